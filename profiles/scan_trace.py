"""Per-tile timeline of the persistent posterior-scan kernel: python profiles/scan_trace.py [B] [T] [step]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dreamer_b200 import _lib as L, ops, synthetic as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
J0 = int(sys.argv[3]) if len(sys.argv) > 3 else 20
cfg = dict(W.REF_CONFIG, horizon=T, sequence_length=T, batch_size=B)
dev = "cuda"
lib = L.load()
sd = {k: v.to(dev) for k, v in W.make_state_dict(cfg, seed=0).items()}
model = ops.PackedRssm.from_state_dict(sd)
vae = ops.PackedVae.from_state_dict(model, sd, (64, 64))
ws = ops.Observe(vae, B, T)
obs, act, rew, cont, u = (x.to(dev) for x in W.sequence_inputs(cfg, B, T, seed=4321))
obs = obs / 255.0 - 0.5
for _ in range(3): ws.scan(obs, act, u)
torch.cuda.synchronize()
L.check(lib.drm_observe_trace(ws.handle, J0, 2, None, 0), "trace on")
ws.scan(obs, act, u); torch.cuda.synchronize()
n_cta = 4 * ((B + 127) // 128) + ((B + 127) // 128) * 19
n_cta = (n_cta + 3) // 4 * 4
nw = n_cta * 32 * 8 * 2
buf = np.zeros(nw, dtype=np.uint64)
L.check(lib.drm_observe_trace(ws.handle, 0, 0, buf.ctypes.data_as(C.c_void_p), nw), "trace read")
rec = buf[: nw // 2].reshape(n_cta, 32, 8)
rows = []
for cta in range(n_cta):
    for s in range(32):
        r = rec[cta, s]
        if r[1] == 0: continue
        code = int(r[0]); role, layer, j, m = (code >> 24) - 4, (code >> 16) & 15, (code >> 8) & 255, code & 255
        rows.append((int(r[1]), role, layer, j, m, cta, [int(x) for x in r[1:8]]))
t0 = min(r[0] for r in rows)
print(f"{'tile':<14}{'cta':>4} {'start':>8}{'dep':>8}{'operands':>9}{'epi rdy':>8}{'acc':>8}{'epi end':>8}{'publ':>8}")
gru = {}
for t_start, role, layer, j, m, cta, ts in sorted(rows):
    rel = [(x - t0) / 1e3 if x else float('nan') for x in ts]
    if role == 3:
        gru.setdefault(j, []).append(rel); continue
    name = {0: "postL1", 2: "sample"}.get(layer, str(layer)) + f" j={j}"
    print(f"{name:<14}{cta:>4} " + "".join(f"{x:>8.2f}" if i != 2 else f"{x:>9.2f}" for i, x in enumerate(rel)))
for j, lst in sorted(gru.items()):
    a = np.array(lst)
    print(f"gru s={j}: {len(lst)} tiles; start {a[:,0].min():.2f}..{a[:,0].max():.2f} dep {a[:,1].min():.2f}..{a[:,1].max():.2f} operands {a[:,2].min():.2f}..{a[:,2].max():.2f} "
          f"acc {np.nanmin(a[:,4]):.2f}..{np.nanmax(a[:,4]):.2f} epi end {a[:,5].min():.2f}..{a[:,5].max():.2f} publ {a[:,6].min():.2f}..{a[:,6].max():.2f}")
