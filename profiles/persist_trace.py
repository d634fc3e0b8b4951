"""Per-tile timeline of the persistent rollout kernel: python profiles/persist_trace.py [rows] [horizon] [state] [m_tile]

Records the tiles of two consecutive states (drm_rollout_trace) and prints, for one m-tile, every tile that touches it in
time order: microseconds relative to the first record -- start of the tile on its CTA, dependency seen by the producer,
first operands landed (MMA can start), epilogue warps past their own waits, accumulator complete, epilogue math done, outputs
published.
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dreamer_b200 import _lib as L, ops, synthetic as W

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
H = int(sys.argv[2]) if len(sys.argv) > 2 else 15
J0 = int(sys.argv[3]) if len(sys.argv) > 3 else 6
MT = int(sys.argv[4]) if len(sys.argv) > 4 else 0
cfg = dict(W.REF_CONFIG, horizon=H)
dev = "cuda"
lib = L.load()
model = ops.PackedRssm.from_state_dict({k: v.to(dev) for k, v in W.make_state_dict(cfg, seed=0, actor_mu_zero=True).items()})
ro = ops.Rollout(model, B, H)
z0, h0, u, n = (t.to(dev) for t in W.rollout_inputs(cfg, B, H, seed=1))
info = ro.info()
print("info", info)
assert info["persistent"]
for _ in range(3):
    ro.run(z0, h0, u, n, want_idx=False)
torch.cuda.synchronize()
L.check(lib.drm_rollout_trace(ro.handle, J0, 2, None, 0), "trace on")
ro.run(z0, h0, u, n, want_idx=False)
torch.cuda.synchronize()
nw = info["ctas"] * 32 * 8 * 2
buf = np.zeros(nw, dtype=np.uint64)
L.check(lib.drm_rollout_trace(ro.handle, 0, 0, buf.ctypes.data_as(C.c_void_p), nw), "trace read")
rec = buf[: nw // 2].reshape(info["ctas"], 32, 8)
laps = buf[nw // 2:].reshape(info["ctas"], 32, 8)
KIND = {0: "chain", 2: "rc", 3: "gru"}
LAYER = {0: "prior1", 1: "prior2", 2: "sample", 8: "actor1", 9: "actor2", 10: "actor3"}
rows = []
for cta in range(rec.shape[0]):
    for s in range(32):
        r = rec[cta, s]
        if r[1] == 0:
            continue
        code = int(r[0])
        kind, layer, j, m = code >> 24, (code >> 16) & 15, (code >> 8) & 255, code & 255
        rows.append((int(r[1]), kind, layer, j, m, cta, [int(x) for x in r[1:8]], [int(x) for x in laps[cta, s]]))
t0 = min(r[0] for r in rows)
print(f"{len(rows)} tiles recorded; us relative to the first")
print(f"{'tile':<18}{'cta':>4} {'start':>8}{'dep':>8}{'operands':>9}{'epi rdy':>8}{'acc':>8}{'epi end':>8}{'publ':>8}   main  epi  publ")
gru_done = {}
for t_start, kind, layer, j, m, cta, ts, lp in sorted(rows):
    if m != MT:
        continue
    rel = [(x - t0) / 1e3 if x else float("nan") for x in ts]
    name = f"{LAYER.get(layer, layer) if kind == 0 else KIND[kind] + str(layer)} j={j}"
    if kind == 3:
        gru_done.setdefault(j, []).append(rel)
        if cta % 8 == 0 and any(lp):
            print(f"  gru cta {cta} j={j} acc {rel[4]:.2f} | action seen, gates done, barrier, copied out (absolute us): " + " ".join(f"{(x - t0) / 1e3:.2f}" for x in lp if x) + f" | published {rel[6]:.2f}")
        continue
    print(f"{name:<18}{cta:>4} " + "".join(f"{x:>8.2f}" if i != 2 else f"{x:>9.2f}" for i, x in enumerate(rel)) +
          f"   {rel[4] - rel[2]:5.2f} {rel[5] - rel[4]:4.2f} {rel[6] - rel[5]:4.2f}" +
          ("   laps(abs) " + " ".join(f"{(x - t0) / 1e3:.2f}" for x in lp if x) if any(lp) else ""))
for j, lst in sorted(gru_done.items()):
    a = np.array(lst)
    print(f"gru j={j}: {len(lst)} tiles; start {a[:,0].min():.2f}..{a[:,0].max():.2f}  dep {a[:,1].min():.2f}..{a[:,1].max():.2f}  operands {a[:,2].min():.2f}..{a[:,2].max():.2f}  "
          f"epi rdy {a[:,3].min():.2f}..{a[:,3].max():.2f}  acc {a[:,4].min():.2f}..{a[:,4].max():.2f}  epi end {a[:,5].min():.2f}..{a[:,5].max():.2f}  publ {a[:,6].min():.2f}..{a[:,6].max():.2f}  "
          f"main loop {np.nanmean(a[:,4]-a[:,2]):.2f} (mean)")
