"""In-kernel timeline of CTA (0,0) of each fused stage (last launch of a rollout).  Diagnostic."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamer_b200 import ops, _lib as L
from dreamer_b200 import synthetic as W
B, H = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 15
cfg = dict(W.REF_CONFIG, horizon=H)
sd = W.make_state_dict(cfg, seed=0, actor_mu_zero=True)
dev = torch.device("cuda")
model = ops.PackedRssm.from_state_dict({k: v.to(dev) for k, v in sd.items()})
ro = ops.Rollout(model, B, H)
z0, h0, u, n = (t.to(dev) for t in W.rollout_inputs(cfg, B, H, seed=1234))
for _ in range(3):
    ro.run(z0, h0, u, n, want_idx=False)
torch.cuda.synchronize()
lib = L.load()
lib.drm_debug_timeline(1, None)
ro.run(z0, h0, u, n, want_idx=False)
buf = (C.c_uint64 * (8 * (16 + 1024)))()
lib.drm_debug_timeline(0, buf)
names = ["gru", "prior_l1", "prior_l2", "prior_cat", "heads_l1", "heads_l2", "heads_out", "other"]
pts = ["entry", "setup", "tma0", "ops0", "mma_end", "acc_rdy", "epi_end", "freed"]
for s, nm in enumerate(names):
    t = [buf[s * 16 + 2 * i] for i in range(8)]
    c = [buf[s * 16 + 2 * i + 1] for i in range(8)]
    if t[0] == 0:
        continue
    print(f"{nm:10s} ns since entry: " + "  ".join(f"{p}={t[i]-t[0]:6d}" for i, p in enumerate(pts)))
    print(f"{'':10s} cycles       : " + "  ".join(f"{p}={c[i]-c[0]:6d}" for i, p in enumerate(pts)))

# chained kernels (option "chain"): 15 probes starting at the prior_l1 / heads_l1 slots
cpts = ["entry", "setup", "staged", "acc0", "lstat0", "xstat0", "y0", "acc1", "lstat1", "xstat1", "y1", "acc2", "epi_end", "synced", "exit", "l1_math", "l1_fence", "l1_bar", "mma_y0", "mma_l1", "mma_y1", "mma_l2"]
for s, nm in ((1, "prior chain"), (4, "heads chain")):
    t = [buf[s * 16 + 2 * i] for i in range(22)]
    c = [buf[s * 16 + 2 * i + 1] for i in range(22)]
    if t[8] == 0:
        continue
    print(f"{nm} ns: " + "  ".join(f"{p}={t[i]-t[0]}" for i, p in enumerate(cpts)))
    print(f"{nm} cy: " + "  ".join(f"{p}={c[i]-c[0]}" for i, p in enumerate(cpts)))

# per-CTA records of every stage: when did each CTA enter, get past its dependency wait, and exit (relative to the GRU stage's first entry)
recs = {}
for s, nm in enumerate(names):
    base = 8 * 16 + 1024 * s
    rows = [(buf[base + 4 * i], buf[base + 4 * i + 1], buf[base + 4 * i + 2], buf[base + 4 * i + 3]) for i in range(256) if buf[base + 4 * i]]
    if rows:
        recs[nm] = rows
if recs:
    t0 = min(r[0] for rows in recs.values() for r in rows)
    for nm, rows in recs.items():
        ent = [r[0] - t0 for r in rows]; dep = [r[1] - t0 for r in rows]; ex = [r[2] - t0 for r in rows]
        print(f"{nm:10s} ctas={len(rows):3d} sms={len(set(r[3] for r in rows)):3d} entry {min(ent)/1e3:7.1f}..{max(ent)/1e3:7.1f} us  dep-over {min(dep)/1e3:7.1f}..{max(dep)/1e3:7.1f}  exit {min(ex)/1e3:7.1f}..{max(ex)/1e3:7.1f}")
