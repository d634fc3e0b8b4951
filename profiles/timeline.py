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
buf = (C.c_uint64 * (8 * 16))()
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
