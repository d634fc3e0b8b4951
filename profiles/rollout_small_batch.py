import os, sys, statistics
sys.path.insert(0, "/root/repo")
import torch
from dreamer_b200 import _lib as L, ops, synthetic as W
dev = torch.device("cuda")
for B, H in ((16, 30), (50, 30), (64, 30), (128, 30)):
    cfg = dict(W.REF_CONFIG, horizon=H)
    sd = {k: v.to(dev) for k, v in W.make_state_dict(cfg, seed=0).items()}
    model = ops.PackedRssm.from_state_dict(sd)
    ro = ops.Rollout(model, B, H)
    z0, h0, u, n = (t.to(dev) for t in W.rollout_inputs(cfg, B, H, seed=1))
    for _ in range(3): ro.run(z0, h0, u, n, want_idx=False)
    ts = []
    for i in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ro.run(z0, h0, u, n, want_idx=False); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"B={B} H={H}: {statistics.median(ts):.4f} ms ({1e3 * statistics.median(ts) / H:.1f} us per step) info {ro.info()['ctas']} CTAs")
