"""Eager launches vs one CUDA-graph replay of a whole rollout (device-resident inputs)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamer_b200 import ops, synthetic as W
from dreamer_b200.graphs import StepGraph
B, H = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024), 15
cfg = dict(W.REF_CONFIG, horizon=H)
dev = torch.device("cuda")
model = ops.PackedRssm.from_state_dict({k: v.to(dev) for k, v in W.make_state_dict(cfg, seed=0, actor_mu_zero=True).items()})
ro = ops.Rollout(model, B, H)
z0, h0, u, n = (t.to(dev) for t in W.rollout_inputs(cfg, B, H, seed=1))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = StepGraph(lambda a, b, c, d: ro.run(a, b, c, d, want_idx=False), warmup=2)
for name, fn in (("eager", lambda: ro.run(z0, h0, u, n, want_idx=False)), ("graph", lambda: g(z0, h0, u, n))):
    ts = []
    for i in range(30):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 8:
            ts.append(a.elapsed_time(b))
    print(f"{name}: {statistics.median(ts):.4f} ms per rollout ({B} x {H})")
