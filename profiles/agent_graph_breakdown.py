"""Kernel-level breakdown of ONE graph-replayed Agent.train_step on a reference-configuration rollout (2500 start states x horizon 30)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from dreamer_b200 import synthetic as W
from dreamer_b200.hotpath import HotPath
dev = torch.device("cuda")
cfg = dict(W.REF_CONFIG, buffer_size=8192)
hp = HotPath(cfg, dev)
rng = np.random.default_rng(1)
n = 4096
hp.buffer.add_batch(rng.integers(0, 256, size=(n, 3, 64, 64)).astype(np.uint8), rng.uniform(-1, 1, (n, 3)).astype(np.float32),
                    rng.standard_normal(n).astype(np.float32), (rng.random(n) > 0.02).astype(np.float32))
hp.world_model.enable_cuda_graphs(1); hp.agent.enable_cuda_graphs(1)
for _ in range(5):
    hp.train_world_model(); hp.train_Agent()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    hp.train_Agent()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_time > 0]
t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
busy = sum(e.device_time for e in evs)
print(f"train_Agent ({cfg['AC_epochs']} epochs: sample + warm start + rollout + Agent.train_step): span {1e-3 * (t1 - t0):.3f} ms, kernel time {1e-3 * busy:.3f} ms in {len(evs)} kernels")
agg = collections.defaultdict(lambda: [0.0, 0])
for e in evs:
    agg[e.name[:80]][0] += e.device_time; agg[e.name[:80]][1] += 1
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:32]:
    print(f"{1e-3 * t:8.3f} ms x{n:5d}  {k}")
