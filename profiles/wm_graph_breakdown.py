"""Kernel-level breakdown of ONE graph-replayed world-model training step (batch 16 x seq 64): device time per kernel family and the
idle share of the step (torch profiler / CUPTI sees the kernels inside a graph launch)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from dreamer_b200 import synthetic as W
cfg = dict(W.REF_CONFIG, horizon=64, sequence_length=64, batch_size=16)
dev = torch.device("cuda")
wm, _ = W.build_learners(cfg, W.make_state_dict(cfg, seed=0), dev)
obs, act, rew, cont, uu = (x.to(dev) for x in W.sequence_inputs(cfg, 16, 64, seed=4321))
wm.enable_cuda_graphs(1)
for _ in range(6):
    wm.training_step(obs, act, rew, cont, uniforms=uu)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); wm.training_step(obs, act, rew, cont, uniforms=uu); b.record(); torch.cuda.synchronize()
print(f"graph-replayed step: {a.elapsed_time(b):.3f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    wm.training_step(obs, act, rew, cont, uniforms=uu)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_time > 0]
t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
busy = sum(e.device_time for e in evs)
print(f"span {1e-3 * (t1 - t0):.3f} ms, kernel time {1e-3 * busy:.3f} ms in {len(evs)} kernels ({100 * busy / (t1 - t0):.0f} % busy)")
agg = collections.defaultdict(lambda: [0.0, 0])
for e in evs:
    k = e.name[:70]
    agg[k][0] += e.device_time; agg[k][1] += 1
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"{1e-3 * t:8.3f} ms x{n:5d}  {k}")
