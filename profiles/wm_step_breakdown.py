"""Where the world-model training step's GPU time goes (torch profiler over one eager step at BASELINE configs[2])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from dreamer_b200 import synthetic as W
cfg = dict(W.REF_CONFIG, horizon=64, sequence_length=64, batch_size=16)
dev = torch.device("cuda")
wm, _ = W.build_learners(cfg, W.make_state_dict(cfg, seed=0), dev)
obs, act, rew, cont, uu = (x.to(dev) for x in W.sequence_inputs(cfg, 16, 64, seed=4321))
for _ in range(3):
    wm.training_step(obs, act, rew, cont, uniforms=uu)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    wm.training_step(obs, act, rew, cont, uniforms=uu)
    torch.cuda.synchronize()
evs = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.device_time_total for e in evs)
n = sum(e.count for e in evs)
print(f"total device time {tot/1e3:.2f} ms in {n} kernels")
for e in sorted(evs, key=lambda e: -e.device_time_total)[:28]:
    print(f"{e.device_time_total/1e3:8.3f} ms  x{e.count:5d}  {e.key[:110]}")
