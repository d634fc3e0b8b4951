import os, sys
sys.path.insert(0, "/root/repo")
import torch
from torch.profiler import profile, ProfilerActivity
from dreamer_b200 import _lib as L, ops, synthetic as W
B, T = int(sys.argv[1]), int(sys.argv[2])
cfg = dict(W.REF_CONFIG, horizon=T, sequence_length=T, batch_size=B)
dev = "cuda"
sd = {k: v.to(dev) for k, v in W.make_state_dict(cfg, seed=0).items()}
model = ops.PackedRssm.from_state_dict(sd)
vae = ops.PackedVae.from_state_dict(model, sd, (64, 64))
ws = ops.Observe(vae, B, T)
obs, act, rew, cont, u = (x.to(dev) for x in W.sequence_inputs(cfg, B, T, seed=4321))
obs = obs / 255.0 - 0.5
for _ in range(3): ws.scan(obs, act, u)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ws.scan(obs, act, u); torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
    print(f"{e.device_time_total/1e3:8.3f} ms x{e.count:4d} {e.key[:90]}")
