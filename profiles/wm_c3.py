"""World-model loss forward at BASELINE configs[2] (batch 16 x seq 64) -- driver for ncu launch lists / timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamer_b200 import synthetic as W
B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 64)
cfg = dict(W.REF_CONFIG, horizon=T, sequence_length=T, batch_size=B)
dev = torch.device("cuda")
wm, _ = W.build_learners(cfg, W.make_state_dict(cfg, seed=0), dev)
obs, act, rew, cont, u = (x.to(dev) for x in W.sequence_inputs(cfg, B, T, seed=4321))
for _ in range(3):
    wm.loss_forward(obs, act, rew, cont, uniforms=u)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); a.record(); total, _ = wm.loss_forward(obs, act, rew, cont, uniforms=u); b.record(); t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"wm loss forward B={B} T={T}: host enqueue {1e3*(t1-t0):.2f} ms, device {a.elapsed_time(b):.2f} ms, loss {total.item():.4f}")
import statistics
ts = []
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); wm.loss_forward(obs, act, rew, cont, uniforms=u); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"median of 20: device {statistics.median(ts):.3f} ms (min {min(ts):.3f})")
