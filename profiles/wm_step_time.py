"""World-model training step at BASELINE configs[2] (batch 16 x seq 64): eager vs CUDA graph, BPTT vs autograd tail."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamer_b200 import synthetic as W
cfg = dict(W.REF_CONFIG, horizon=64, sequence_length=64, batch_size=16)
dev = torch.device("cuda")
obs, act, rew, cont, uu = (x.to(dev) for x in W.sequence_inputs(cfg, 16, 64, seed=4321))


def t(fn, reps=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


for mode in ("autograd", "bptt"):
    wm, _ = W.build_learners(cfg, W.make_state_dict(cfg, seed=0), dev)
    wm.grad_mode = mode
    e = t(lambda: wm.training_step(obs, act, rew, cont, uniforms=uu))
    wm.enable_cuda_graphs(1)
    g = t(lambda: wm.training_step(obs, act, rew, cont, uniforms=uu))
    print(f"{mode:9s} eager {e:7.2f} ms   graph {g:7.2f} ms   ({1e3/g:.1f} steps/s)")
    del wm

# phase times of the BPTT gradient (eager, device time between phase marks)
from dreamer_b200 import bptt
wm, _ = W.build_learners(cfg, W.make_state_dict(cfg, seed=0), dev)
for _ in range(3):
    wm.training_step(obs, act, rew, cont, uniforms=uu)
a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a0.record()
total, parts = wm.loss_forward(obs, act, rew, cont, uu)
a1.record()
marks = []
wm.optimiser.zero_grad()
bptt.world_model_backward(wm, parts["obs_norm"], act, rew, cont, wm.last["scan"]["idx"], wm.last["scan"]["hidden"], parts, marks=marks)
b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
b0.record(); wm.optimiser.step(); b1.record()
torch.cuda.synchronize()
print(f"  {a0.elapsed_time(a1):7.2f} ms  loss forward (scan + heads kernels)")
for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
    print(f"  {e0.elapsed_time(e1):7.2f} ms  {n1}")
print(f"  {b0.elapsed_time(b1):7.2f} ms  clip + AdamW (fused)")
