"""Where one training iteration at the reference configuration goes (CUDA-event times around each call of HotPath's loops)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
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
for _ in range(4):
    hp.train_world_model(); hp.train_Agent()
torch.cuda.synchronize()
acc = {}


def timed(name, fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = fn(); b.record(); torch.cuda.synchronize()
    acc.setdefault(name, []).append(a.elapsed_time(b))
    return out


for _ in range(5):
    obs, act, rew, cont, _ = timed("sample (WM)", lambda: hp.buffer.sample_sequences(batch_size=hp.batch_size))
    timed("WorldModel.training_step", lambda: hp.world_model.training_step(obs, act, rew, cont))
    obs, act, _, _, L = timed("sample (AC)", lambda: hp.buffer.sample_sequences(batch_size=hp.batch_size))
    z0, h0 = timed("warm_start_generator", lambda: hp.warm_start_generator(obs, act, L))
    out = timed("dream_episodes", lambda: hp.dream_episodes(z0, h0))
    z, h, a, r, c, mu, sg = out
    timed("Agent.train_step", lambda: hp.agent.train_step(z, h, r, c, a, mu, sg))
for k, v in acc.items():
    print(f"{statistics.median(v):8.2f} ms  {k}")
