"""A/B of the backward passes' GEMM backend (bptt.GEMM_BATCHED / GEMM_STEP: this library's drm_gemm_tf32 vs the library GEMM behind
torch.mm): world-model training step (16 x 64) and agent training step (1024 x 15, and 50 x 30 as in car_racer_config.yaml) as
CUDA graphs.  python profiles/gemm_backend_ab.py [wm | iter <batched> <step> <heads> | agent]   (one process per configuration of the
iteration section: capturing many training-step graphs of different models in one process is not what the product does)"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamer_b200 import synthetic as W, bptt
dev = torch.device("cuda")


def t(fn, reps=10, warm=4):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


section = sys.argv[1] if len(sys.argv) > 1 else "wm"
cfg = dict(W.REF_CONFIG, horizon=64, sequence_length=64, batch_size=16)
obs, act, rew, cont, uu = (x.to(dev) for x in W.sequence_inputs(cfg, 16, 64, seed=4321))
for batched, step, heads in ((("torch", "torch", "autograd"), ("drm", "torch", "autograd"), ("drm", "drm", "autograd"), ("drm", "drm", "drm"),
                              ("torch", "torch", "autograd"), ("drm", "drm", "drm")) if section == "wm" else ()):
    bptt.GEMM_BATCHED, bptt.GEMM_STEP, bptt.HEADS_BACKWARD = batched, step, heads
    wm, _ = W.build_learners(cfg, W.make_state_dict(cfg, seed=0), dev)
    wm.enable_cuda_graphs(1)
    g = t(lambda: wm.training_step(obs, act, rew, cont, uniforms=uu))
    print(f"world-model step 16 x 64: batched GEMMs {batched:5s} step GEMMs {step:5s} heads backward {heads:8s}: {g:7.3f} ms ({1e3 / g:6.1f} steps/s)", flush=True)
    del wm

# one training iteration at car_racer_config.yaml (batch 50 x seq 50, horizon 30): world-model step and agent step as graphs
import numpy as np
from dreamer_b200.hotpath import HotPath
for batched, step, heads in ((tuple(sys.argv[2:5]),) if section == "iter" else ()):
    bptt.GEMM_BATCHED, bptt.GEMM_STEP, bptt.HEADS_BACKWARD = batched, step, heads
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
    obs, act, rew, cont, L = hp.buffer.sample_sequences(batch_size=hp.batch_size)
    z0, h0 = hp.warm_start_generator(obs, act, L)
    z, h, a, r, c, mu, sg = hp.dream_episodes(z0, h0)
    t_wm = t(lambda: hp.world_model.training_step(obs, act, rew, cont))
    t_ag = t(lambda: hp.agent.train_step(z, h, r, c, a, mu, sg))
    print(f"car_racer_config (50 x 50, horizon 30): batched {batched:5s} step {step:5s} heads {heads:8s}: world-model step {t_wm:7.3f} ms, agent step {t_ag:7.3f} ms", flush=True)
    del hp


# agent training step at 1024 start states x horizon 15 (c5's agent step): the step GEMMs have 1024 gradient rows
cfg = dict(W.REF_CONFIG, horizon=15, batch_size=1024)
bptt.GEMM_BATCHED, bptt.GEMM_STEP, bptt.HEADS_BACKWARD = "drm", "drm", "drm"
for large in (("torch", "drm", "torch", "drm") if section == "agent" else ()):
    bptt.GEMM_STEP_LARGE = large
    wm, ag = W.build_learners(cfg, W.make_state_dict(cfg, seed=0), dev)
    ag.attach_world_model(wm)
    g = torch.Generator(device=dev).manual_seed(3)
    B, H, Dh = 1024, 15, cfg["hidden_state_dims"]
    z = torch.nn.functional.one_hot(torch.randint(0, 32, (B, H + 1, 32), device=dev, generator=g), 32).float()
    h = torch.tanh(torch.randn(B, H + 1, Dh, device=dev, generator=g))
    a = torch.tanh(torch.randn(B, H, 3, device=dev, generator=g))
    mu = torch.randn(B, H, 3, device=dev, generator=g) * 0.3
    sg = torch.rand(B, H, 3, device=dev, generator=g) + 0.1
    r = torch.randn(B, H, 1, device=dev, generator=g)
    c = torch.ones(B, H, 1, device=dev)
    ag.enable_cuda_graphs(1)
    t_ag = t(lambda: ag.train_step(z, h, r, c, a, mu, sg))
    print(f"agent step 1024 x 15: step GEMMs above 64 rows on {large:5s}: {t_ag:7.3f} ms", flush=True)
    del wm, ag
