"""A/B of the conv path: implicit GEMM (im2col-mode TMA) vs patch gather + GEMM.  python profiles/conv_ab.py [frames]
Times encoder convs + head (Observe.encode) and decoder (Observe.decode) at `frames` frames, and the world-model loss forward."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamer_b200 import _lib as L, ops, synthetic as W
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cfg = dict(W.REF_CONFIG)
dev = "cuda"
lib = L.load()
sd = {k: v.to(dev) for k, v in W.make_state_dict(cfg, seed=0).items()}
model = ops.PackedRssm.from_state_dict(sd)
vae = ops.PackedVae.from_state_dict(model, sd, (64, 64))
ws = ops.Observe(vae, N, 1)
g = torch.Generator(device=dev).manual_seed(1)
h = torch.tanh(torch.randn(N, 600, device=dev, generator=g))
obs = torch.rand(N, 3, 64, 64, device=dev, generator=g) - 0.5
z = torch.nn.functional.one_hot(torch.randint(0, 32, (N, 32), device=dev, generator=g), 32).float().reshape(N, 1024)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=20):
    ts = []
    for i in range(n + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b))
    return statistics.median(ts)
cfg3 = dict(W.REF_CONFIG, horizon=64, sequence_length=64, batch_size=16)
wm, _ = W.build_learners(cfg3, W.make_state_dict(cfg3, seed=0), torch.device(dev))
o3, a3, r3, c3, u3 = (x.to(dev) for x in W.sequence_inputs(cfg3, 16, 64, seed=4321))
for flag, kps in ((1, 0), (1, 1), (1, 2), (1, 4), (0, 0)):
    L.check(lib.drm_set_option(b"conv_implicit", flag), "opt"); L.check(lib.drm_set_option(b"conv_kps", kps), "opt")
    te = timeit(lambda: ws.encode(h, obs))
    td = timeit(lambda: ws.decode(h, z))
    tl = timeit(lambda: wm.loss_forward(o3, a3, r3, c3, uniforms=u3))
    print(f"conv_implicit={flag} conv_kps={kps}: encode {N} frames {te:.3f} ms, decode {td:.3f} ms, world-model loss forward (16 x 64) {tl:.3f} ms")
