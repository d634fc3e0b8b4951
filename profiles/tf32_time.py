"""bf16 (persistent kernel, launch-per-stage) vs TF32 mode (launch-per-stage): graph-replayed 1024 x 15 rollout, L2 flushed."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamer_b200 import _lib as L, ops, synthetic as W
from dreamer_b200.graphs import StepGraph
B, H = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024), 15
cfg = dict(W.REF_CONFIG, horizon=H)
dev = torch.device("cuda")
lib = L.load()
sd = {k: v.to(dev) for k, v in W.make_state_dict(cfg, seed=0, actor_mu_zero=True).items()}
z0, h0, u, n = (t.to(dev) for t in W.rollout_inputs(cfg, B, H, seed=1))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for label, prec, persist in (("bf16 persistent kernel", "bf16", 1), ("bf16 launch-per-stage", "bf16", 0), ("tf32 launch-per-stage", "tf32", 0)):
    L.check(lib.drm_set_option(b"persist", persist), "opt")
    model = ops.PackedRssm.from_state_dict(sd, precision=prec)
    ro = ops.Rollout(model, B, H)
    g = StepGraph(lambda a, b, c, d: ro.run(a, b, c, d, want_idx=False), warmup=2)
    ts = []
    for i in range(30):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g(z0, h0, u, n); b.record(); torch.cuda.synchronize()
        if i >= 8: ts.append(a.elapsed_time(b))
    ms = statistics.median(ts)
    print(f"{label}: {ms:.4f} ms per rollout ({B} x {H}) = {B * H / ms / 1e3:.2f} M states/s")
L.check(lib.drm_set_option(b"persist", 1), "opt")
