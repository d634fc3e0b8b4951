"""A/B a drm_set_option on whole rollouts: python profiles/rollout_ab.py <rows> <option> <value_a> <value_b> [horizon]."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamer_b200 import ops, _lib as L, synthetic as W
B, opt, va, vb = int(sys.argv[1]), sys.argv[2].encode(), int(sys.argv[3]), int(sys.argv[4])
H = int(sys.argv[5]) if len(sys.argv) > 5 else 15
cfg = dict(W.REF_CONFIG, horizon=H)
dev = torch.device("cuda")
model = ops.PackedRssm.from_state_dict({k: v.to(dev) for k, v in W.make_state_dict(cfg, seed=0, actor_mu_zero=True).items()})
ro = ops.Rollout(model, B, H)
z0, h0, u, n = (t.to(dev) for t in W.rollout_inputs(cfg, B, H, seed=1))
lib = L.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for v in (va, vb, va, vb):
    L.check(lib.drm_set_option(opt, v), "set_option")
    ts = []
    for i in range(25):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ro.run(z0, h0, u, n, want_idx=False); b.record(); torch.cuda.synchronize()
        if i >= 5:
            ts.append(a.elapsed_time(b))
    print(f"{opt.decode()}={v}: {statistics.median(ts) * 1e3 / H:7.1f} us per imagined step ({B} rows)")
