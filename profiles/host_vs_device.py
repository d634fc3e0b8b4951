"""Host enqueue time vs device time of one rollout (diagnostic used for profiles/)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamer_b200 import ops
from dreamer_b200 import synthetic as W
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
B, H, over = {"c2": (1024, 15, {}), "c2x16": (16384, 15, {}), "c4": (16384, 15, {"hidden_state_dims": 4096})}[wl]
cfg = dict(W.REF_CONFIG, horizon=H, **over)
sd = W.make_state_dict(cfg, seed=0, actor_mu_zero=True)
dev = torch.device("cuda")
model = ops.PackedRssm.from_state_dict({k: v.to(dev) for k, v in sd.items()})
ro = ops.Rollout(model, B, H)
z0, h0, u, n = (t.to(dev) for t in W.rollout_inputs(cfg, B, H, seed=1234))
for _ in range(3):
    ro.run(z0, h0, u, n, want_idx=False)
torch.cuda.synchronize()
for i in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record(); ro.run(z0, h0, u, n, want_idx=False); b.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{wl}: host enqueue {1e3*(t1-t0):.3f} ms, device {a.elapsed_time(b):.3f} ms, wall {1e3*(t2-t0):.3f} ms")
