"""Per-stage CUDA-event times of the rollout for a list of batch sizes (profiling mode: PDL off)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamer_b200 import ops, _lib as L, synthetic as W
H = 15
cfg = dict(W.REF_CONFIG, horizon=H)
sd = {k: v.cuda() for k, v in W.make_state_dict(cfg, seed=0, actor_mu_zero=True).items()}
model = ops.PackedRssm.from_state_dict(sd)
lib = L.load()
names = ["gru", "prior_l1", "prior_l2", "prior_cat", "heads_l1", "heads_l2", "heads_out", "other"]
for B in [int(x) for x in sys.argv[1:]] or [1024]:
    ro = ops.Rollout(model, B, H)
    z0, h0, u, n = (t.cuda() for t in W.rollout_inputs(cfg, B, H, seed=1))
    for _ in range(3):
        ro.run(z0, h0, u, n, want_idx=False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        ro.run(z0, h0, u, n, want_idx=False)
    b.record(); torch.cuda.synchronize()
    total = a.elapsed_time(b) / 5
    lib.drm_profile_enable(1)
    for _ in range(5):
        ro.run(z0, h0, u, n, want_idx=False)
    torch.cuda.synchronize()
    lib.drm_profile_enable(0)
    out = {}
    for i, nm in enumerate(names):
        ms, cnt = C.c_double(), C.c_int64()
        lib.drm_profile_read(i, C.byref(ms), C.byref(cnt))
        if cnt.value:
            out[nm] = round(1e3 * ms.value / cnt.value, 1)
    print(f"B={B:6d} rollout {total*1e3:8.1f} us  ({B*H/total/1e3:7.2f} M states/s)  per-launch us: {out}")
