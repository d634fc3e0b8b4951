"""Time the CTA-pair GRU stage of one library build: python profiles/pair_sweep.py <lib.so> [D ...]  (16 384 rows x 15).
Used with the -DDRM_PAIR_KPS64 / ST64 / KPS32 / ST32 builds of csrc/gru_pair.cuh (pipeline shape sweep, round 2)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamer_b200 import _lib as L
L.LIB_PATH = os.path.abspath(sys.argv[1])
from dreamer_b200 import ops, synthetic as W
H, B = 15, 16384
lib = L.load()
for D in [int(x) for x in sys.argv[2:]] or [600, 4096]:
    cfg = dict(W.REF_CONFIG, horizon=H, hidden_state_dims=D)
    sd = {k: v.cuda() for k, v in W.make_state_dict(cfg, seed=0, actor_mu_zero=True).items()}
    model = ops.PackedRssm.from_state_dict(sd)
    ro = ops.Rollout(model, B, H)
    z0, h0, u, n = (t.cuda() for t in W.rollout_inputs(cfg, B, H, seed=1))
    for _ in range(2):
        ro.run(z0, h0, u, n, want_idx=False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        ro.run(z0, h0, u, n, want_idx=False)
    b.record(); torch.cuda.synchronize()
    total = a.elapsed_time(b) / 3
    lib.drm_profile_enable(1)
    for _ in range(3):
        ro.run(z0, h0, u, n, want_idx=False)
    torch.cuda.synchronize()
    lib.drm_profile_enable(0)
    ms, cnt = C.c_double(), C.c_int64()
    lib.drm_profile_read(0, C.byref(ms), C.byref(cnt))
    gru_us = 1e3 * ms.value / cnt.value
    ZP = 1024 + 64
    flop = 2.0 * B * (ZP + ((D + 63) // 64) * 64) * 3 * D
    print(f"{os.path.basename(sys.argv[1])} D={D}: rollout {total:8.3f} ms ({B*H/total/1e3:6.2f} M states/s)  GRU stage {gru_us:8.1f} us = {flop/gru_us/1e6:7.1f} TFLOP/s", flush=True)
    del ro, model
