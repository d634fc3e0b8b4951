"""Four representative drm_gemm_tf32 calls for an ncu capture (profiles/prof_r2_gemm_tf32_summary.txt):
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_kernel -c 4 -o gpurun_out/r2e_gemm python profiles/gemm_prof.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamer_b200 import ops
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g)
dGI, X = r(1024, 1800), r(1024, 1028)[:, :1027]
W = r(1800, 1027)
out = torch.zeros(1800, 1027, device=dev)
ops.mm_nt(dGI.t(), X.t(), out=out, accumulate=True, a_direct=True, b_direct=True)   # weight gradient, X unaligned -> packed; dGI read K-last in place
Hp, Whh = r(1024, 600), r(1800, 600)
ops.mm_nt(Hp, Whh, r(1800), a_direct=True, b_direct=True)                           # forward re-evaluation, both in place (K-first), split-K tickets
dgh, Whh_r = r(16, 1800), ops.pack_tf32(Whh)
ops.mm(dgh, Whh_r, b_direct=True)                                                    # one BPTT step: swapped, cluster split-K, weight read K-last in place
dgi50, Wz = r(50, 1800), ops.pack_tf32(r(1800, 1024))
ops.mm(dgi50, Wz, b_direct=True)                                                     # 50 sequences (car_racer_config batch)
torch.cuda.synchronize()
print("ok")
