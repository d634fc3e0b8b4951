"""Driver for ncu captures of the HBM-bound kernels: replay gather (1024 x 50 windows) and the stand-alone categorical (4 Mi rows)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamer_b200 import ops
dev = torch.device("cuda")
cap, B, L = 60000, 1024, 50
ring = [torch.randint(0, 256, (cap, 3, 64, 64), dtype=torch.uint8, device=dev), torch.rand(cap, 3, device=dev),
        torch.rand(cap, 1, device=dev), torch.ones(cap, 1, device=dev)]
starts = torch.randint(0, cap - L, (B,), device=dev)
for _ in range(3):
    out = ops.replay_gather(*ring, starts, L)
n = 1 << 22
lg = torch.randn(n, 32, device=dev); u = torch.rand(n, device=dev)
for _ in range(3):
    ops.categorical32(lg, u)
torch.cuda.synchronize()
print("done")
