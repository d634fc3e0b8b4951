import os, sys, statistics
sys.path.insert(0, "/root/repo")
import torch
from dreamer_b200 import ops, _lib as L
torch.backends.cuda.matmul.allow_tf32 = True
dev = torch.device("cuda")
def graph_time(fn, n=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3 / n)
    return statistics.median(ts)
def mm_flags(a, b, extra, a_direct=True):
    at, lda, ta = ops._gemm_operand(a); bt, ldb, tb = ops._gemm_operand(b)
    M, K = at.shape; N = bt.shape[0]
    out = torch.empty(M, N, device=dev)
    lib = L.load(); ws = ops._gemm_workspace(dev, lib.drm_gemm_tf32_workspace_bytes(M, N, K))
    flags = (1 if ta else 0) | (2 if tb else 0) | (8 if a_direct else 0) | 16 | extra
    L.check(lib.drm_gemm_tf32(M, N, K, L.ptr(at), lda, L.ptr(bt), ldb, L.ptr(out), out.stride(0), None, flags, L.ptr(ws), ws.numel(), L.stream()), "g")
    return out
for (M, N, K) in ((1024, 200, 64), (1024, 200, 1624), (1024, 600, 1800), (16, 600, 1800), (16, 256, 1024), (50, 1024, 1800)):
    a, b = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
    ref = mm_flags(a, b, 32, M > 64)
    assert torch.equal(ref, mm_flags(a, b, 0, M > 64))
    print(f"M={M} N={N} K={K}: cuBLAS {graph_time(lambda: torch.mm(a, b.t())):5.1f} us | ours no PDL {graph_time(lambda: mm_flags(a, b, 32, M > 64)):5.1f} | PDL {graph_time(lambda: mm_flags(a, b, 0, M > 64)):5.1f}", flush=True)
