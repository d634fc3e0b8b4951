"""drm_gemm_tf32 against torch.mm (cuBLAS, TF32 allowed) on the shapes of the backward passes: python profiles/gemm_tf32_time.py"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamer_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = True
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=20, flush_l2=True):
    ts = []
    for i in range(reps + 3):
        if flush_l2:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)


def graph_time(fn, n=20):
    """per-call time of n back-to-back calls replayed as one CUDA graph (what the training step sees)"""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    g.replay(); torch.cuda.synchronize()
    return timeit(g.replay, reps=10, flush_l2=False) / n


cases = [  # name, M, N, K, kind
    ("fwd  GI = X W_ih^T      ", 1024, 1800, 1027, "nt"),
    ("fwd  A1 = X1 W1^T       ", 1024, 256, 1624, "nt"),
    ("dgrad dA1 W1            ", 1024, 1624, 256, "nn"),
    ("wgrad dGI^T X           ", 1800, 1027, 1024, "tn"),
    ("wgrad dGH^T Hprev       ", 1800, 600, 1024, "tn"),
    ("wgrad dA1^T X1          ", 256, 1624, 1024, "tn"),
    ("step  dGH[t] W_hh (16)  ", 16, 600, 1800, "nn"),
    ("step  dLG[t] W2 (16)    ", 16, 256, 1024, "nn"),
    ("step  dA1 W1h (16)      ", 16, 600, 256, "nn"),
    ("step  dGI[t] Wih_z (16) ", 16, 1024, 1800, "nn"),
    ("step  actor (1024 rows) ", 1024, 600, 1800, "nn"),
]
for name, M, N, K, kind in cases:
    if kind == "nt":
        a, b = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
        ours = lambda: ops.mm_nt(a, b)
        lib = lambda: torch.mm(a, b.t())
    elif kind == "nn":
        a, b = torch.randn(M, K, device=dev), torch.randn(K, N, device=dev)
        ours = lambda: ops.mm(a, b)
        lib = lambda: torch.mm(a, b)
    else:
        a, b = torch.randn(K, M, device=dev), torch.randn(K, N, device=dev)
        ours = lambda: ops.mm_nt(a.t(), b.t())
        lib = lambda: torch.mm(a.t(), b)
    # weights pre-rounded once (drm_pack_tf32) and read in place -- the way bptt.py calls the input-gradient GEMMs
    if kind == "nn":
        bp = ops.pack_tf32(b)                      # [K, N] row-major, rounded: read K-last (MN-major) in place
        ours_d = lambda: ops.mm(a, bp, b_direct=True)
    elif kind == "nt":
        bp = ops.pack_tf32(b)
        ours_d = lambda: ops.mm_nt(a, bp, b_direct=True)
    else:
        ours_d = ours
    err = max((ours() - lib()).abs().max().item(), (ours_d() - lib()).abs().max().item())
    t_o, t_l = timeit(ours), timeit(lib)
    g_o, g_l, g_d = graph_time(ours), graph_time(lib), graph_time(ours_d)
    fl = 2.0 * M * N * K
    print(f"{name} M={M:5d} N={N:5d} K={K:5d}: ours {t_o:7.1f} us ({fl/t_o/1e6:6.1f} TF/s) cuBLAS {t_l:7.1f} us ({fl/t_l/1e6:6.1f})"
          f" | in a graph (warm L2): ours {g_o:6.1f} us, weights direct {g_d:6.1f} us, cuBLAS {g_l:6.1f} us | max diff {err:.2e}", flush=True)
