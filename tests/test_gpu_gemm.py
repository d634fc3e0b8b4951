"""drm_gemm_tf32 (csrc/gemm_tf32.cu) through the C-ABI against the numpy TF32 oracle (oracle/gemm.py).

Bound: both sides multiply identically rounded TF32 operands, so the only difference is the fp32 accumulation (order
unspecified on the tensor core, fixed-order split-K reduction) against the oracle's float64 sums:
|err| <= 4e-6 * sum_k |a||b| (a few fp32 ulps of the absolute-value product) -- 250x tighter than one TF32 rounding step."""
import numpy as np
import pytest
import torch

from oracle import gemm as G

pytestmark = pytest.mark.gpu

REL = 4e-6


def _check(got, a, b, bias=None, c=None):
    ref = G.gemm_tf32(a, b, bias, c)
    bound = REL * (G.abs_bound(a, b) + (np.abs(c) if c is not None else 0) + (np.abs(bias)[None] if bias is not None else 0)) + 1e-30
    err = np.abs(got.astype(np.float64) - ref)
    assert np.all(err <= bound), f"max err/bound {np.max(err / bound):.3f} at {np.unravel_index(np.argmax(err / bound), err.shape)}"


SHAPES = [
    (1024, 1800, 1027),   # GRU input pre-activations of a 16 x 64 world-model batch (K % 4 != 0: packed pitch)
    (1024, 256, 1624),    # posterior MLP layer 1
    (16, 600, 1800),      # one BPTT step: dgh W_hh (swapped, split-K)
    (50, 1024, 256),      # one actor-backward step
    (1800, 1027, 1024),   # weight-gradient shape
    (50, 3, 1800),        # d(action) of one actor-backward step: few rows AND few columns
    (50, 600, 1800), (64, 1024, 2048), (24, 6, 256),
    (1, 1, 1), (3, 5, 7), (129, 17, 33), (64, 65, 40), (65, 64, 40), (300, 3, 256), (2, 1030, 96), (257, 255, 8),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_nt(M, N, K):
    from dreamer_b200 import ops
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    a = rng.standard_normal((M, K)).astype(np.float32)
    b = rng.standard_normal((N, K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    dev = torch.device("cuda")
    ta, tb, tbias = (torch.from_numpy(x).to(dev) for x in (a, b, bias))
    _check(ops.mm_nt(ta, tb).cpu().numpy(), a, b)
    _check(ops.mm_nt(ta, tb, tbias).cpu().numpy(), a, b, bias)
    # K-last operands (transposed views are read in place by the pack kernel)
    a_kl = ta.t().contiguous().t()
    b_kl = tb.t().contiguous().t()
    assert a_kl.stride(0) == 1 or M == 1 or K == 1
    _check(ops.mm_nt(a_kl, tb).cpu().numpy(), a, b)
    _check(ops.mm_nt(ta, b_kl, tbias).cpu().numpy(), a, b, bias)
    # accumulate into a strided output
    c = rng.standard_normal((M, N)).astype(np.float32)
    big = torch.zeros(M, N + 8, device=dev)
    out = big[:, 4:4 + N]
    out.copy_(torch.from_numpy(c))
    ops.mm_nt(a_kl, b_kl, out=out, accumulate=True)
    _check(out.cpu().numpy(), a, b, None, c)
    assert float(big[:, :4].abs().max()) == 0 and float(big[:, 4 + N:].abs().max()) == 0     # nothing written outside the view


DIRECT_SHAPES = [(1024, 1624, 256), (16, 600, 1800), (16, 1024, 1800), (50, 600, 256), (256, 1624, 1024), (1800, 600, 1024), (40, 24, 64),
                 (1024, 600, 1800), (200, 1000, 36)]


@pytest.mark.parametrize("M,N,K", DIRECT_SHAPES)
def test_gemm_direct_operands(M, N, K):
    """pre-rounded, aligned operands read in place by TMA in both orientations (K-first: 128-byte swizzle; K-last: MN-major
    descriptors on SWIZZLE_128B_ATOM_32B boxes), on either side of the MMA (swapped skinny problems put B on the M side)"""
    from dreamer_b200 import ops
    rng = np.random.default_rng(M + N * 5 + K * 11)
    a = G.tf32_round(rng.standard_normal((M, K)).astype(np.float32))
    b = G.tf32_round(rng.standard_normal((N, K)).astype(np.float32))
    dev = torch.device("cuda")
    ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    a_kl, b_kl = ta.t().contiguous().t(), tb.t().contiguous().t()
    ref = ops.mm_nt(ta, tb)
    _check(ref.cpu().numpy(), a, b)
    for x, ad in ((ta, True), (a_kl, True), (ta, False)):
        for y, bd in ((tb, True), (b_kl, True), (tb, False)):
            got = ops.mm_nt(x, y, a_direct=ad, b_direct=bd)
            assert torch.equal(got, ref), f"a {'K-last' if x is a_kl else 'K-first'} direct={ad}, b {'K-last' if y is b_kl else 'K-first'} direct={bd}: max diff {(got - ref).abs().max().item()}"


def test_gemm_packed_and_deterministic():
    from dreamer_b200 import ops
    rng = np.random.default_rng(5)
    M, N, K = 16, 600, 1801
    a = rng.standard_normal((M, K)).astype(np.float32)
    w = rng.standard_normal((K, N)).astype(np.float32)     # dX = dY W: W given [K, N]
    dev = torch.device("cuda")
    ta, tw = torch.from_numpy(a).to(dev), torch.from_numpy(w).to(dev)
    wp = ops.pack_tf32(tw.t())
    assert wp.shape == (N, K) and wp.stride(0) % 4 == 0
    assert np.array_equal(wp.cpu().numpy(), G.tf32_round(w.T))
    r1 = ops.mm_nt(ta, wp, b_direct=True)                   # a: rounded inside the kernel (skinny); wp: pre-rounded, read in place
    r2 = ops.mm(ta, tw)
    r3 = ops.mm_nt(ta, wp, b_direct=True)
    _check(r1.cpu().numpy(), a, w.T.copy())
    assert torch.equal(r1, r2) and torch.equal(r1, r3)      # same arithmetic whichever way the operand arrives; run-to-run identical
    # split-K tickets are left zero: a different shape on the same workspace right after
    b2 = rng.standard_normal((300, K)).astype(np.float32)
    tb2 = torch.from_numpy(b2).to(dev)
    _check(ops.mm_nt(ta, tb2).cpu().numpy(), a, b2)
    for _ in range(3):
        assert torch.equal(ops.mm_nt(ta, wp, b_direct=True), r1)


def test_gemm_in_graph_and_side_stream():
    """captured in a CUDA graph and replayed; a second stream gets its own workspace"""
    from dreamer_b200 import ops
    dev = torch.device("cuda")
    g0 = torch.Generator(device=dev).manual_seed(1)
    a = torch.randn(16, 1024, device=dev, generator=g0)
    w = torch.randn(1024, 256, device=dev, generator=g0)
    x = torch.randn(1024, 512, device=dev, generator=g0)
    dy = torch.randn(1024, 256, device=dev, generator=g0)
    ref1, ref2 = ops.mm(a, w), ops.mm_nt(dy.t(), x.t())
    out1, out2 = torch.empty_like(ref1), torch.empty_like(ref2)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ops.mm(a, w, out=out1); ops.mm_nt(dy.t(), x.t(), out=out2)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        out1.zero_(); out2.zero_()
        with torch.cuda.graph(graph, stream=s):
            ops.mm(a, w, out=out1)
            ops.mm_nt(dy.t(), x.t(), out=out2)
    for _ in range(3):
        out1.zero_(); out2.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out1, ref1) and torch.equal(out2, ref2)
    _check(ref2.cpu().numpy(), dy.t().cpu().numpy(), x.t().cpu().numpy())


def test_gemm_errors():
    from dreamer_b200 import ops
    dev = torch.device("cuda")
    with pytest.raises(RuntimeError):
        ops.mm_nt(torch.zeros(4, 8, device=dev), torch.zeros(4, 9, device=dev))
    with pytest.raises(RuntimeError):
        ops.mm_nt(torch.zeros(4, 8, device=dev), torch.zeros(4, 8, device=dev), out=torch.zeros(4, 5, device=dev))
    with pytest.raises(RuntimeError):
        ops.mm_nt(torch.zeros(4, 8), torch.zeros(4, 8))
