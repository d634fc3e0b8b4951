"""GPU parity of the HBM-bound kernels and the raw TMA/tcgen05 main loop, through the C-ABI."""
import numpy as np
import pytest
import torch

from oracle import rssm as O
from oracle.replay import ReplayOracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from dreamer_b200 import ops as _ops
    return _ops


DEV = "cuda"


@pytest.mark.parametrize("n_rows,scale", [(1, 1.0), (33, 3.0), (32 * 1024, 2.0), (5000, 8.0)])
def test_categorical32_bit_exact_indices(ops, n_rows, scale):
    g = torch.Generator().manual_seed(n_rows)
    logits = torch.randn(n_rows, 32, generator=g) * scale
    u = torch.rand(n_rows, generator=g)
    u = O.interior_uniforms(O.unimix_probs(logits), u, 0.0, 1e-5)   # keep draws off CDF edges by 1e-5
    z, idx, p = O.categorical_st(logits, u)
    out = ops.categorical32(logits.to(DEV), u.to(DEV), want_probs=True, want_bf16=True)
    assert torch.equal(out["idx"].cpu().long(), idx)                 # bit-exact
    assert torch.allclose(out["probs"].cpu(), p, atol=1e-6)          # fp32 tolerance 1e-4 relative is far looser
    assert torch.allclose(out["z"].cpu(), z, atol=2e-7)
    assert torch.equal(out["z_bf16"].cpu().float(), torch.nn.functional.one_hot(idx, 32).float())


def test_categorical32_empty_and_extremes(ops):
    out = ops.categorical32(torch.zeros(0, 32, device=DEV), torch.zeros(0, device=DEV))
    assert out["idx"].numel() == 0
    logits = torch.zeros(4, 32)
    logits[1, 7] = 80.0
    logits[2, 31] = -80.0
    u = torch.tensor([0.0, 0.5, 0.999999, 0.99999994])
    _, idx, _ = O.categorical_st(logits, u)
    got = ops.categorical32(logits.to(DEV), u.to(DEV))["idx"].cpu().long()
    assert torch.equal(got, idx)


def test_categorical32_kl(ops):
    g = torch.Generator().manual_seed(0)
    post = torch.randn(7, 5, 32, 32, generator=g) * 2
    prior = torch.randn(7, 5, 32, 32, generator=g)
    ref = O.categorical_kl_terms(post, prior)
    got = ops.categorical32_kl(post.to(DEV), prior.to(DEV)).cpu()
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4)


def _filled_ring(cap, L, fill, seed=7):
    ro = ReplayOracle(cap, L, 3, (64, 64))
    rng = np.random.Generator(np.random.PCG64(seed))
    for i in range(fill):
        ro.add(rng.integers(0, 256, size=(3, 64, 64)).astype(np.uint8), rng.uniform(-1, 1, 3).astype(np.float32),
               float(rng.standard_normal() * 5), float(i % 9 != 8))
    return ro


@pytest.mark.parametrize("fill", [20, 37 + 11])
def test_replay_gather_bit_exact(ops, fill):
    cap, L, B = 37, 8, 16
    ro = _filled_ring(cap, L, fill)
    starts = np.random.RandomState(5).randint(0, cap, size=B)          # includes windows that wrap modulo capacity
    o, a, r, c, _ = ro.gather(starts)
    ring = [torch.from_numpy(x).to(DEV) for x in (ro.obs, ro.act, ro.rew, ro.con)]
    go, ga, gr, gc = ops.replay_gather(*ring, torch.from_numpy(starts), L)
    assert torch.equal(go.cpu(), torch.from_numpy(o))                  # u8 -> f32 is exact
    assert torch.equal(ga.cpu(), torch.from_numpy(a)) and torch.equal(gr.cpu(), torch.from_numpy(r)) and torch.equal(gc.cpu(), torch.from_numpy(c))
    gn = ops.replay_gather(*ring, torch.from_numpy(starts), L, normalise=True)[0]
    assert torch.allclose(gn.cpu(), torch.from_numpy(o) / 255.0 - 0.5, atol=1e-7)
    # empty batch
    e = ops.replay_gather(*ring, torch.zeros(0, dtype=torch.int64), L)
    assert e[0].shape[0] == 0


def test_replay_insert_matches_reference_ring(ops):
    cap, L = 16, 4
    ro = ReplayOracle(cap, L, 3, (64, 64))
    rng = np.random.Generator(np.random.PCG64(1))
    ring = [torch.zeros(cap, 3, 64, 64, dtype=torch.uint8, device=DEV), torch.zeros(cap, 3, device=DEV),
            torch.zeros(cap, 1, device=DEV), torch.zeros(cap, 1, device=DEV)]
    nxt = 0
    for chunk in (5, 7, 9):                                            # third chunk wraps
        obs = rng.integers(0, 256, size=(chunk, 3, 64, 64)).astype(np.uint8)
        act = rng.uniform(-1, 1, (chunk, 3)).astype(np.float32)
        rew = (rng.standard_normal(chunk) * 5).astype(np.float32)
        con = (rng.random(chunk) > 0.2).astype(np.float32)
        for i in range(chunk):
            ro.add(obs[i], act[i], rew[i], con[i])
        ops.replay_insert(*ring, torch.from_numpy(obs).to(DEV), torch.from_numpy(act).to(DEV), torch.from_numpy(rew).to(DEV),
                          torch.from_numpy(con).to(DEV), nxt)
        nxt = (nxt + chunk) % cap
    assert nxt == ro.next_idx
    assert torch.equal(ring[0].cpu(), torch.from_numpy(ro.obs))
    assert torch.equal(ring[1].cpu(), torch.from_numpy(ro.act))
    assert torch.allclose(ring[2].cpu(), torch.from_numpy(ro.rew), atol=1e-6)
    assert torch.equal(ring[3].cpu(), torch.from_numpy(ro.con))


def test_lambda_return(ops):
    g = torch.Generator().manual_seed(3)
    B, H = 257, 15
    rew = torch.randn(B, H, 1, generator=g)
    cont = (torch.rand(B, H, 1, generator=g) > 0.1).float() * 0.97
    val = torch.randn(B, H + 1, 1, generator=g) * 3
    ref = O.lambda_returns(rew, cont, val, 0.99, 0.95)
    got = ops.lambda_return(rew.to(DEV), cont.to(DEV), val.to(DEV), 0.99, 0.95).cpu()
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5)              # fp32: 1e-4
    one = ops.lambda_return(rew[:, :1].to(DEV), cont[:, :1].to(DEV), val[:, :2].to(DEV), 0.99, 0.95).cpu()
    assert torch.allclose(one, O.lambda_returns(rew[:, :1], cont[:, :1], val[:, :2], 0.99, 0.95), atol=1e-5)


def test_twohot_ce_and_bucket_value(ops):
    g = torch.Generator().manual_seed(4)
    b = torch.linspace(-20.0, 20.0, 255)
    logits = torch.randn(1000, 255, generator=g) * 2
    v = torch.cat([torch.randn(989, 1, generator=g) * 6,
                   torch.tensor([[-25.0], [-20.0], [0.0], [7.45e-8], [19.9999], [20.0], [31.0], [float(b[10])], [float(b[200])], [1e-3], [-1e-3]])])
    ref = O.twohot_ce(logits, v, b)
    got = ops.twohot_ce(logits.to(DEV), v.to(DEV), b.to(DEV)).cpu()
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4)
    ref2 = O.twohot_ce(logits, O.symlog(v * 50), b)
    got2 = ops.twohot_ce(logits.to(DEV), (v * 50).to(DEV), b.to(DEV), apply_symlog=True).cpu()
    assert torch.allclose(got2, ref2, rtol=1e-4, atol=1e-4)
    refv = O.symexp((torch.softmax(logits, -1) * b).sum(-1, keepdim=True))
    gotv = ops.bucket_value(logits.to(DEV), b.to(DEV)).cpu()
    assert torch.allclose(gotv, refv, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 256), (1, 8, 16), (300, 700, 1027), (1024, 1800, 1627), (257, 255, 200)])
def test_tcgen05_gemm_mainloop(ops, M, N, K):
    """The TMA + tcgen05 main loop against a bf16-rounded fp32 reference (same rounding of the inputs)."""
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    Wt = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g)
    ref = A.bfloat16().float() @ Wt.bfloat16().float().T + bias
    got = ops.test_gemm(A.to(DEV), Wt.to(DEV), bias.to(DEV)).cpu()
    assert torch.allclose(got, ref, rtol=1e-4, atol=2e-4), (got - ref).abs().max()


def test_nan_targets_propagate_like_torch():
    """torch.clamp / sign * log keep a NaN (DreamerUtils.py:29-50): the two-hot CE of a NaN target is NaN, so the reference's
    NaN-loss skip (and the fused optimiser's device-side skip) sees it; finite rows are untouched."""
    from dreamer_b200 import ops
    r = torch.randn(4, 5, 1, device=DEV)
    r[0, 0, 0] = float("nan")
    R = ops.lambda_return(r, torch.ones_like(r), torch.randn(4, 6, 1, device=DEV), 0.99, 0.95)
    assert torch.isnan(R[0, 0, 0]) and torch.isfinite(R[1:]).all()
    b = torch.linspace(-20, 20, 255, device=DEV)
    lg = torch.randn(4, 5, 255, device=DEV)
    for sym in (False, True):
        ce = ops.twohot_ce(lg, R, b, apply_symlog=sym)
        assert torch.isnan(ce[0, 0, 0]) and torch.isfinite(ce[1:]).all() and torch.isfinite(ce[0, 1:]).all()


def test_straight_through_forward_and_backward_match_autograd():
    """ops.straight_through (teacher-forced ST latent, one kernel each way) against the torch expression the reference
    differentiates (DynamicsPredictors.py:33-39): value bit-exact, gradient to fp32 rounding; ragged row count."""
    from dreamer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    for n in (1, 37, 4096 + 5):
        lg = (torch.randn(n, 32, device=DEV, generator=g) * 3).requires_grad_(True)
        idx = torch.randint(0, 32, (n,), device=DEV, generator=g, dtype=torch.uint8)
        w = torch.randn(n, 32, device=DEV, generator=g)
        z = ops.straight_through(lg, idx)
        (z * w).sum().backward()
        g_kernel = lg.grad.clone()
        lg.grad = None
        p = 0.99 * torch.softmax(lg, -1) + 0.01 / 32
        z_ref = (torch.nn.functional.one_hot(idx.long(), 32).float() + p) - p.detach()
        (z_ref * w).sum().backward()
        assert torch.equal(z.detach() != 0, z_ref.detach() != 0) and torch.allclose(z.detach(), z_ref.detach(), atol=1e-7)
        assert torch.allclose(g_kernel, lg.grad, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("n", [1, 2, 3, 20, 1500, 15360, 122880])
def test_percentile_pair_matches_sort(n):
    """drm_percentile_pair (radix select) against the sort-and-interpolate definition Agent.update_S spells out (Agent.py:78-88):
    bit-exact order statistics, identical interpolation arithmetic."""
    import torch
    from dreamer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(n)
    for kind in range(4):
        x = torch.randn(n, device="cuda", generator=g) * (50.0 if kind == 1 else 1.0)
        if kind == 2:
            x = torch.round(x * 2) / 2                     # many ties, +-0
        if kind == 3:
            x = x.abs() * 1e-3 + 7.0                        # narrow range: the leading digits all collide
        s, _ = torch.sort(x)

        def q(p):
            pos = p * (n - 1)
            lo = int(pos)
            hi = min(lo + 1, n - 1)
            return s[lo] + (s[hi] - s[lo]) * torch.tensor(pos - lo, dtype=torch.float32, device="cuda")

        out = ops.percentile_pair(x, 0.05, 0.95)
        assert float(out[2]) == 1.0
        assert float(out[0]) == float(q(0.05)) and float(out[1]) == float(q(0.95)), (n, kind, out.tolist(), float(q(0.05)), float(q(0.95)))
    x[n // 2] = float("inf")
    assert float(ops.percentile_pair(x, 0.05, 0.95)[2]) == 0.0
    x[n // 2] = float("nan")
    assert float(ops.percentile_pair(x, 0.05, 0.95)[2]) == 0.0
