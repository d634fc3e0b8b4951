"""CPU checks of the TF32 GEMM oracle (oracle/gemm.py) against known answers."""
import numpy as np

from oracle import gemm as G


def test_tf32_round_known_answers():
    one = np.float32(1.0)
    ulp10 = np.float32(2.0 ** -10)
    x = np.array([1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -11 - 2.0 ** -23, 1.0 + 2.0 ** -11 + 2.0 ** -23, -(1.0 + 2.0 ** -11),
                  3.14159265, 0.0, -0.0, 65504.0, 1e-30], dtype=np.float32)
    r = G.tf32_round(x)
    assert r[0] == one
    assert r[1] == one + ulp10            # tie: away from zero
    assert r[2] == one                    # just below the tie
    assert r[3] == one + ulp10
    assert r[4] == -(one + ulp10)
    assert (r.view(np.uint32) & np.uint32(0x1FFF)).max() == 0          # 13 low mantissa bits cleared
    assert np.all(np.abs(r - x) <= np.abs(x) * 2.0 ** -11 + 1e-45)
    sp = np.array([np.inf, -np.inf, np.nan], dtype=np.float32)
    rs = G.tf32_round(sp)
    assert np.isinf(rs[0]) and np.isinf(rs[1]) and np.isnan(rs[2])


def test_gemm_tf32_close_to_exact():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((7, 50)).astype(np.float32)
    b = rng.standard_normal((9, 50)).astype(np.float32)
    bias = rng.standard_normal(9).astype(np.float32)
    c = rng.standard_normal((7, 9)).astype(np.float32)
    exact = a.astype(np.float64) @ b.astype(np.float64).T + bias + c
    got = G.gemm_tf32(a, b, bias, c)
    assert np.all(np.abs(got - exact) <= 2.0 ** -10 * G.abs_bound(a, b))    # two operand roundings of 2^-11 each
    # operands that are exactly representable in TF32 give the exact product
    ai = rng.integers(-8, 8, (5, 33)).astype(np.float32)
    bi = rng.integers(-8, 8, (4, 33)).astype(np.float32)
    assert np.array_equal(G.gemm_tf32(ai, bi), ai.astype(np.float64) @ bi.astype(np.float64).T)
