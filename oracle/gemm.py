"""CPU oracle of drm_gemm_tf32 (TEST INFRASTRUCTURE ONLY: imported by tests/, never by the product).

The reference's linear layers (torch.nn.Linear / nn.GRUCell inside /root/reference/SequenceModel.py:13-24, the MLPs of
DreamerUtils.py, differentiated by autograd behind WorldModel.py:193 and Agent.py:141-151) run on tensor cores in TF32
(train_car_racer.py:13 sets torch.backends.cuda.matmul.allow_tf32): operands rounded to 10 mantissa bits, products summed in fp32.
This restates that arithmetic with numpy: round-to-nearest (ties away from zero, PTX cvt.rna.tf32.f32) operands, float64 sums --
the GPU's fp32 accumulation order is unspecified, so parity is 'within fp32 summation error of this', with the bound in the test.
"""
import numpy as np


def tf32_round(x: np.ndarray) -> np.ndarray:
    """cvt.rna.tf32.f32: keep 10 mantissa bits, round to nearest, ties away from zero (NaN / Inf pass through)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    bits = x.view(np.uint32)
    finite = (bits & np.uint32(0x7F800000)) != np.uint32(0x7F800000)
    r = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000))
    return np.where(finite, r, bits).astype(np.uint32).view(np.float32)


def tf32_trunc(x: np.ndarray) -> np.ndarray:
    """what the tensor core does to an fp32 operand it reads directly: the 13 low mantissa bits are ignored"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def gemm_tf32(a: np.ndarray, b: np.ndarray, bias=None, c=None, a_mode="rna", b_mode="rna") -> np.ndarray:
    """(c +) tf32(a) [M, K] @ tf32(b) [N, K]^T (+ bias [N]) with float64 accumulation; returns float64.
    *_mode: "rna" (operand went through the pack kernel) or "rz" (read in place: truncated by the tensor core)."""
    cv = {"rna": tf32_round, "rz": tf32_trunc}
    out = cv[a_mode](a).astype(np.float64) @ cv[b_mode](b).astype(np.float64).T
    if bias is not None:
        out = out + np.asarray(bias, dtype=np.float64)[None, :]
    if c is not None:
        out = out + np.asarray(c, dtype=np.float64)
    return out


def abs_bound(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """sum_k |a_mk| |b_nk|: the scale fp32 summation error is proportional to"""
    return np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64).T
