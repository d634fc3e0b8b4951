"""Tensor-level wrappers over the C-ABI (one function per entry point of include/dreamer_b200.h).

All tensors must live on the CUDA device; outputs are allocated with torch (device memory
plumbing only) and filled by the library on the current stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib as L


# --------------------------------------------------------------------------------------------
# (2) categorical head
# --------------------------------------------------------------------------------------------
def categorical32(logits: torch.Tensor, uniforms: torch.Tensor, want_probs: bool = False, want_bf16: bool = False):
    """logits (..., 32) fp32, uniforms (...) -> dict(z=(...,32) straight-through, idx=(...) uint8[, probs, z_bf16]).

    Replaces DynamicsPredictors.py:33-39 / VariationalAutoEncoder.py:88-98.
    """
    L.require_cuda(logits, "logits")
    if logits.shape[-1] != 32:
        raise RuntimeError("dreamer_b200.categorical32: the class dimension must be 32")
    lg = L.f32c(logits)
    u = L.f32c(uniforms)
    n = lg.numel() // 32
    if u.numel() != n:
        raise RuntimeError("dreamer_b200.categorical32: one uniform per categorical row is required")
    z = torch.empty_like(lg)
    idx = torch.empty(lg.shape[:-1], dtype=torch.uint8, device=lg.device)
    probs = torch.empty_like(lg) if want_probs else None
    zb = torch.empty(lg.shape, dtype=torch.bfloat16, device=lg.device) if want_bf16 else None
    L.check(L.load().drm_categorical32_fwd(L.ptr(lg), L.ptr(u), L.ptr(idx), L.ptr(z), L.ptr(probs), L.ptr(zb), n, L.stream()),
            "categorical32_fwd")
    out = dict(z=z, idx=idx)
    if want_probs:
        out["probs"] = probs
    if want_bf16:
        out["z_bf16"] = zb
    return out


class _StraightThrough(torch.autograd.Function):
    """z = onehot(idx) + p - stopgrad(p) on GIVEN classes, one kernel forward and one backward
    (drm_categorical32_st / drm_categorical32_bwd) instead of ~8 + ~12 elementwise launches per call."""

    @staticmethod
    def forward(ctx, logits, idx):
        lg = L.f32c(logits)
        ix = idx.to(torch.uint8).contiguous()
        z = torch.empty_like(lg)
        L.check(L.load().drm_categorical32_st(L.ptr(lg), L.ptr(ix), L.ptr(z), None, lg.numel() // 32, L.stream()), "categorical32_st")
        ctx.save_for_backward(lg)
        return z

    @staticmethod
    def backward(ctx, dz):
        (lg,) = ctx.saved_tensors
        g = L.f32c(dz)
        out = torch.empty_like(lg)
        L.check(L.load().drm_categorical32_bwd(L.ptr(lg), L.ptr(g), None, None, L.ptr(out), lg.numel() // 32, L.stream()), "categorical32_bwd")
        return out, None


def categorical32_bwd(logits, dz, dz2=None, dl_add=None, out=None):
    """dlogits = ST-backward(logits, dz [+ dz2]) [+ dl_add]; all (..., 32) fp32 contiguous (no copies are made)."""
    n = logits.numel() // 32
    if out is None:
        out = torch.empty_like(logits)
    for t in (logits, dz, dz2, dl_add, out):
        if t is not None and not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32 and t.numel() == n * 32):
            raise RuntimeError("dreamer_b200.categorical32_bwd: fp32 contiguous CUDA tensors of one shape are required")
    L.check(L.load().drm_categorical32_bwd(L.ptr(logits), L.ptr(dz), L.ptr(dz2), L.ptr(dl_add), L.ptr(out), n, L.stream()), "categorical32_bwd")
    return out


def ln_silu_bwd(dy, a, gamma, beta, eps: float = 1e-5, want_dln: bool = False, out=None, want_dlnx: bool = False):
    """Backward of SiLU(LayerNorm(a) * gamma + beta): dy, a (..., n) -> da (..., n) [, dln] [, dlnx = dln * xhat with want_dlnx]."""
    n = a.shape[-1]
    for t in (dy, a, gamma, beta):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
            raise RuntimeError("dreamer_b200.ln_silu_bwd: fp32 contiguous CUDA tensors are required")
    if dy.shape != a.shape or gamma.numel() != n or beta.numel() != n:
        raise RuntimeError("dreamer_b200.ln_silu_bwd: shape mismatch")
    da = torch.empty_like(a) if out is None else out
    dln = torch.empty_like(a) if want_dln else None
    if want_dlnx:
        dlnx = torch.empty_like(a)
        L.check(L.load().drm_ln_silu_bwd_affine(L.ptr(dy), L.ptr(a), L.ptr(gamma), L.ptr(beta), L.ptr(da), L.ptr(dln), L.ptr(dlnx),
                                                a.numel() // n, n, eps, L.stream()), "ln_silu_bwd_affine")
        return da, dln, dlnx
    L.check(L.load().drm_ln_silu_bwd(L.ptr(dy), L.ptr(a), L.ptr(gamma), L.ptr(beta), L.ptr(da), L.ptr(dln), a.numel() // n, n, eps, L.stream()),
            "ln_silu_bwd")
    return (da, dln) if want_dln else da


def convt_image_fwd(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """tanh(conv_transpose2d(x, weight, bias, stride=2, padding=1)) for the decoder's image layer: x (N, C_in, H, W) bf16 in
    channels-last memory, weight (C_in, C_out <= 3, 4, 4) fp32, bias (C_out) -> (N, C_out, 2H, 2W) fp32."""
    L.require_cuda(x, "x")
    if x.dtype != torch.bfloat16 or x.dim() != 4 or not x.is_contiguous(memory_format=torch.channels_last):
        raise RuntimeError("dreamer_b200.convt_image_fwd: x must be a bf16 channels-last (N, C, H, W) tensor")
    N, Ci, H, W = x.shape
    Co = weight.shape[1]
    if tuple(weight.shape) != (Ci, Co, 4, 4) or bias.numel() != Co:
        raise RuntimeError("dreamer_b200.convt_image_fwd: weight must be (C_in, C_out, 4, 4), bias (C_out)")
    w, b = L.f32c(weight), L.f32c(bias)
    out = torch.empty((N, Co, 2 * H, 2 * W), dtype=torch.float32, device=x.device)
    L.check(L.load().drm_convt_image_fwd(L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(out), N, H, W, Ci, Co, L.stream()), "convt_image_fwd")
    return out


def colsum(x: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """out [n] (+)= x [rows, n].sum(0) (x may be a column slice of a wider matrix; fp32, or bf16 with even n / pitch); deterministic."""
    L.require_cuda(x, "x")
    bf16 = x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1 and x.shape[1] % 2 == 0 and x.stride(0) % 2 == 0 \
        and x.data_ptr() % 4 == 0 and x.shape[0] >= 1
    if not bf16 and (x.dtype != torch.float32 or x.dim() != 2 or (x.stride(1) != 1 and x.shape[1] != 1)):
        x = L.f32c(x.reshape(-1, x.shape[-1]))
    rows, n = x.shape
    if out is None:
        if accumulate:
            raise RuntimeError("dreamer_b200.colsum: accumulate needs `out`")
        out = torch.empty(n, dtype=torch.float32, device=x.device)
    elif out.dtype != torch.float32 or out.numel() != n or not out.is_contiguous():
        raise RuntimeError("dreamer_b200.colsum: `out` must be a contiguous fp32 tensor of n elements")
    lib = L.load()
    if bf16:
        nscr = lib.drm_colsum_bf16_scratch_bytes(rows, n)
        scratch = torch.empty(nscr, dtype=torch.uint8, device=x.device)
        L.check(lib.drm_colsum_bf16(L.ptr(x), rows, n, x.stride(0) if rows > 1 else n, L.ptr(out), 1 if accumulate else 0, L.ptr(scratch),
                                    L.stream()), "colsum_bf16")
        return out
    nscr = lib.drm_colsum_scratch_bytes(rows, n)
    scratch = torch.empty(nscr, dtype=torch.uint8, device=x.device) if nscr else None
    L.check(lib.drm_colsum(L.ptr(x), rows, n, x.stride(0) if rows > 1 else n, L.ptr(out), 1 if accumulate else 0, L.ptr(scratch), L.stream()),
            "colsum")
    return out


def gru_bwd(dh, gi, gh, h_prev, dgi, dgh, dh_prev=None, accumulate: bool = False, dh_add=None):
    """Elementwise backward of one GRUCell step (drm_gru_bwd); writes dgi, dgh (rows, 3D) and dh_prev (=|+=) dh * u.
    ``dh_add`` (optional, same shape as dh) is added to dh (drm_gru_bwd_add)."""
    rows, Dh = dh.shape
    for t in (dh, gi, gh, h_prev, dgi, dgh, dh_prev, dh_add):
        if t is not None and not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
            raise RuntimeError("dreamer_b200.gru_bwd: fp32 contiguous CUDA tensors are required")
    if gi.shape != (rows, 3 * Dh) or gh.shape != gi.shape or dgi.shape != gi.shape or dgh.shape != gi.shape:
        raise RuntimeError("dreamer_b200.gru_bwd: shape mismatch")
    if dh_add is not None:
        L.check(L.load().drm_gru_bwd_add(L.ptr(dh), L.ptr(dh_add), L.ptr(gi), L.ptr(gh), L.ptr(h_prev), L.ptr(dgi), L.ptr(dgh), L.ptr(dh_prev),
                                         1 if accumulate else 0, rows, Dh, L.stream()), "gru_bwd_add")
        return
    L.check(L.load().drm_gru_bwd(L.ptr(dh), L.ptr(gi), L.ptr(gh), L.ptr(h_prev), L.ptr(dgi), L.ptr(dgh), L.ptr(dh_prev),
                                 1 if accumulate else 0, rows, Dh, L.stream()), "gru_bwd")


def straight_through(logits: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """Differentiable straight-through latent on given classes: logits (..., 32), idx (...) -> z (..., 32).
    DynamicsPredictors.py:33-39 / VariationalAutoEncoder.py:88-98 with the draw replaced by `idx` (teacher forcing)."""
    L.require_cuda(logits, "logits")
    if logits.shape[-1] != 32 or tuple(idx.shape) != tuple(logits.shape[:-1]):
        raise RuntimeError("dreamer_b200.straight_through: logits (..., 32) and idx (...) are required")
    return _StraightThrough.apply(logits, idx)


def categorical32_kl(post_logits: torch.Tensor, prior_logits: torch.Tensor) -> torch.Tensor:
    """(..., R, 32) x2 -> (...) sum over R of KL(Cat(post)||Cat(prior)).  WorldModel.py:175-181."""
    L.require_cuda(post_logits, "post_logits")
    a, b = L.f32c(post_logits), L.f32c(prior_logits)
    if a.shape != b.shape or a.shape[-1] != 32:
        raise RuntimeError("dreamer_b200.categorical32_kl: shapes must match and end in 32")
    R = a.shape[-2]
    out = torch.empty(a.shape[:-2], dtype=torch.float32, device=a.device)
    L.check(L.load().drm_categorical32_kl(L.ptr(a), L.ptr(b), L.ptr(out), out.numel(), R, L.stream()), "categorical32_kl")
    return out


# --------------------------------------------------------------------------------------------
# (4) replay ring
# --------------------------------------------------------------------------------------------
def replay_gather(ring_obs, ring_act, ring_rew, ring_con, starts, L_seq: int, normalise: bool = False):
    """Buffer.py:49-61.  ring_obs (cap, 3, H, W) uint8 on device; starts (B,) int64 -> obs (B, L, 3, H, W) fp32, ..."""
    L.require_cuda(ring_obs, "ring_obs")
    cap = ring_obs.shape[0]
    frame = ring_obs[0].numel()
    A = ring_act.shape[1]
    B = starts.numel()
    starts = starts.to(device=ring_obs.device, dtype=torch.int64).contiguous()
    obs = torch.empty((B, L_seq) + tuple(ring_obs.shape[1:]), dtype=torch.float32, device=ring_obs.device)
    act = torch.empty((B, L_seq, A), dtype=torch.float32, device=ring_obs.device)
    rew = torch.empty((B, L_seq, 1), dtype=torch.float32, device=ring_obs.device)
    con = torch.empty((B, L_seq, 1), dtype=torch.float32, device=ring_obs.device)
    L.check(L.load().drm_replay_gather(L.ptr(ring_obs), L.ptr(ring_act), L.ptr(ring_rew), L.ptr(ring_con), L.ptr(starts),
                                       L.ptr(obs), L.ptr(act), L.ptr(rew), L.ptr(con), B, L_seq, cap, frame, A,
                                       1 if normalise else 0, L.stream()), "replay_gather")
    return obs, act, rew, con


def replay_insert(ring_obs, ring_act, ring_rew, ring_con, obs_u8, act, rew, con, next_idx: int):
    """Buffer.py:19-30 for n transitions already on the device."""
    L.require_cuda(ring_obs, "ring_obs")
    n = obs_u8.shape[0]
    L.check(L.load().drm_replay_insert(L.ptr(ring_obs), L.ptr(ring_act), L.ptr(ring_rew), L.ptr(ring_con),
                                       L.ptr(obs_u8.contiguous()), L.ptr(L.f32c(act)), L.ptr(L.f32c(rew)), L.ptr(L.f32c(con)),
                                       next_idx, n, ring_obs.shape[0], ring_obs[0].numel(), ring_act.shape[1], L.stream()),
            "replay_insert")


# --------------------------------------------------------------------------------------------
# returns / losses
# --------------------------------------------------------------------------------------------
def lambda_return(rew, cont, value, gamma: float, lam: float):
    """Agent.py:158-171.  rew, cont (B,H,1) or (B,H); value (B,H+1,1) or (B,H+1) -> same rank as rew."""
    L.require_cuda(rew, "rew")
    r, c, v = L.f32c(rew), L.f32c(cont), L.f32c(value)
    B, H = r.shape[0], r.shape[1]
    if v.shape[1] != H + 1:
        raise RuntimeError("dreamer_b200.lambda_return: value must have H + 1 steps")
    out = torch.empty_like(r)
    L.check(L.load().drm_lambda_return(L.ptr(r), L.ptr(c), L.ptr(v), L.ptr(out), B, H, gamma, lam, L.stream()), "lambda_return")
    return out


def twohot_ce(logits, value, buckets, apply_symlog: bool = False):
    """sum(twohot(value) * log_softmax(logits)): logits (..., NB), value (..., 1) -> (..., 1)."""
    L.require_cuda(logits, "logits")
    lg, v, b = L.f32c(logits), L.f32c(value), L.f32c(buckets)
    NB = lg.shape[-1]
    N = lg.numel() // NB
    out = torch.empty(lg.shape[:-1] + (1,), dtype=torch.float32, device=lg.device)
    L.check(L.load().drm_twohot_ce(L.ptr(lg), L.ptr(v), L.ptr(b), L.ptr(out), N, NB, 1 if apply_symlog else 0, L.stream()), "twohot_ce")
    return out


def tanh_normal_logp(a, mu, sigma, coef=None, want_logp: bool = True, want_grad: bool = False):
    """Agent.py:110-115: a, mu, sigma (..., A) -> logp (...) and / or (g_mu, g_sigma) (..., A) = coef (...) * d logp / d(mu, sigma)."""
    L.require_cuda(a, "a")
    av, m, sg = L.f32c(a), L.f32c(mu), L.f32c(sigma)
    A = av.shape[-1]
    rows = av.numel() // A
    cf = L.f32c(coef) if coef is not None else None
    if m.shape != av.shape or sg.shape != av.shape or (cf is not None and cf.numel() != rows):
        raise RuntimeError("dreamer_b200.tanh_normal_logp: shape mismatch")
    logp = torch.empty(av.shape[:-1], dtype=torch.float32, device=av.device) if want_logp else None
    gm = torch.empty_like(av) if want_grad else None
    gs = torch.empty_like(av) if want_grad else None
    L.check(L.load().drm_tanh_normal_logp(L.ptr(av), L.ptr(m), L.ptr(sg), L.ptr(cf), L.ptr(logp), L.ptr(gm), L.ptr(gs), rows, A, L.stream()),
            "tanh_normal_logp")
    if want_logp and want_grad:
        return logp, gm, gs
    return logp if want_logp else (gm, gs)


def percentile_pair(x: torch.Tensor, p_lo: float, p_hi: float) -> torch.Tensor:
    """(q(p_lo), q(p_hi), all-finite flag) of the flattened x as a 3-element device tensor (torch.quantile's linear rule, Agent.py:78-88)."""
    L.require_cuda(x, "x")
    xv = L.f32c(x).reshape(-1)
    out = torch.empty(3, dtype=torch.float32, device=xv.device)
    L.check(L.load().drm_percentile_pair(L.ptr(xv), xv.numel(), float(p_lo), float(p_hi), L.ptr(out), L.stream()), "percentile_pair")
    return out


def twohot_ce_bwd(logits, value, buckets, coef=None, scale_dev=None, scale: float = 1.0, apply_symlog: bool = False):
    """d/dlogits of sum(coef * twohot_ce(logits, value)) * scale * scale_dev: logits (..., NB), value (..., 1), coef (..., 1) or None,
    scale_dev a 0-d device tensor or None -> (..., NB)."""
    L.require_cuda(logits, "logits")
    lg, v, b = L.f32c(logits), L.f32c(value), L.f32c(buckets)
    NB = lg.shape[-1]
    N = lg.numel() // NB
    cf = L.f32c(coef) if coef is not None else None
    sd = L.f32c(scale_dev) if scale_dev is not None else None
    if v.numel() != N or (cf is not None and cf.numel() != N):
        raise RuntimeError("dreamer_b200.twohot_ce_bwd: one value (and coefficient) per row is required")
    out = torch.empty_like(lg)
    L.check(L.load().drm_twohot_ce_bwd(L.ptr(lg), L.ptr(v), L.ptr(b), L.ptr(cf), L.ptr(sd), float(scale), L.ptr(out), N, NB,
                                       1 if apply_symlog else 0, L.stream()), "twohot_ce_bwd")
    return out


def onehot32(idx: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """(..., ) uint8 classes < 32 -> (..., 32) fp32 one-hot (into `out` when given)."""
    L.require_cuda(idx, "idx")
    if idx.dtype != torch.uint8:
        raise RuntimeError("dreamer_b200.onehot32: idx must be uint8")
    ix = idx.contiguous()
    if out is None:
        out = torch.empty(tuple(ix.shape) + (32,), dtype=torch.float32, device=ix.device)
    elif out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != ix.numel() * 32:
        raise RuntimeError("dreamer_b200.onehot32: out must be a contiguous fp32 tensor with 32 values per class index")
    L.check(L.load().drm_onehot32(L.ptr(ix), L.ptr(out), ix.numel(), L.stream()), "onehot32")
    return out


def bucket_value(logits, buckets):
    """symexp(sum(softmax(logits) * buckets)): (..., NB) -> (..., 1)."""
    L.require_cuda(logits, "logits")
    lg, b = L.f32c(logits), L.f32c(buckets)
    NB = lg.shape[-1]
    out = torch.empty(lg.shape[:-1] + (1,), dtype=torch.float32, device=lg.device)
    L.check(L.load().drm_bucket_value(L.ptr(lg), L.ptr(b), L.ptr(out), lg.numel() // NB, NB, L.stream()), "bucket_value")
    return out


def test_gemm(A, W, bias=None):
    """out = bf16(A) @ bf16(W)^T + bias through the TMA/tcgen05 main loop (test hook)."""
    L.require_cuda(A, "A")
    a, w = L.f32c(A), L.f32c(W)
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    b = L.f32c(bias) if bias is not None else None
    L.check(L.load().drm_test_gemm(L.ptr(a), L.ptr(w), L.ptr(b), L.ptr(out), M, N, K, L.stream()), "test_gemm")
    return out


# --------------------------------------------------------------------------------------------
# contractions of the backward passes (drm_gemm_tf32)
# --------------------------------------------------------------------------------------------
GEMM_TRANS_A, GEMM_TRANS_B, GEMM_ACCUMULATE, GEMM_A_DIRECT, GEMM_B_DIRECT = 1, 2, 4, 8, 16

# one persistent workspace per (device, stream): its first 1 KB are the split-K tile tickets, which must start zero and which every
# call leaves zero; calls on one stream are ordered, so they can share it (a call on another stream gets its own)
_GEMM_WS: Dict[tuple, torch.Tensor] = {}
_GEMM_WS_RETIRED: list = []


def _gemm_workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _GEMM_WS.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            # a captured graph may replay kernels that point into the old buffer, and releasing a block of a destroyed graph's private
            # pool while another capture is running makes the allocator cudaFree inside the capture: outgrown buffers are kept
            _GEMM_WS_RETIRED.append(ws)
        ws = torch.empty(max(nbytes, 2 * ws.numel() if ws is not None else 0, 32 << 20), dtype=torch.uint8, device=dev)
        ws[:1024].zero_()        # the tickets (under graph capture this is one 1 KB memset node)
        _GEMM_WS[key] = ws
    return ws


def _gemm_operand(t: torch.Tensor):
    """2-D fp32 view -> (tensor, leading dimension, K-first?) without copying when it is row-major or a transposed row-major view."""
    if t.dtype != torch.float32:
        t = t.float()
    r, c = t.shape
    if (t.stride(1) == 1 or c == 1) and t.stride(0) >= c:
        return t, t.stride(0), False
    if (t.stride(0) == 1 or r == 1) and t.stride(1) >= r:
        return t, t.stride(1), True
    t = t.contiguous()
    return t, t.stride(0), False


def mm_nt(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
          accumulate: bool = False, a_direct: bool = False, b_direct: bool = False) -> torch.Tensor:
    """out [M, N] (+)= a [M, K] @ b [N, K]^T (+ bias [N]) on the library's TF32 tcgen05 GEMM (fp32 in / out).

    `a` / `b` may be transposed views (x.t()): they are read in place either way.  The three shapes of a linear layer:
    forward mm_nt(x, W, bias), input gradient mm_nt(dy, W.t()), weight gradient mm_nt(dy.t(), x.t(), out=W.grad, accumulate=True).
    Operands are rounded to nearest TF32 on the way in (a pack launch, or inside the kernel for the few-row operand of a skinny
    problem); `a_direct` / `b_direct`: that operand is already rounded (round_tf32 / pack_tf32) and, if aligned, is read in place."""
    L.require_cuda(a, "a")
    at, lda, ta = _gemm_operand(a)
    bt, ldb, tb = _gemm_operand(b)
    M, K = at.shape
    N = bt.shape[0]
    if bt.shape[1] != K:
        raise RuntimeError(f"dreamer_b200.mm_nt: inner dimensions differ ({tuple(at.shape)} x {tuple(bt.shape)}^T)")
    if out is None:
        if accumulate:
            raise RuntimeError("dreamer_b200.mm_nt: accumulate needs `out`")
        out = torch.empty((M, N), dtype=torch.float32, device=at.device)
    elif out.dtype != torch.float32 or out.shape != (M, N) or (out.stride(1) != 1 and N != 1) or out.stride(0) < N:
        raise RuntimeError("dreamer_b200.mm_nt: `out` must be a row-major fp32 [M, N] tensor")
    flags = (GEMM_TRANS_A if ta else 0) | (GEMM_TRANS_B if tb else 0) | (GEMM_ACCUMULATE if accumulate else 0) \
        | (GEMM_A_DIRECT if a_direct else 0) | (GEMM_B_DIRECT if b_direct else 0)
    lib = L.load()
    nbytes = lib.drm_gemm_tf32_workspace_bytes(M, N, K)
    ws = _gemm_workspace(at.device, nbytes)
    bs = L.f32c(bias) if bias is not None else None
    L.check(lib.drm_gemm_tf32(M, N, K, L.ptr(at), lda, L.ptr(bt), ldb, L.ptr(out), out.stride(0), L.ptr(bs), flags, L.ptr(ws),
                              ws.numel(), L.stream()), "gemm_tf32")
    return out


def mm(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False,
       a_direct: bool = False, b_direct: bool = False) -> torch.Tensor:
    """out (+)= a [M, K] @ b [K, N]"""
    return mm_nt(a, b.t(), out=out, accumulate=accumulate, a_direct=a_direct, b_direct=b_direct)


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [rows, in] @ weight [out, in]^T + bias (torch.nn.functional.linear on 2-D inputs)"""
    return mm_nt(x, weight, bias)


def pack_tf32(w: torch.Tensor) -> torch.Tensor:
    """K-major, TF32-rounded copy of the 2-D operand w [rows, K] (a transposed view is read in place) with a pitch that is a multiple
    of 4 (mm_nt(..., b_direct=True) then reads it in place; for weights that stay fixed over the steps of a recurrence)."""
    L.require_cuda(w, "w")
    wt, ld, tr = _gemm_operand(w)
    rows, K = wt.shape
    Kp = (K + 3) // 4 * 4
    out = torch.empty((rows, Kp), dtype=torch.float32, device=wt.device)
    L.check(L.load().drm_pack_tf32(rows, K, L.ptr(wt), ld, 1 if tr else 0, L.ptr(out), Kp, L.stream()), "pack_tf32")
    return out[:, :K]


# --------------------------------------------------------------------------------------------
# packed RSSM + rollout workspace
# --------------------------------------------------------------------------------------------
def _mlp_struct(sd: Dict[str, torch.Tensor], prefix: Optional[str], keep: list, final: bool = True) -> L.DrmMlpW:
    s = L.DrmMlpW()
    if prefix is None:
        return s
    names = [("w0", "0.weight"), ("b0", "0.bias"), ("g0", "1.weight"), ("be0", "1.bias"),
             ("w1", "3.weight"), ("b1", "3.bias"), ("g1", "4.weight"), ("be1", "4.bias")]
    if final:
        names += [("w2", "6.weight"), ("b2", "6.bias")]
    for field, key in names:
        t = L.f32c(sd[f"{prefix}.{key}"].detach())
        L.require_cuda(t, f"{prefix}.{key}")
        keep.append(t)
        setattr(s, field, t.data_ptr())
    return s


class _NoCopy:
    """Owners of a native handle: a copy would free the handle twice (double cudaFree / stale TMA descriptors)."""

    def __copy__(self):
        raise RuntimeError(f"{type(self).__name__} owns a native handle and cannot be copied; build a new one")

    def __deepcopy__(self, memo):
        raise RuntimeError(f"{type(self).__name__} owns a native handle and cannot be copied; build a new one")

    def __reduce__(self):
        raise RuntimeError(f"{type(self).__name__} owns a native handle and cannot be pickled")


class PackedRssm(_NoCopy):
    """drm_rssm handle: bf16 tile-packed copies of the RSSM / head weights.

    ``state_dict`` uses the reference's key names (``world_model.*`` / ``agent.*``).
    Call :meth:`pack` again after every optimiser step -- the packed weights are a cache.
    """

    PRECISIONS = {"bf16": 0, "tf32": 1}   # DRM_PRECISION_*

    def __init__(self, D: int, R: int, C_: int, A: int, NB: int, h_prior, h_head, precision: str = "bf16"):
        """precision "tf32": fp32 operands rounded to TF32 (tcgen05.mma kind::tf32) -- the class the reference's own GPU runs use
        (train_car_racer.py:13); rollout / step-level calls only, on the launch-per-stage kernels."""
        self.dims = L.DrmDims(D, R, C_, A, NB, (C.c_int32 * 2)(*h_prior), (C.c_int32 * 2)(*h_head))
        self.D, self.R, self.C, self.A, self.NB = D, R, C_, A, NB
        self.precision = precision
        self.handle = C.c_void_p()
        L.check(L.load().drm_rssm_create_ex(C.byref(self.dims), self.PRECISIONS[precision], C.byref(self.handle)), "rssm_create")

    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], R: int = 32, C_: int = 32, D: Optional[int] = None, A: int = 3,
                        precision: str = "bf16"):
        """Build from any subset of the reference state_dict (a full Dreamer, or one module's slice)."""
        g = "world_model.sequence_model.GRU."
        heads = [p for p in ("world_model.reward_predictor.logit_net", "world_model.continue_predictor.logit_generator",
                             "agent.actor.base_net", "agent.critic.value_net", "agent.target_critic.value_net")
                 if f"{p}.0.weight" in sd]
        if g + "weight_hh" in sd:
            D = sd[g + "weight_hh"].shape[1]
            A = sd[g + "weight_ih"].shape[1] - R * C_
        elif "world_model.dynamics_predictor.logit_net.0.weight" in sd:
            D = sd["world_model.dynamics_predictor.logit_net.0.weight"].shape[1]
        elif heads:
            D = sd[f"{heads[0]}.0.weight"].shape[1] - R * C_
        if D is None:
            raise RuntimeError("dreamer_b200: cannot infer the GRU width D from this state_dict")
        if "agent.actor.mu_head.weight" in sd:
            A = sd["agent.actor.mu_head.weight"].shape[0]
        NB = 255
        for k in ("world_model.reward_predictor.buckets_rew", "agent.critic.buckets_crit", "agent.target_critic.buckets_crit"):
            if k in sd:
                NB = sd[k].shape[0]
        pk = "world_model.dynamics_predictor.logit_net"
        hp = (sd[f"{pk}.0.weight"].shape[0], sd[f"{pk}.3.weight"].shape[0]) if f"{pk}.0.weight" in sd else (32, 32)
        sizes = {(sd[f"{p}.0.weight"].shape[0], sd[f"{p}.3.weight"].shape[0]) for p in heads} or {(32, 32)}
        if len(sizes) != 1:
            raise RuntimeError("dreamer_b200: reward / continue / actor / critic MLPs must share hidden sizes "
                               f"(got {sorted(sizes)}); the fused head stage batches them in one launch")
        obj = cls(D, R, C_, A, NB, hp, sizes.pop(), precision=precision)
        obj.pack(sd)
        return obj

    def pack(self, sd: Dict[str, torch.Tensor]):
        """(Re-)pack whatever weight groups `sd` contains (reference key names)."""
        keep = []
        w = L.DrmRssmWeights()

        def put(field, key):
            if key in sd:
                t = L.f32c(sd[key].detach()); L.require_cuda(t, key); keep.append(t); setattr(w, field, t.data_ptr())

        g = "world_model.sequence_model.GRU."
        for field, key in (("gru_w_ih", "weight_ih"), ("gru_w_hh", "weight_hh"), ("gru_b_ih", "bias_ih"), ("gru_b_hh", "bias_hh")):
            put(field, g + key)

        def mlp(prefix, final=True):
            return _mlp_struct(sd, prefix if f"{prefix}.0.weight" in sd else None, keep, final)

        w.prior = mlp("world_model.dynamics_predictor.logit_net")
        w.reward = mlp("world_model.reward_predictor.logit_net")
        w.cont = mlp("world_model.continue_predictor.logit_generator")
        w.actor = mlp("agent.actor.base_net", final=False)
        for field, key in (("actor_mu_w", "mu_head.weight"), ("actor_mu_b", "mu_head.bias"),
                           ("actor_ls_w", "log_sig_head.weight"), ("actor_ls_b", "log_sig_head.bias")):
            put(field, "agent.actor." + key)
        w.critic = mlp("agent.critic.value_net")
        w.target_critic = mlp("agent.target_critic.value_net")
        put("buckets_rew", "world_model.reward_predictor.buckets_rew")
        put("buckets_crit", "agent.critic.buckets_crit")
        if not w.buckets_crit:
            put("buckets_crit", "agent.target_critic.buckets_crit")
        L.check(L.load().drm_rssm_pack(self.handle, C.byref(w), L.stream()), "rssm_pack")
        self._keep = keep  # sources must stay alive until the pack kernels have run

    def __del__(self):
        try:
            if self.handle:
                L.load().drm_rssm_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


class Rollout(_NoCopy):
    """drm_rollout handle: state buffers + TMA descriptors for up to B rows and horizon H."""

    def __init__(self, model: PackedRssm, B: int, H: int):
        self.model, self.B, self.H = model, B, H
        self.handle = C.c_void_p()
        L.check(L.load().drm_rollout_create(model.handle, B, H, C.byref(self.handle)), "rollout_create")

    def info(self):
        """Which kernel path `run` takes: dict(persistent, gru_tile, sample_tile, ctas, timeouts=[(code, seen, want, thread, cta), ...])."""
        buf = (C.c_uint32 * 45)()
        L.check(L.load().drm_rollout_info(self.handle, buf, 45), "rollout_info")
        recs = [tuple(buf[5 + 5 * i + k] for k in range(5)) for i in range(8)]
        return dict(persistent=bool(buf[0]), gru_tile=int(buf[1]), sample_tile=int(buf[2]), ctas=int(buf[3]), n_timeouts=int(buf[4]),
                    timeouts=[r for r in recs if any(r)])

    def run_graphed(self, z0, h0, uniforms, normals, want_idx: bool = True):
        """`run` replayed as ONE CUDA graph after two eager calls (graphs.StepGraph): the inputs are copied into static device
        buffers and the ~110 launches of a rollout become one graph launch (0.99 -> 0.94 ms at 1024 x 15, and no host launch
        work).  The returned tensors are the graph's static outputs: they are overwritten by the next call."""
        from .graphs import StepGraph
        cache = self.__dict__.setdefault("_graphs", {})
        if want_idx not in cache:
            cache[want_idx] = StepGraph(lambda a, b, c, d: self.run(a, b, c, d, want_idx=want_idx), warmup=2)
        return cache[want_idx](L.f32c(z0), L.f32c(h0), L.f32c(uniforms), L.f32c(normals))

    def run(self, z0, h0, uniforms, normals, want_idx: bool = True):
        """Dreamer.dream_episodes (Dreamer.py:143-175).

        z0 (B,1,R,C) or (B,R*C); h0 (B,1,D) or (B,D); uniforms (H,B,R); normals (H,B,A).
        Returns the reference 7-tuple (+ idx (B,H,R) uint8 when want_idx).
        """
        m, B, H = self.model, self.B, self.H
        L.require_cuda(z0, "z0")
        dev = z0.device
        z0f = L.f32c(z0).reshape(B, m.R * m.C)
        h0f = L.f32c(h0).reshape(B, m.D)
        u = L.f32c(uniforms)
        n = L.f32c(normals)
        if tuple(u.shape) != (H, B, m.R) or tuple(n.shape) != (H, B, m.A):
            raise RuntimeError(f"dreamer_b200.Rollout.run: uniforms must be {(H, B, m.R)} and normals {(H, B, m.A)}")
        f = dict(dtype=torch.float32, device=dev)
        latent = torch.empty((B, H + 1, m.R, m.C), **f)
        hidden = torch.empty((B, H + 1, m.D), **f)
        actions = torch.empty((B, H, m.A), **f)
        mu = torch.empty((B, H, m.A), **f)
        sigma = torch.empty((B, H, m.A), **f)
        rewards = torch.empty((B, H, 1), **f)
        conts = torch.empty((B, H, 1), **f)
        idx = torch.empty((B, H, m.R), dtype=torch.uint8, device=dev) if want_idx else None
        L.check(L.load().drm_rollout_run(self.handle, L.ptr(z0f), L.ptr(h0f), L.ptr(u), L.ptr(n), L.ptr(latent), L.ptr(hidden),
                                         L.ptr(actions), L.ptr(rewards), L.ptr(conts), L.ptr(mu), L.ptr(sigma), L.ptr(idx),
                                         L.stream()), "rollout_run")
        out = (latent, hidden, actions, rewards, conts, mu, sigma)
        return out + (idx,) if want_idx else out

    # ---- step-level calls -------------------------------------------------------------------
    def gru_step(self, z, h, a):
        m = self.model
        N = h.shape[0]
        zf, hf, af = L.f32c(z).reshape(N, -1), L.f32c(h), L.f32c(a)
        out = torch.empty_like(hf)
        L.check(L.load().drm_gru_step(self.handle, L.ptr(zf), L.ptr(hf), L.ptr(af), L.ptr(out), N, L.stream()), "gru_step")
        return out

    def prior(self, h, uniforms=None, want_logits: bool = True):
        m = self.model
        N = h.shape[0]
        hf = L.f32c(h)
        logits = torch.empty((N, m.R, m.C), dtype=torch.float32, device=hf.device) if want_logits else None
        z = idx = None
        u = None
        if uniforms is not None:
            u = L.f32c(uniforms)
            z = torch.empty((N, m.R, m.C), dtype=torch.float32, device=hf.device)
            idx = torch.empty((N, m.R), dtype=torch.uint8, device=hf.device)
        L.check(L.load().drm_prior(self.handle, L.ptr(hf), L.ptr(u), L.ptr(logits), L.ptr(z), L.ptr(idx), N, L.stream()), "prior")
        return dict(logits=logits, z=z, idx=idx)

    def heads(self, h, z, heads: int, normals=None, want_logits: bool = False):
        m = self.model
        N = h.shape[0]
        hf, zf = L.f32c(h), L.f32c(z).reshape(N, -1)
        f = dict(dtype=torch.float32, device=hf.device)
        o = L.DrmHeadsOut()
        res = {}

        def alloc(name, shape):
            t = torch.empty(shape, **f)
            res[name] = t
            setattr(o, name, t.data_ptr())

        if heads & L.HEAD_REWARD:
            alloc("reward", (N, 1))
            if want_logits:
                alloc("reward_logits", (N, m.NB))
        if heads & L.HEAD_CONT:
            alloc("cont_prob", (N, 1)); alloc("cont_logit", (N, 1))
        nf = None
        if heads & L.HEAD_ACTOR:
            alloc("mu", (N, m.A)); alloc("sigma", (N, m.A))
            if normals is not None:
                nf = L.f32c(normals)
                alloc("action", (N, m.A))
        if heads & L.HEAD_CRITIC:
            alloc("value", (N, 1))
            if want_logits:
                alloc("value_logits", (N, m.NB))
        if heads & L.HEAD_TARGET_CRITIC:
            alloc("target_value", (N, 1))
        L.check(L.load().drm_heads(self.handle, L.ptr(hf), L.ptr(zf), L.ptr(nf), heads, C.byref(o), N, L.stream()), "heads")
        return res

    def __del__(self):
        try:
            if self.handle:
                L.load().drm_rollout_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


# --------------------------------------------------------------------------------------------
# (3) VAE encoder / decoder + observe scan
# --------------------------------------------------------------------------------------------
def neg_sse_rows(a: torch.Tensor, b: torch.Tensor, row_dims: int = 3) -> torch.Tensor:
    """-sum((a - b)^2) over the last `row_dims` dims (WorldModel.py:129)."""
    L.require_cuda(a, "a")
    x, y = L.f32c(a), L.f32c(b)
    lead = x.shape[:-row_dims]
    length = 1
    for d in x.shape[-row_dims:]:
        length *= d
    out = torch.empty(lead, dtype=torch.float32, device=x.device)
    L.check(L.load().drm_neg_sse_rows(L.ptr(x), L.ptr(y), L.ptr(out), out.numel(), length, L.stream()), "neg_sse_rows")
    return out


class PackedVae(_NoCopy):
    """drm_vae handle: packed Encoder / Decoder weights (reference keys world_model.encoder.*, world_model.decoder.*)."""

    def __init__(self, model: PackedRssm, H: int, W: int, e1: int, e2: int, d1: int, d2: int, h_enc: int, h_dec: int):
        self.model = model
        self.dims = L.DrmVaeDims(H, W, e1, e2, d1, d2, h_enc, h_dec)
        self.H, self.W = H, W
        self.handle = C.c_void_p()
        L.check(L.load().drm_vae_create(model.handle, C.byref(self.dims), C.byref(self.handle)), "vae_create")

    @classmethod
    def from_state_dict(cls, model: PackedRssm, sd: Dict[str, torch.Tensor], obs_hw=(64, 64)):
        e = "world_model.encoder."
        d = "world_model.decoder."
        e1 = sd[e + "feature_extractor.0.weight"].shape[0]
        e2 = sd[e + "feature_extractor.2.weight"].shape[0]
        d2 = sd[d + "image_builder.2.weight"].shape[1]
        d1 = sd[d + "image_builder.4.weight"].shape[1]
        obj = cls(model, obs_hw[0], obs_hw[1], e1, e2, d1, d2, sd[e + "latent_mapper.0.weight"].shape[0],
                  sd[d + "upscaler.0.weight"].shape[0])
        obj.pack(sd)
        return obj

    def pack(self, sd: Dict[str, torch.Tensor]):
        keep = []
        w = L.DrmVaeWeights()

        def dp(key):
            t = L.f32c(sd[key].detach()); L.require_cuda(t, key); keep.append(t)
            return t.data_ptr()

        e = "world_model.encoder."
        d = "world_model.decoder."
        for i, k in enumerate((0, 2, 4, 6)):
            w.enc_conv_w[i] = dp(f"{e}feature_extractor.{k}.weight"); w.enc_conv_b[i] = dp(f"{e}feature_extractor.{k}.bias")
            w.dec_conv_w[i] = dp(f"{d}image_builder.{k}.weight"); w.dec_conv_b[i] = dp(f"{d}image_builder.{k}.bias")
        for field, key in (("enc_l1_w", e + "latent_mapper.0.weight"), ("enc_l1_b", e + "latent_mapper.0.bias"),
                           ("enc_ln_g", e + "latent_mapper.1.weight"), ("enc_ln_b", e + "latent_mapper.1.bias"),
                           ("enc_l2_w", e + "latent_mapper.3.weight"), ("enc_l2_b", e + "latent_mapper.3.bias"),
                           ("dec_l1_w", d + "upscaler.0.weight"), ("dec_l1_b", d + "upscaler.0.bias"),
                           ("dec_ln_g", d + "upscaler.1.weight"), ("dec_ln_b", d + "upscaler.1.bias"),
                           ("dec_l2_w", d + "upscaler.3.weight"), ("dec_l2_b", d + "upscaler.3.bias")):
            setattr(w, field, dp(key))
        L.check(L.load().drm_vae_pack(self.handle, C.byref(w), L.stream()), "vae_pack")
        self._keep = keep

    def __del__(self):
        try:
            if self.handle:
                L.load().drm_vae_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


class Observe(_NoCopy):
    """drm_observe handle: workspace for B sequences x T steps (posterior scan, batched heads, encoder / decoder calls)."""

    def __init__(self, vae: PackedVae, B: int, T: int):
        self.vae, self.model, self.B, self.T = vae, vae.model, B, T
        self.handle = C.c_void_p()
        L.check(L.load().drm_observe_create(self.model.handle, vae.handle, B, T, C.byref(self.handle)), "observe_create")

    def scan(self, obs, act, uniforms, warm_start: bool = False, want_logits: bool = True, want_idx: bool = True):
        """obs (B,T,3,H,W) normalised fp32, act (B,T,A), uniforms (T,B,R) -> dict(latent (B,T,R,C), hidden (B,T,D), logits, idx)."""
        m, B, T = self.model, self.B, self.T
        L.require_cuda(obs, "obs")
        o, a, u = L.f32c(obs), L.f32c(act), L.f32c(uniforms)
        if tuple(o.shape[:2]) != (B, T) or tuple(a.shape[:2]) != (B, T) or tuple(u.shape) != (T, B, m.R):
            raise RuntimeError(f"dreamer_b200.Observe.scan: expected obs/act with leading {(B, T)} and uniforms {(T, B, m.R)}")
        f = dict(dtype=torch.float32, device=o.device)
        latent = torch.empty((B, T, m.R, m.C), **f)
        hidden = torch.empty((B, T, m.D), **f)
        logits = torch.empty((B, T, m.R, m.C), **f) if want_logits else None
        idx = torch.empty((B, T, m.R), dtype=torch.uint8, device=o.device) if want_idx else None
        L.check(L.load().drm_observe_scan(self.handle, L.ptr(o), L.ptr(a), L.ptr(u), 1 if warm_start else 0, L.ptr(latent), L.ptr(hidden),
                                          L.ptr(logits), L.ptr(idx), L.stream()), "observe_scan")
        return dict(latent=latent, hidden=hidden, logits=logits, idx=idx)

    def heads(self, prior=True, decoder=True, reward=True, cont=True):
        """Batched heads on the states of the last scan (WorldModel.py:116-119)."""
        m, B, T = self.model, self.B, self.T
        dev = torch.device("cuda", torch.cuda.current_device())
        f = dict(dtype=torch.float32, device=dev)
        out = dict(prior_logits=torch.empty((B, T, m.R, m.C), **f) if prior else None,
                   dec_mu=torch.empty((B, T, 3, self.vae.H, self.vae.W), **f) if decoder else None,
                   reward_logits=torch.empty((B, T - 1, m.NB), **f) if reward and T > 1 else None,
                   cont_logit=torch.empty((B, T - 1, 1), **f) if cont and T > 1 else None)
        L.check(L.load().drm_observe_heads(self.handle, L.ptr(out["prior_logits"]), L.ptr(out["dec_mu"]), L.ptr(out["reward_logits"]),
                                           L.ptr(out["cont_logit"]), L.stream()), "observe_heads")
        return out

    def encode(self, h, obs, uniforms=None):
        """Encoder.forward / .encode: h (N,D), obs (N,3,H,W) -> dict(logits (N,R,C)[, z, idx])."""
        m = self.model
        N = h.shape[0]
        hf, of = L.f32c(h), L.f32c(obs)
        f = dict(dtype=torch.float32, device=hf.device)
        logits = torch.empty((N, m.R, m.C), **f)
        z = idx = u = None
        if uniforms is not None:
            u = L.f32c(uniforms)
            z = torch.empty((N, m.R, m.C), **f)
            idx = torch.empty((N, m.R), dtype=torch.uint8, device=hf.device)
        L.check(L.load().drm_encoder_fwd(self.handle, L.ptr(hf), L.ptr(of), L.ptr(u), L.ptr(logits), L.ptr(z), L.ptr(idx), N, L.stream()),
                "encoder_fwd")
        return dict(logits=logits, z=z, idx=idx)

    def decode(self, h, z):
        """Decoder.forward: h (N,D), z (N,R,C) -> mu (N,3,H,W)."""
        N = h.shape[0]
        hf, zf = L.f32c(h), L.f32c(z).reshape(N, -1)
        mu = torch.empty((N, 3, self.vae.H, self.vae.W), dtype=torch.float32, device=hf.device)
        L.check(L.load().drm_decoder_fwd(self.handle, L.ptr(hf), L.ptr(zf), L.ptr(mu), N, L.stream()), "decoder_fwd")
        return mu

    def __del__(self):
        try:
            if self.handle:
                L.load().drm_observe_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


def actor_head_bwd(g_mu, g_sigma, da, a, eps, log_sigma, out):
    """d_head (rows, 2A) = [dmu | dls] of one BPTT step through the actor head (drm_actor_head_bwd); da may be None."""
    rows, A = g_mu.shape
    for t in (g_mu, g_sigma, da, a, eps, log_sigma, out):
        if t is not None and not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
            raise RuntimeError("dreamer_b200.actor_head_bwd: fp32 contiguous CUDA tensors are required")
    if out.shape != (rows, 2 * A):
        raise RuntimeError("dreamer_b200.actor_head_bwd: out must be (rows, 2A)")
    L.check(L.load().drm_actor_head_bwd(L.ptr(g_mu), L.ptr(g_sigma), L.ptr(da), L.ptr(a), L.ptr(eps), L.ptr(log_sigma), L.ptr(out), rows, A,
                                        L.stream()), "actor_head_bwd")
    return out
