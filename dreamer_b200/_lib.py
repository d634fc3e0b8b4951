"""ctypes binding of libdreamer_b200.so (the C-ABI declared in include/dreamer_b200.h).

There is no CPU fallback: if the shared library is missing or the device is not sm_100-class,
every call raises.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdreamer_b200.so")

c_f32p = C.c_void_p
c_stream = C.c_void_p


class DrmDims(C.Structure):
    _fields_ = [("D", C.c_int32), ("R", C.c_int32), ("C", C.c_int32), ("A", C.c_int32), ("NB", C.c_int32),
                ("h_prior", C.c_int32 * 2), ("h_head", C.c_int32 * 2)]


class DrmMlpW(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w0", "b0", "g0", "be0", "w1", "b1", "g1", "be1", "w2", "b2")]


class DrmRssmWeights(C.Structure):
    _fields_ = [("gru_w_ih", C.c_void_p), ("gru_w_hh", C.c_void_p), ("gru_b_ih", C.c_void_p), ("gru_b_hh", C.c_void_p),
                ("prior", DrmMlpW), ("reward", DrmMlpW), ("cont", DrmMlpW), ("actor", DrmMlpW),
                ("actor_mu_w", C.c_void_p), ("actor_mu_b", C.c_void_p), ("actor_ls_w", C.c_void_p), ("actor_ls_b", C.c_void_p),
                ("critic", DrmMlpW), ("target_critic", DrmMlpW),
                ("buckets_rew", C.c_void_p), ("buckets_crit", C.c_void_p)]


class DrmHeadsOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("reward", "reward_logits", "cont_prob", "cont_logit", "mu", "sigma", "action",
                                          "value", "value_logits", "target_value")]


class DrmVaeDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("H", "W", "e1", "e2", "d1", "d2", "h_enc", "h_dec")]


class DrmVaeWeights(C.Structure):
    _fields_ = ([("enc_conv_w", C.c_void_p * 4), ("enc_conv_b", C.c_void_p * 4)] +
                [(n, C.c_void_p) for n in ("enc_l1_w", "enc_l1_b", "enc_ln_g", "enc_ln_b", "enc_l2_w", "enc_l2_b",
                                           "dec_l1_w", "dec_l1_b", "dec_ln_g", "dec_ln_b", "dec_l2_w", "dec_l2_b")] +
                [("dec_conv_w", C.c_void_p * 4), ("dec_conv_b", C.c_void_p * 4)])


HEAD_REWARD, HEAD_CONT, HEAD_ACTOR, HEAD_CRITIC, HEAD_TARGET_CRITIC = 1, 2, 4, 8, 16

# name -> (restype, argtypes); kept in sync with include/dreamer_b200.h (tests/test_cabi.py checks
# that every symbol the header declares is exported and listed here).
SIGNATURES = {
    "drm_abi_version": (C.c_int, []),
    "drm_last_error": (C.c_char_p, []),
    "drm_device_check": (C.c_int, []),
    "drm_launch_count": (C.c_int64, []),
    "drm_set_option": (C.c_int, [C.c_char_p, C.c_int32]),
    "drm_debug_timeline": (C.c_int, [C.c_int32, C.c_void_p]),
    "drm_profile_enable": (C.c_int, [C.c_int32]),
    "drm_profile_read": (C.c_int, [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "drm_categorical32_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, c_stream]),
    "drm_onehot32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, c_stream]),
    "drm_categorical32_st": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, c_stream]),
    "drm_categorical32_bwd": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, c_stream]),
    "drm_ln_silu_bwd": (C.c_int, [C.c_void_p] * 6 + [C.c_int64, C.c_int32, C.c_float, c_stream]),
    "drm_actor_head_bwd": (C.c_int, [C.c_void_p] * 7 + [C.c_int64, C.c_int32, c_stream]),
    "drm_gru_bwd": (C.c_int, [C.c_void_p] * 7 + [C.c_int32, C.c_int64, C.c_int32, c_stream]),
    "drm_gru_bwd_add": (C.c_int, [C.c_void_p] * 8 + [C.c_int32, C.c_int64, C.c_int32, c_stream]),
    "drm_categorical32_kl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, c_stream]),
    "drm_replay_gather": (C.c_int, [C.c_void_p] * 9 + [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, c_stream]),
    "drm_replay_insert": (C.c_int, [C.c_void_p] * 8 + [C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_int32, c_stream]),
    "drm_lambda_return": (C.c_int, [C.c_void_p] * 4 + [C.c_int32, C.c_int32, C.c_float, C.c_float, c_stream]),
    "drm_tanh_normal_logp": (C.c_int, [C.c_void_p] * 7 + [C.c_int64, C.c_int32, c_stream]),
    "drm_percentile_pair": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, c_stream]),
    "drm_ln_silu_bwd_affine": (C.c_int, [C.c_void_p] * 7 + [C.c_int64, C.c_int32, C.c_float, c_stream]),
    "drm_convt_image_fwd": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] * 5 + [c_stream]),
    "drm_colsum_bf16_scratch_bytes": (C.c_int64, [C.c_int64, C.c_int32]),
    "drm_colsum_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, c_stream]),
    "drm_colsum_scratch_bytes": (C.c_int64, [C.c_int64, C.c_int32]),
    "drm_colsum": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, c_stream]),
    "drm_twohot_ce_bwd": (C.c_int, [C.c_void_p] * 5 + [C.c_float, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, c_stream]),
    "drm_twohot_ce": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int32, C.c_int32, c_stream]),
    "drm_bucket_value": (C.c_int, [C.c_void_p] * 3 + [C.c_int64, C.c_int32, c_stream]),
    "drm_rssm_create": (C.c_int, [C.POINTER(DrmDims), C.POINTER(C.c_void_p)]),
    "drm_rssm_create_ex": (C.c_int, [C.POINTER(DrmDims), C.c_int32, C.POINTER(C.c_void_p)]),
    "drm_rssm_pack": (C.c_int, [C.c_void_p, C.POINTER(DrmRssmWeights), c_stream]),
    "drm_rssm_destroy": (C.c_int, [C.c_void_p]),
    "drm_rollout_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "drm_rollout_destroy": (C.c_int, [C.c_void_p]),
    "drm_rollout_run": (C.c_int, [C.c_void_p] * 13 + [c_stream]),
    "drm_rollout_info": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "drm_rollout_trace": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]),
    "drm_observe_trace": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]),
    "drm_gru_step": (C.c_int, [C.c_void_p] * 5 + [C.c_int32, c_stream]),
    "drm_prior": (C.c_int, [C.c_void_p] * 6 + [C.c_int32, c_stream]),
    "drm_heads": (C.c_int, [C.c_void_p] * 4 + [C.c_int32, C.POINTER(DrmHeadsOut), C.c_int32, c_stream]),
    "drm_vae_create": (C.c_int, [C.c_void_p, C.POINTER(DrmVaeDims), C.POINTER(C.c_void_p)]),
    "drm_vae_pack": (C.c_int, [C.c_void_p, C.POINTER(DrmVaeWeights), c_stream]),
    "drm_vae_destroy": (C.c_int, [C.c_void_p]),
    "drm_observe_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "drm_observe_destroy": (C.c_int, [C.c_void_p]),
    "drm_observe_scan": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] + [C.c_void_p] * 4 + [c_stream]),
    "drm_observe_heads": (C.c_int, [C.c_void_p] * 5 + [c_stream]),
    "drm_encoder_fwd": (C.c_int, [C.c_void_p] * 7 + [C.c_int32, c_stream]),
    "drm_decoder_fwd": (C.c_int, [C.c_void_p] * 4 + [C.c_int32, c_stream]),
    "drm_neg_sse_rows": (C.c_int, [C.c_void_p] * 3 + [C.c_int64, C.c_int32, c_stream]),
    "drm_adamw_scratch_bytes": (C.c_int64, []),
    "drm_adamw_step": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_void_p, C.c_void_p] + [C.c_float] * 6 + [C.c_void_p, C.c_float, C.c_int32, c_stream]),
    "drm_grad_norm": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, c_stream]),
    "drm_gemm_tf32_workspace_bytes": (C.c_int64, [C.c_int32] * 3),
    "drm_gemm_tf32": (C.c_int, [C.c_int32] * 3 + [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                C.c_int32, C.c_void_p, C.c_int64, c_stream]),
    "drm_pack_tf32": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, c_stream]),
    "drm_test_gemm": (C.c_int, [C.c_void_p] * 4 + [C.c_int32, C.c_int32, C.c_int32, c_stream]),
}

_lib = None


def load():
    """Load the shared library (raises if it has not been built: `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C dreamer_b200/csrc` "
                               "(__graft_entry__.build()). dreamer_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().drm_last_error()
        raise RuntimeError(f"dreamer_b200 {what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32, contiguous, on the current CUDA device."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"dreamer_b200: `{name}` must be a CUDA tensor (there is no CPU path)")
