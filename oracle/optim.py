"""ORACLE (test infrastructure, never shipped): CPU restatement of the reference's optimiser tail.

    nn.utils.clip_grad_norm_(params, 100.0)            WorldModel.py:198, Agent.py:147-148
    torch.optim.AdamW(..., weight_decay=1e-6).step()   WorldModel.py:46,199; Agent.py:30-31,150-151
    Agent.soft_update_target(tau=0.02)                 Agent.py:90-94

The arithmetic is PyTorch's (requirements.txt: torch, unpinned; installed 2.11.0): torch/nn/utils/clip_grad.py
(total norm = 2-norm of the per-tensor 2-norms, coefficient max_norm / (total + 1e-6) clamped to 1) and
torch/optim/adamw.py `_single_tensor_adamw` (decoupled decay, lerp first moment, bias-corrected step).
Pinned in tests/test_oracle_golden.py::test_optim_oracle_matches_torch against the installed torch on CPU.
"""
import numpy as np


def clip_coef(grads, max_norm=100.0):
    total = np.sqrt(sum(float(np.sum(g.astype(np.float64) ** 2)) for g in grads)).astype(np.float32)
    return np.float32(min(1.0, float(np.float32(max_norm) / (total + np.float32(1e-6))))), total


def adamw_step(params, grads, exp_avg, exp_avg_sq, step, lr, betas, eps, weight_decay, max_norm=100.0):
    """One clipped AdamW step over lists of fp32 arrays (updated in place).  Returns (new step count, grad norm)."""
    coef, total = clip_coef(grads, max_norm)
    if not np.isfinite(total):
        return step, total
    step += 1
    b1, b2 = np.float32(betas[0]), np.float32(betas[1])
    bc1 = 1.0 - float(betas[0]) ** step
    bc2 = 1.0 - float(betas[1]) ** step
    step_size = np.float32(lr / bc1)
    inv_sqrt_bc2 = np.float32(1.0 / np.sqrt(bc2))
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        g = g * coef
        p *= np.float32(1.0 - lr * weight_decay)
        m += (np.float32(1.0) - b1) * (g - m)
        v *= b2
        v += (np.float32(1.0) - b2) * g * g
        denom = np.sqrt(v) * inv_sqrt_bc2 + np.float32(eps)
        p -= step_size * (m / denom)
    return step, total


def soft_update(target, current, tau=0.02):
    for t, c in zip(target, current):
        t *= np.float32(1.0 - tau)
        t += np.float32(tau) * c
