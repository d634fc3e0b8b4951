"""Data parallelism for the hot path: one process per GPU, start states / replay sequences sharded by
rank, NO collective in the forward data path.  Collectives (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU
tests) are used only for
  * one flat gradient bucket all-reduce (SUM) per optimiser group after backward, and
  * one packed vector of scalars per loss (mask sum, KL sums, element counts, NaN flag) so that every rank evaluates the
    *global* loss -- the free-bits ``max(1, mean KL)`` of WorldModel.py:187-188 is nonlinear in the global mean, and all
    ranks must take the same NaN/Inf skip decision (WorldModel.py:191; Agent.py:137).
The reference has no distributed code at all; single-process runs take the no-op branches below.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


_FORCE_SINGLE = False   # tests flip this to evaluate the single-process (full-batch) path inside a distributed job


def is_dist() -> bool:
    return (not _FORCE_SINGLE) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world() -> int:
    return dist.get_world_size() if is_dist() else 1


def rank() -> int:
    return dist.get_rank() if is_dist() else 0


def shard_bounds(n: int, r: int = None, w: int = None):
    """Contiguous shard [lo, hi) of n units for rank r of w (the first n % w ranks get one extra unit)."""
    r = rank() if r is None else r
    w = world() if w is None else w
    base, extra = divmod(n, w)
    lo = r * base + min(r, extra)
    return lo, lo + base + (1 if r < extra else 0)


def shard(t: torch.Tensor, dim: int = 0, r: int = None, w: int = None) -> torch.Tensor:
    """This rank's slice of a global-batch tensor along `dim` (uniforms / normals are generated for the global batch and
    sliced, so the per-rank results concatenate to the single-GPU result)."""
    lo, hi = shard_bounds(t.shape[dim], r, w)
    return t.narrow(dim, lo, hi - lo)


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if is_dist():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def all_gather_cat(t: torch.Tensor) -> torch.Tensor:
    """Concatenate equally-shaped per-rank tensors along dim 0 (used for the global return percentiles, Agent.py:82-84)."""
    if not is_dist():
        return t
    parts = [torch.empty_like(t) for _ in range(world())]
    dist.all_gather(parts, t.contiguous())
    return torch.cat(parts, 0)


class FlatBucket:
    """Gradients of a parameter group flattened into one contiguous fp32 buffer: one all-reduce per optimiser group
    (world model 7.76 M parameters = 31 MB, actor 1.47 MB, critic 1.67 MB at the reference sizes)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None

    def all_reduce(self, average: bool = False):
        """SUM (or mean) the gradients over ranks in place.  Parameters without a gradient contribute zeros."""
        if not self.params:
            return
        dev, dt = self.params[0].device, torch.float32
        if self.flat is None or self.flat.device != dev:
            self.flat = torch.zeros(self.numel, dtype=dt, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        all_reduce_sum_(self.flat)
        if average:
            self.flat.div_(world())
        off = 0
        for p in self.params:
            n = p.numel()
            g = self.flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n


def world_model_loss_from_sums(local: torch.Tensor, betas: Sequence[float]):
    """Global WorldModel loss from per-rank sums.

    local = [obs_ll_sum, rew_ll_sum, cont_bce_sum, mask_sum, kl_sum (masked), n_elements] of this rank's sequences
    (WorldModel.py:170-188).  Returns (total, dict) where every rank holds the same values."""
    g = all_reduce_sum_(local.clone())
    denom = g[3] + 1e-5
    loss_pred = (-g[0] - g[1] + g[2]) / denom
    kl_mean = g[4] / g[5]
    one = torch.ones((), dtype=g.dtype, device=g.device)
    total = betas[0] * loss_pred + betas[1] * torch.maximum(one, kl_mean) + betas[2] * torch.maximum(one, kl_mean)
    return total, dict(loss_pred=loss_pred, kl_mean=kl_mean, denom=denom, n_elements=g[5])


def any_rank_flag(flag: bool, device) -> bool:
    """True on every rank if `flag` is True on any rank (the shared NaN/Inf skip decision)."""
    t = torch.tensor([1.0 if flag else 0.0], device=device)
    all_reduce_sum_(t)
    return bool(t.item() > 0)
