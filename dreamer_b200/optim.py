"""Flat-buffer optimiser for the three parameter groups of the hot path.

``FlatAdamW`` replaces, for one parameter group,

    nn.utils.clip_grad_norm_(params, 100.0); torch.optim.AdamW(...).step()        WorldModel.py:195-200, Agent.py:141-151
    Agent.soft_update_target(tau)                                                  Agent.py:90-94   (with ``ema_params``)

by ``drm_adamw_step``: three launches on ONE flat fp32 buffer, no host synchronisation, gradients zeroed in the same pass.
The parameters (and their ``.grad``) are re-pointed at slices of the flat buffers, so autograd accumulates straight into the
bucket the data-parallel all-reduce sends (dist.FlatBucket needs no copy in or out), and the whole training tail is
capturable in a CUDA graph.  A step whose gradient norm is not finite is skipped on the device (what GradScaler /
the reference's NaN checks do on the host).

It is a ``torch.optim.Optimizer``: ``param_groups`` (lr, betas, eps, weight_decay, max_norm are read at every step),
``state_dict()`` / ``load_state_dict()`` use torch.optim.AdamW's layout (``step``, ``exp_avg``, ``exp_avg_sq`` per parameter).
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from . import _lib as L
from . import dist as D

_ALIGN = 4  # elements: every parameter starts on a 16-byte boundary (vector loads, TMA-friendly views)


def _flatten(params, flat: Optional[torch.Tensor] = None):
    """Re-point every parameter's storage at a slice of one flat fp32 buffer (values preserved)."""
    offs, off = [], 0
    for p in params:
        offs.append(off)
        off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
    if flat is None:
        flat = torch.zeros(max(off, _ALIGN), dtype=torch.float32, device=params[0].device)
    with torch.no_grad():
        for p, o in zip(params, offs):
            view = flat[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
    return flat, offs


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_norm: float = 100.0, ema_params: Optional[Iterable[torch.nn.Parameter]] = None,
                 tau: float = 0.02):
        params = [p for p in params]
        if not params:
            raise ValueError("FlatAdamW: empty parameter list")
        for p in params:
            L.require_cuda(p, "parameter")
            if p.dtype != torch.float32:
                raise TypeError("FlatAdamW: parameters must be fp32")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, max_norm=max_norm, tau=tau))
        self._params = params
        self.flat, self._offs = _flatten(params)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.opt_state = torch.zeros(8, dtype=torch.float32, device=self.flat.device)   # see drm_adamw_step
        self._scratch = torch.zeros(int(L.load().drm_adamw_scratch_bytes()) // 8, dtype=torch.float64, device=self.flat.device)
        self.ema, self._ema_params = None, None
        if ema_params is not None:
            ema_params = [p for p in ema_params]
            if [p.shape for p in ema_params] != [p.shape for p in params]:
                raise ValueError("FlatAdamW: ema_params must mirror params")
            self.ema, _ = _flatten(ema_params)
            self._ema_params = ema_params
        for p, o in zip(params, self._offs):
            p.grad = self.grad[o:o + p.numel()].view(p.shape)
            self.state[p] = {"step": self.opt_state[0], "exp_avg": self.exp_avg[o:o + p.numel()].view(p.shape),
                             "exp_avg_sq": self.exp_avg_sq[o:o + p.numel()].view(p.shape)}

    # ---- torch.optim.Optimizer surface ----------------------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = False):
        """Gradients live in the flat bucket: they are zeroed in place (step() already leaves them zeroed)."""
        self._rebind_grads()
        self.grad.zero_()

    def _check_aliasing(self):
        """The kernel updates `self.flat` through raw pointers: a parameter whose storage was replaced since construction
        (module.to(...), .half(), load_state_dict(assign=True)) would silently stop training.  Re-point it at its slice (keeping
        its current values) when only the storage moved; refuse when its dtype / device / size changed."""
        for p, o in zip(self._params, self._offs):
            if p.data_ptr() == self.flat.data_ptr() + 4 * o:
                continue
            if p.dtype != torch.float32 or p.device != self.flat.device or p.numel() != self.flat[o:o + p.numel()].numel():
                raise RuntimeError("FlatAdamW: a parameter no longer matches the flat bucket (dtype / device / size changed); "
                                   "rebuild the optimiser after converting the module")
            view = self.flat[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view

    def _rebind_grads(self):
        self._check_aliasing()
        for p, o in zip(self._params, self._offs):
            g = p.grad
            if g is None or g.data_ptr() != self.grad.data_ptr() + 4 * o:
                view = self.grad[o:o + p.numel()].view(p.shape)
                if g is not None:           # somebody replaced .grad (zero_grad(set_to_none=True) + backward): keep its value
                    view.copy_(g)
                p.grad = view

    @torch.no_grad()
    def step(self, closure=None, all_reduce: bool = False, zero_grad: bool = True):
        """clip + AdamW (+ EMA of ``ema_params``) on the flat bucket.  ``all_reduce`` sums the bucket over ranks first
        (each rank's loss already carries its share of the global normalisation, dist.py)."""
        if closure is not None:
            raise NotImplementedError("FlatAdamW.step: closures are not supported")
        self._rebind_grads()
        if all_reduce and D.is_dist():
            D.all_reduce_sum_(self.grad)
        g = self.param_groups[0]
        L.check(L.load().drm_adamw_step(L.ptr(self.flat), L.ptr(self.grad), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq),
                                        self.flat.numel(), L.ptr(self.opt_state), L.ptr(self._scratch), g["lr"], g["betas"][0],
                                        g["betas"][1], g["eps"], g["weight_decay"], g["max_norm"] or 0.0, L.ptr(self.ema),
                                        g["tau"], 1 if zero_grad else 0, L.stream()), "adamw_step")
        self.mark_updated()

    def mark_updated(self):
        """The kernel writes through raw pointers: bump the autograd version counters so packed-weight caches
        (modules._Packed, _VaeEngine) see the update.  Call it after replaying a CUDA graph that contains step()."""
        torch.autograd.graph.increment_version(self._params)
        if self._ema_params is not None:
            torch.autograd.graph.increment_version(self._ema_params)

    def grad_norm(self) -> torch.Tensor:
        """||g||_2 of the flat bucket (device scalar; after step(): the norm that step saw is ``last_grad_norm``)."""
        out = torch.empty(1, dtype=torch.float32, device=self.flat.device)
        L.check(L.load().drm_grad_norm(L.ptr(self.grad), self.grad.numel(), L.ptr(self._scratch), L.ptr(out), L.stream()), "grad_norm")
        return out[0]

    @property
    def last_grad_norm(self) -> torch.Tensor:
        return self.opt_state[1]

    @property
    def last_step_skipped(self) -> torch.Tensor:
        return self.opt_state[3] != 0

    def load_state_dict(self, state_dict):
        """Accepts torch.optim.AdamW-format state; values are copied into the flat buffers."""
        groups = state_dict["param_groups"]
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in groups[0]:
                self.param_groups[0][k] = tuple(groups[0][k]) if k == "betas" else groups[0][k]
        for k in ("max_norm", "tau"):
            if k in groups[0]:
                self.param_groups[0][k] = groups[0][k]
        ids = groups[0]["params"]
        with torch.no_grad():
            for pid, p in zip(ids, self._params):
                st = state_dict["state"].get(pid)
                if st is None:
                    continue
                self.state[p]["exp_avg"].copy_(st["exp_avg"])
                self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
                self.opt_state[0] = float(st["step"])


def make_adamw(params, lr, betas, eps, weight_decay=1e-6, max_norm=100.0, ema_params=None, tau=0.02):
    """FlatAdamW for parameters that live on the GPU.  Modules constructed on the CPU (state_dict / surface work only:
    no kernel accepts CPU tensors) get a plain torch.optim.AdamW so that construction itself never needs a device."""
    params = list(params)
    if params and params[0].is_cuda:
        return FlatAdamW(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_norm=max_norm, ema_params=ema_params, tau=tau)
    return torch.optim.AdamW(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
