"""Whole-rollout entry points: the fused replacement of Dreamer.dream_episodes (Dreamer.py:143-175).

``dream_episodes`` takes device tensors; ``dream_episodes_host`` is the host-buffer variant (pinned
host inputs are copied H2D on the current stream, the per-step rewards / continue probabilities are
read back D2H) that bench.py times as the end-to-end number.
"""
from __future__ import annotations

import torch

from . import ops


def dream_episodes(rollout: ops.Rollout, z0, h0, uniforms=None, normals=None, generator=None, want_idx=False):
    """Returns (latent (B,H+1,R,C), hidden (B,H+1,D), actions, rewards, continues, mu, sigma) exactly like
    the reference.  ``uniforms`` (H,B,R) / ``normals`` (H,B,A) default to fresh device-side draws."""
    m, B, H = rollout.model, rollout.B, rollout.H
    dev = z0.device
    if uniforms is None:
        uniforms = torch.rand((H, B, m.R), device=dev, generator=generator)
    if normals is None:
        normals = torch.randn((H, B, m.A), device=dev, generator=generator)
    return rollout.run(z0, h0, uniforms, normals, want_idx=want_idx)


def dream_episodes_host(rollout: ops.Rollout, z0, h0, uniforms=None, normals=None, generator=None):
    """Host-buffer call: the start states are (pinned) CPU tensors; the draws are made on the device exactly as
    ``Dreamer.dream_episodes`` does (pass ``uniforms`` / ``normals`` host tensors to supply them instead).
    Returns dict(device=7-tuple, host=(rewards, continues))."""
    dev = torch.device("cuda", torch.cuda.current_device())
    m, B, H = rollout.model, rollout.B, rollout.H
    zd, hd = z0.to(dev, non_blocking=True), h0.to(dev, non_blocking=True)
    ud = uniforms.to(dev, non_blocking=True) if uniforms is not None else torch.rand((H, B, m.R), device=dev, generator=generator)
    nd = normals.to(dev, non_blocking=True) if normals is not None else torch.randn((H, B, m.A), device=dev, generator=generator)
    out = rollout.run(zd, hd, ud, nd, want_idx=False)
    host = [out[3].to("cpu", non_blocking=True), out[4].to("cpu", non_blocking=True)]
    torch.cuda.current_stream().synchronize()
    return dict(device=out, host=host)


def dream_episodes_modules(world_model, agent, starting_latent_state_batch, starting_hidden_state_batch, horizon=None,
                           uniforms=None, normals=None, generator=None):
    """Drop-in body for ``Dreamer.dream_episodes`` (Dreamer.py:143-175) on the mirrored modules: one fused rollout instead
    of ``horizon`` x (Actor.act -> WorldModel.imagine_step).  Returns the reference's 7-tuple."""
    H = world_model.horizon if horizon is None else horizon
    B = starting_hidden_state_batch.shape[0]
    world_model.attach_actor(agent.actor)
    ro = world_model._engine.rollout(B, H)
    return dream_episodes(ro, starting_latent_state_batch, starting_hidden_state_batch, uniforms, normals, generator)
