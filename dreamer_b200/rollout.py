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


def dream_episodes_host(rollout: ops.Rollout, z0, h0, uniforms, normals):
    """Host-buffer call: inputs are (pinned) CPU tensors; returns dict(device=7-tuple, host=(rewards, continues))."""
    dev = torch.device("cuda", torch.cuda.current_device())
    d = [t.to(dev, non_blocking=True) for t in (z0, h0, uniforms, normals)]
    out = rollout.run(*d, want_idx=False)
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
