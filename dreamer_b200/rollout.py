"""Whole-rollout entry points: the fused replacement of Dreamer.dream_episodes (Dreamer.py:143-175).

``dream_episodes`` takes device tensors; ``dream_episodes_host`` is the host-buffer variant (pinned
host inputs are copied H2D on the current stream, the per-step rewards / continue probabilities are
read back D2H) that bench.py times as the end-to-end number.
"""
from __future__ import annotations

import torch

from . import ops


def dream_episodes(rollout: ops.Rollout, z0, h0, uniforms=None, normals=None, generator=None, want_idx=False, graphed=False):
    """Returns (latent (B,H+1,R,C), hidden (B,H+1,D), actions, rewards, continues, mu, sigma) exactly like
    the reference.  ``uniforms`` (H,B,R) / ``normals`` (H,B,A) default to fresh device-side draws.
    ``graphed``: replay the rollout as one CUDA graph (ops.Rollout.run_graphed); the returned tensors are then the graph's
    static outputs and are overwritten by the next call on this workspace."""
    m, B, H = rollout.model, rollout.B, rollout.H
    dev = z0.device
    if uniforms is None:
        uniforms = torch.rand((H, B, m.R), device=dev, generator=generator)
    if normals is None:
        normals = torch.randn((H, B, m.A), device=dev, generator=generator)
    run = rollout.run_graphed if graphed else rollout.run
    return run(z0, h0, uniforms, normals, want_idx=want_idx)


def dream_episodes_host(rollout: ops.Rollout, z0, h0, uniforms=None, normals=None, generator=None):
    """Host-buffer call: the start states are (pinned) CPU tensors; the draws are made on the device exactly as
    ``Dreamer.dream_episodes`` does (pass ``uniforms`` / ``normals`` host tensors to supply them instead).
    ``z0`` is either the fp32 latent (B, 1, R, 32) -- 4 KB per state over the host link -- or, since a sampled latent is a
    one-hot, its uint8 class indices (B, R): 32 bytes per state, expanded on the device (drm_onehot32).
    Returns dict(device=7-tuple, host=(rewards, continues)); the host tensors are pinned buffers reused by the next call.

    The workspace keeps static device inputs and, after two eager calls, replays the rollout's ~110 launches as ONE CUDA graph
    (graphs.StepGraph): the host then issues two copies, two RNG fills, one graph launch and two read-backs per call instead of
    half a millisecond of kernel launches, so the call's wall time is the device time plus the host link."""
    dev = torch.device("cuda", torch.cuda.current_device())
    m, B, H = rollout.model, rollout.B, rollout.H
    st = rollout.__dict__.get("_host_state")
    if st is None:
        f = dict(dtype=torch.float32, device=dev)
        st = dict(z=torch.empty((B, 1, m.R, m.C), **f), h=torch.empty((B,) + tuple(h0.shape[1:]), **f),
                  u=torch.empty((H, B, m.R), **f), n=torch.empty((H, B, m.A), **f), host=None,
                  zi=torch.empty((B, m.R), dtype=torch.uint8, device=dev))
        rollout.__dict__["_host_state"] = st
    if z0.dtype == torch.uint8:
        st["zi"].copy_(z0.reshape(B, m.R), non_blocking=True)
        ops.onehot32(st["zi"], st["z"])
    else:
        st["z"].copy_(z0.reshape(st["z"].shape), non_blocking=True)
    st["h"].copy_(h0, non_blocking=True)
    if uniforms is not None:
        st["u"].copy_(uniforms, non_blocking=True)
    else:
        st["u"].uniform_(generator=generator)
    if normals is not None:
        st["n"].copy_(normals, non_blocking=True)
    else:
        st["n"].normal_(generator=generator)
    out = rollout.run_graphed(st["z"], st["h"], st["u"], st["n"], want_idx=False)
    # results come back into PINNED host buffers owned by the workspace (a pageable destination would be staged and synchronous)
    if st["host"] is None:
        st["host"] = [torch.empty(out[3].shape, dtype=out[3].dtype).pin_memory(), torch.empty(out[4].shape, dtype=out[4].dtype).pin_memory()]
    st["host"][0].copy_(out[3], non_blocking=True)
    st["host"][1].copy_(out[4], non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return dict(device=out, host=st["host"])


class HostRolloutQueue:
    """Two-deep submit / result queue around ``dream_episodes_host``: the host -> device copy of call i + 1's start states runs on a
    copy stream WHILE call i's rollout computes, so the host link (2.5 MB of start states per 1024 x 15 rollout, ~0.1 - 0.2 ms)
    leaves the critical path of a stream of rollouts.  Every call still moves its own inputs from pinned host memory and its own
    results (rewards, continues) back into pinned host memory.

        q = HostRolloutQueue(rollout)
        t0 = q.submit(z0_idx_u8_pinned, h0_pinned)      # returns at once
        t1 = q.submit(...)                               # copies overlap the first rollout
        rewards, continues = q.result(t0)                # pinned buffers of that call, valid until its slot is submitted to again

    Results of the device-side 7-tuple are the graph's static outputs and are overwritten by the next rollout (as with
    ``Rollout.run_graphed``); only the host read-back is per call."""

    def __init__(self, rollout: ops.Rollout, depth: int = 2, generator=None):
        self.ro, self.depth, self.generator = rollout, int(depth), generator
        m, B, H = rollout.model, rollout.B, rollout.H
        dev = torch.device("cuda", torch.cuda.current_device())
        f = dict(dtype=torch.float32, device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.zi = [torch.empty((B, m.R), dtype=torch.uint8, device=dev) for _ in range(self.depth)]
        self.h = [torch.empty((B, m.D), **f) for _ in range(self.depth)]
        self.z = torch.empty((B, 1, m.R, m.C), **f)
        self.u = torch.empty((H, B, m.R), **f)
        self.n = torch.empty((H, B, m.A), **f)
        self.host = [None] * self.depth
        self.in_done = [torch.cuda.Event() for _ in range(self.depth)]
        self.read_done = [None] * self.depth      # the compute stream has consumed slot k's staging buffers
        self.out_done = [None] * self.depth
        self.count = 0

    def _graph_inputs(self):
        """(z, uniforms, normals) to fill for the next call: the captured graph's static inputs when it exists, else own buffers"""
        g = self.ro.__dict__.get("_graphs", {}).get(False)
        if g is not None:
            for ent in g.cache.values():
                st = ent.get("static")
                if st is not None and st[0].shape == self.z.shape and st[2].shape == self.u.shape and st[3].shape == self.n.shape:
                    return st[0], st[2], st[3]
        return self.z, self.u, self.n

    def submit(self, z0_idx: torch.Tensor, h0: torch.Tensor) -> int:
        """z0_idx (B, R) uint8 classes of the one-hot start latent, h0 (B, 1, D) / (B, D) fp32 -- pinned host tensors."""
        k = self.count % self.depth
        main = torch.cuda.current_stream()
        if self.read_done[k] is not None:
            self.copy_stream.wait_event(self.read_done[k])
        with torch.cuda.stream(self.copy_stream):
            self.zi[k].copy_(z0_idx.reshape(self.zi[k].shape), non_blocking=True)
            self.h[k].copy_(h0.reshape(self.h[k].shape), non_blocking=True)
            self.in_done[k].record(self.copy_stream)
        main.wait_event(self.in_done[k])
        # once the rollout graph exists, the one-hot start latent and the draws are written straight into ITS static input buffers
        # (run_graphed then has nothing to copy but h0): three device-to-device copies (6 MB) less per call
        z_t, u_t, n_t = self._graph_inputs()
        ops.onehot32(self.zi[k], z_t)
        u_t.uniform_(generator=self.generator)
        n_t.normal_(generator=self.generator)
        out = self.ro.run_graphed(z_t, self.h[k], u_t, n_t, want_idx=False)
        self.read_done[k] = torch.cuda.Event()
        self.read_done[k].record(main)
        if self.host[k] is None:
            self.host[k] = [torch.empty(out[3].shape, dtype=out[3].dtype).pin_memory(), torch.empty(out[4].shape, dtype=out[4].dtype).pin_memory()]
        self.host[k][0].copy_(out[3], non_blocking=True)
        self.host[k][1].copy_(out[4], non_blocking=True)
        self.out_done[k] = torch.cuda.Event()
        self.out_done[k].record(main)
        self.device_out = out
        t = self.count
        self.count += 1
        return t

    def result(self, ticket: int):
        """Blocks until call `ticket` has finished; -> (rewards, continues) pinned host tensors of that call."""
        if ticket < self.count - self.depth or ticket >= self.count:
            raise RuntimeError("HostRolloutQueue.result: that call's slot has been reused (or it was never submitted)")
        k = ticket % self.depth
        self.out_done[k].synchronize()
        return self.host[k]


def dream_episodes_modules(world_model, agent, starting_latent_state_batch, starting_hidden_state_batch, horizon=None,
                           uniforms=None, normals=None, generator=None, graphed=False):
    """Drop-in body for ``Dreamer.dream_episodes`` (Dreamer.py:143-175) on the mirrored modules: one fused rollout instead
    of ``horizon`` x (Actor.act -> WorldModel.imagine_step).  Returns the reference's 7-tuple."""
    H = world_model.horizon if horizon is None else horizon
    B = starting_hidden_state_batch.shape[0]
    world_model.attach_actor(agent.actor)
    ro = world_model._engine.rollout(B, H)
    return dream_episodes(ro, starting_latent_state_batch, starting_hidden_state_batch, uniforms, normals, generator, graphed=graphed)
