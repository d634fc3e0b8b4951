"""The B = 1 acting path: the inner loop of Dreamer.rollout_policy / evaluate_agent
(Dreamer.py:177-226, 295-322) for ONE environment on the mirrored modules.

Per environment step the reference does  actor.act(h, z) -> env.step -> buffer.add_to_buffer(obs, a, r, c) ->
world_model.observe_step(z, h, a, obs')  as ~100 small launches plus three host<->device tensor constructions.  Here the
agent state (h, z, last action, current frame) lives in fixed device tensors, the frame travels through ONE pinned staging
buffer as uint8 (12 KB instead of 48 KB of fp32), the transition goes from those device tensors straight into the HBM replay
ring, and  observe_step + act  is one captured CUDA graph (graphs.StepGraph) whose only host interaction is the 12-byte
action read-back the environment needs.

    acting = ActingPath(world_model, agent, buffer)
    acting.reset(frame_u8_chw)                       # Dreamer.py:184-191: h = 0, z = encode(h, obs)
    a = acting.act()                                 # numpy (A,)
    while ...:
        frame, reward, done = env_step(a)
        a = acting.step(frame, reward, 1 - done)     # record (obs, a, r, c) -> observe_step(obs') -> act
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib as L
from .graphs import StepGraph


class ActingPath:
    def __init__(self, world_model, agent, buffer=None, deterministic: bool = False, use_graphs: bool = True, warmup: int = 3):
        self.wm, self.agent, self.buffer, self.deterministic = world_model, agent, buffer, deterministic
        p = next(world_model.parameters())
        L.require_cuda(p, "world model parameters")
        dev = p.device
        R, C, D, A = world_model.latent_num_rows, world_model.latent_num_columns, world_model.hidden_dims, world_model.action_dims
        H, W = world_model.observation_dim_x, world_model.observation_dim_y
        f = dict(dtype=torch.float32, device=dev)
        self.hidden = torch.zeros(1, 1, D, **f)
        self.latent = torch.zeros(1, 1, R, C, **f)
        self.action = torch.zeros(1, 1, A, **f)
        self.frame = torch.zeros(3, H, W, dtype=torch.uint8, device=dev)          # the observation (h, z) were computed from
        self.meta = torch.zeros(2, **f)                                           # [reward, continue] of the pending transition
        self._frame_host = torch.zeros(3, H, W, dtype=torch.uint8).pin_memory()
        self._meta_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        self._action_host = torch.zeros(A, dtype=torch.float32).pin_memory()
        self._action_in_host = torch.zeros(A, dtype=torch.float32).pin_memory()      # host -> device staging of set_action (not the read-back buffer)
        # one event per pinned staging buffer: recorded after its asynchronous copy is enqueued, waited for before the buffer is rewritten
        self._staged = {"frame": None, "meta": None, "action": None}
        lut = (((np.arange(256, dtype=np.uint8).astype(np.float32) / 255.0) - 0.5 + 0.5) * 255.0).astype(np.uint8)
        self._roundtrip_lut = torch.from_numpy(lut).to(dev)
        world_model.attach_actor(agent.actor)
        if use_graphs:
            self._reset = StepGraph(self._reset_body, warmup)
            self._observe_act = StepGraph(self._observe_act_body, warmup)
            self._observe = StepGraph(self._observe_body, warmup)
            self._act = StepGraph(self._act_body, warmup)
        else:
            self._reset, self._observe_act, self._observe, self._act = (self._reset_body, self._observe_act_body, self._observe_body,
                                                                        self._act_body)

    # ---- captured bodies: device tensors in, device tensors updated in place, no host synchronisation ----------------------
    def _obs(self):
        return (self.frame.to(torch.float32) / 255.0 - 0.5).view(1, 1, *self.frame.shape)

    def _reset_body(self, uniforms=None):
        self.hidden.zero_()
        z, _ = self.wm.encoder.encode(self.hidden, self._obs(), uniforms)
        self.latent.copy_(z.view_as(self.latent))
        return self.latent

    def _act_body(self, normals=None):
        a, _, _ = self.agent.actor.act(self.hidden, self.latent, deterministic=self.deterministic, normals=normals)
        self.action.copy_(a.view_as(self.action))
        return self.action

    def _observe_body(self, uniforms=None):
        z, h, _ = self.wm.observe_step(self.latent, self.hidden, self.action, self._obs(), uniforms)
        self.hidden.copy_(h.view_as(self.hidden))
        self.latent.copy_(z.view_as(self.latent))
        return self.latent

    def _observe_act_body(self, uniforms=None, normals=None):
        self._observe_body(uniforms)
        return self._act_body(normals)

    # ---- host API -----------------------------------------------------------------------------------------------------
    def sync_weights(self):
        """Re-pack the bf16 weight caches the captured graphs read if a parameter changed since the last call (the packed
        buffers are updated in place, so the graphs pick the new weights up).  Called by reset(); call it yourself after
        training when an episode continues across a training phase."""
        self.wm._engine.refresh()
        self.wm.sequence_model._pk.get()
        self.agent.actor._pk.get()

    def _stage(self, which, host, dev_view, value):
        """Write `value` into the pinned buffer `host` and enqueue its copy to the device -- after the PREVIOUS copy out of that
        buffer has executed (the GPU may still hold a backlog, e.g. a replayed training step)."""
        ev = self._staged[which]
        if ev is not None:
            ev.synchronize()
        host.copy_(value)
        dev_view.copy_(host, non_blocking=True)
        if ev is None:
            ev = self._staged[which] = torch.cuda.Event()
        ev.record()

    def _upload(self, frame_u8_chw):
        self._stage("frame", self._frame_host, self.frame, torch.as_tensor(np.ascontiguousarray(frame_u8_chw), dtype=torch.uint8))

    def _download_action(self) -> np.ndarray:
        self._action_host.copy_(self.action.view(-1), non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the one host wait of an environment step
        return self._action_host.numpy().copy()

    def reset(self, frame_u8_chw, uniforms: Optional[torch.Tensor] = None):
        """Start of an episode (Dreamer.py:184-191 / 218-222): h = 0, z ~ posterior(h, obs)."""
        self.sync_weights()
        self._upload(frame_u8_chw)
        self._reset(uniforms)

    def act(self, normals: Optional[torch.Tensor] = None) -> np.ndarray:
        """Agent.py:202-210 on the current state; returns the action as a host array."""
        self._act(normals)
        return self._download_action()

    def set_action(self, action):
        """Use an externally chosen action (Dreamer.rollout_policy's random_policy branch, Dreamer.py:195-198) as the last action."""
        self._stage("action", self._action_in_host, self.action.view(-1), torch.as_tensor(np.asarray(action, dtype=np.float32).reshape(-1)))

    def observe(self, next_frame_u8_chw, uniforms: Optional[torch.Tensor] = None):
        """observe_step on the new frame with the last action, without choosing the next one (WorldModel.py:79-82)."""
        self._upload(next_frame_u8_chw)
        self._observe(uniforms)

    def record(self, reward: float, continue_: float):
        """buffer.add_to_buffer(current frame, last action, reward, continue) without leaving the device (Dreamer.py:211-212)."""
        if self.buffer is None:
            return
        self._stage("meta", self._meta_host, self.meta, torch.tensor([float(reward), float(continue_)], dtype=torch.float32))
        # Dreamer.rollout_policy keeps the current frame normalised and stores ((obs / 255 - 0.5) + 0.5) * 255 truncated to uint8
        # (Dreamer.py:186, 209): in fp32 that round trip lands just below the integer for 63 of the 256 pixel values, so the reference's
        # ring holds x - 1 there.  Reproduced bit for bit (a 12 KB table look-up per step) so that the ring contents are identical.
        # The 256-entry table is evaluated once on the host in numpy fp32, exactly as the reference evaluates it; the device only gathers.
        stored = self._roundtrip_lut[self.frame.long()]
        self.buffer.add_batch(stored[None], self.action.view(1, -1), self.meta[0:1], self.meta[1:2])

    def step(self, next_frame_u8_chw, reward: float, continue_: float, uniforms: Optional[torch.Tensor] = None,
             normals: Optional[torch.Tensor] = None) -> np.ndarray:
        """One environment step after the first act(): record the transition, absorb the new frame (observe_step,
        WorldModel.py:79-82) and choose the next action -- one graph replay, one 12-byte read-back."""
        self.record(reward, continue_)
        self._upload(next_frame_u8_chw)
        self._observe_act(uniforms, normals)
        return self._download_action()
