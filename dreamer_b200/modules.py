"""Mirrors of the reference's Python module surface on the sm_100a kernels.

Same class names, constructor argument order, method names, return tuples and ``state_dict`` keys as
SequenceModel.py, DynamicsPredictors.py, VariationalAutoEncoder.py, Agent.py (Actor / Critic),
WorldModel.py, Agent.py and Buffer.py of the reference, so ``Dreamer.py`` / ``train_car_racer.py`` can
import them unchanged (``dreamer_b200.dropin.install()`` registers them under the reference's module
names).  ``torch.nn`` layers are used only as parameter containers (identical key names and default
initialisation); every forward pass runs through ``libdreamer_b200.so``.

Random draws: the reference samples from the global torch RNG.  Here every sampling method takes an
optional ``uniforms`` / ``normals`` tensor (the C-ABI's host-supplied randomness); when omitted they
are drawn on the device with ``torch.rand`` / ``torch.randn``.

Forward outputs carry no autograd graph.  ``WorldModel.training_step`` and ``Agent.train_step``
compute their reported losses on the kernels and obtain gradients from ``_autograd_tail`` (a
device-side torch graph, teacher-forced on the kernels' sampled indices) -- the BPTT kernels that
replace it are the first "next" row of SURVEY.md section 8f.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops


# --------------------------------------------------------------------------------------------
# DreamerUtils.py:29-50 (elementwise helpers kept as tensor expressions; the fused forms live in the kernels)
# --------------------------------------------------------------------------------------------
def symlog(x):
    return torch.sign(x) * torch.log(1.0 + torch.abs(x))


def symlog_np(x):
    return np.sign(x) * np.log(1.0 + np.abs(x))


def symexp(x):
    x = torch.clamp(x, -20.0, 20.0)
    return torch.sign(x) * (torch.exp(torch.abs(x).float()) - 1.0)


def to_twohot(value: torch.Tensor, buckets: torch.Tensor):
    """DreamerUtils.py:39-50 (the losses never materialise this; kept for API compatibility)."""
    v = torch.clamp(value, min=buckets.min(), max=buckets.max())
    idx = torch.clamp(torch.searchsorted(buckets, v.contiguous(), right=True) - 1, max=len(buckets) - 2)
    w = (v - buckets[idx]) / (buckets[idx + 1] - buckets[idx] + 1e-8)
    out = torch.zeros(value.shape[:-1] + (buckets.shape[0],), dtype=torch.float32, device=value.device)
    out = torch.scatter(out, -1, idx, 1.0 - w)
    return torch.scatter(out, -1, idx + 1, w)


def _mlp(d_in, h1, h2, d_out, device):
    layers = [nn.Linear(d_in, h1, device=device), nn.LayerNorm(h1, device=device), nn.SiLU(),
              nn.Linear(h1, h2, device=device), nn.LayerNorm(h2, device=device), nn.SiLU()]
    if d_out is not None:
        layers.append(nn.Linear(h2, d_out, device=device))
    return nn.Sequential(*layers)


class _Packed:
    """Cache of the packed-weight handle and row workspaces of one module (re-packed when a parameter changes)."""

    def __init__(self, owner: nn.Module, prefix: str, R: int = 32, C: int = 32, D: Optional[int] = None, A: int = 3):
        self.owner, self.prefix, self.R, self.C, self.D, self.A = owner, prefix, R, C, D, A
        self.model: Optional[ops.PackedRssm] = None
        self.versions = None
        self.ws: Dict[int, ops.Rollout] = {}

    def __deepcopy__(self, memo):
        """copy.deepcopy(module) (snapshots, Agent's target critic) must not share native handles: the copy starts with an empty
        cache bound to the COPIED owner and packs its own weights on first use."""
        import copy as _copy
        return _Packed(_copy.deepcopy(self.owner, memo), self.prefix, self.R, self.C, self.D, self.A)

    def __getstate__(self):
        return dict(owner=self.owner, prefix=self.prefix, R=self.R, C=self.C, D=self.D, A=self.A, model=None, versions=None, ws={})

    def sd(self):
        return {self.prefix + k: v for k, v in self.owner.state_dict(keep_vars=True).items()}

    def get(self) -> ops.PackedRssm:
        sd = self.sd()
        vers = tuple((v.data_ptr(), v._version) for v in sd.values())
        if self.model is None:
            for k, v in sd.items():
                L.require_cuda(v, k)
            self.model = ops.PackedRssm.from_state_dict(sd, self.R, self.C, D=self.D, A=self.A)
            self.versions = vers
        elif vers != self.versions:
            self.model.pack(sd)
            self.versions = vers
        return self.model

    def invalidate(self):
        """Force the next use to re-pack (a captured CUDA graph must contain the pack kernels)."""
        self.versions = None

    def workspace(self, n_rows: int) -> ops.Rollout:
        model = self.get()
        cap = max(128, 1 << (max(n_rows, 1) - 1).bit_length())
        if cap not in self.ws:
            self.ws[cap] = ops.Rollout(model, cap, 1)
        return self.ws[cap]


def _flat2(x):
    """(B, S, ...) -> (B*S, ...)"""
    return x.reshape((x.shape[0] * x.shape[1],) + tuple(x.shape[2:]))


# --------------------------------------------------------------------------------------------
# SequenceModel.py
# --------------------------------------------------------------------------------------------
class SequenceModel(nn.Module):
    def __init__(self, latent_num_rows, latent_num_columns, hidden_dim, action_dim, *, num_layers=1, device='cpu'):
        super().__init__()
        self.latent_dim = latent_num_columns * latent_num_rows
        self.hidden_dim = hidden_dim
        self.num_layers = num_layers
        self.device = device
        self.flatten = nn.Flatten(start_dim=2)
        self.GRU = nn.GRUCell(input_size=self.latent_dim + action_dim, hidden_size=hidden_dim, device=device)
        self._pk = _Packed(self, "world_model.sequence_model.", latent_num_rows, latent_num_columns, hidden_dim, action_dim)

    def forward(self, last_latent_state, last_hidden_state, last_action):
        """(B,1,R,C), (B,1,D), (B,1,A) -> (B,1,D)   [SequenceModel.py:19-24]"""
        B = last_hidden_state.shape[0]
        ws = self._pk.workspace(B)
        h = ws.gru_step(last_latent_state.reshape(B, -1), last_hidden_state.reshape(B, -1), last_action.reshape(B, -1))
        return h.unsqueeze(1)


# --------------------------------------------------------------------------------------------
# DynamicsPredictors.py
# --------------------------------------------------------------------------------------------
class DynamicsPredictor(nn.Module):
    def __init__(self, latent_num_rows, latent_num_columns, hidden_state_size, hidden_L1, hidden_L2, device):
        super().__init__()
        self.latent_num_rows = latent_num_rows
        self.latent_num_columns = latent_num_columns
        self.latent_size = latent_num_rows * latent_num_columns
        self.device = device
        self.logit_net = _mlp(hidden_state_size, hidden_L1, hidden_L2, self.latent_size, device)
        self._pk = _Packed(self, "world_model.dynamics_predictor.", latent_num_rows, latent_num_columns, hidden_state_size)

    def forward(self, x):
        B, S, _ = x.shape
        ws = self._pk.workspace(B * S)
        return ws.prior(x.reshape(B * S, -1))["logits"].view(B, S, self.latent_num_rows, self.latent_num_columns)

    def predict(self, hidden_state, uniforms=None):
        """-> (straight-through latent (B,S,R,C), logits)   [DynamicsPredictors.py:31-40]"""
        B, S, _ = hidden_state.shape
        if uniforms is None:
            uniforms = torch.rand(B * S, self.latent_num_rows, device=hidden_state.device)
        ws = self._pk.workspace(B * S)
        out = ws.prior(hidden_state.reshape(B * S, -1), uniforms.reshape(B * S, -1))
        shp = (B, S, self.latent_num_rows, self.latent_num_columns)
        return out["z"].view(shp), out["logits"].view(shp)


class RewardPredictor(nn.Module):
    def __init__(self, latent_num_rows, latent_num_columns, hidden_state_size, hidden_L1, hidden_L2, num_buckets=255, device='cpu'):
        super().__init__()
        self.latent_size = latent_num_rows * latent_num_columns
        self.buckets = num_buckets
        self.device = device
        self.flatten = nn.Flatten(start_dim=2)
        self.logit_net = _mlp(hidden_state_size + self.latent_size, hidden_L1, hidden_L2, num_buckets, device)
        self.register_buffer('buckets_rew', torch.linspace(-20.0, 20.0, num_buckets, device=device))
        self._pk = _Packed(self, "world_model.reward_predictor.", latent_num_rows, latent_num_columns, hidden_state_size)

    def _run(self, hidden, latent, want_logits):
        B, S, _ = hidden.shape
        ws = self._pk.workspace(B * S)
        return ws.heads(hidden.reshape(B * S, -1), latent.reshape(B * S, -1), L.HEAD_REWARD, want_logits=want_logits), B, S

    def forward(self, hidden, latent):
        out, B, S = self._run(hidden, latent, True)
        return out["reward_logits"].view(B, S, -1)

    def predict(self, hidden_state, latent_state):
        out, B, S = self._run(hidden_state, latent_state, False)
        return out["reward"].view(B, S, 1)


class ContinuePredictor(nn.Module):
    def __init__(self, latent_num_rows, latent_num_columns, hidden_state_size, hidden_L1, hidden_L2, device):
        super().__init__()
        self.latent_size = latent_num_rows * latent_num_columns
        self.device = device
        self.flatten = nn.Flatten(start_dim=2)
        self.logit_generator = _mlp(hidden_state_size + self.latent_size, hidden_L1, hidden_L2, 1, device)
        self._pk = _Packed(self, "world_model.continue_predictor.", latent_num_rows, latent_num_columns, hidden_state_size)

    def forward(self, hidden, latent):
        B, S, _ = hidden.shape
        ws = self._pk.workspace(B * S)
        out = ws.heads(hidden.reshape(B * S, -1), latent.reshape(B * S, -1), L.HEAD_CONT)
        return out["cont_prob"].view(B, S, 1), out["cont_logit"].view(B, S, 1)

    def predict(self, hidden_state, latent_state):
        return self.forward(hidden_state, latent_state)[0]


# --------------------------------------------------------------------------------------------
# Agent.py: Actor / Critic
# --------------------------------------------------------------------------------------------
class Actor(nn.Module):
    def __init__(self, action_dim, latent_column_dim, latent_row_dim, hidden_state_dim, hidden_layer_num_nodes_1,
                 hidden_layer_num_nodes_2, *, device='cpu'):
        super().__init__()
        self.flatten = nn.Flatten(start_dim=2)
        self.base_net = _mlp(latent_row_dim * latent_column_dim + hidden_state_dim, hidden_layer_num_nodes_1, hidden_layer_num_nodes_2, None, device)
        self.mu_head = nn.Linear(hidden_layer_num_nodes_2, action_dim, device=device)
        self.log_sig_head = nn.Linear(hidden_layer_num_nodes_2, action_dim, device=device)
        torch.nn.init.zeros_(self.mu_head.weight)
        torch.nn.init.zeros_(self.mu_head.bias)
        self.action_dim = action_dim
        # (the reference's constructor names are swapped, Agent.py:175: `latent_column_dim` receives latent_dims[0] = the row count)
        self._pk = _Packed(self, "agent.actor.", latent_column_dim, latent_row_dim, hidden_state_dim, action_dim)

    def _run(self, ht, zt, normals):
        B, S, _ = ht.shape
        ws = self._pk.workspace(B * S)
        out = ws.heads(ht.reshape(B * S, -1), zt.reshape(B * S, -1), L.HEAD_ACTOR, normals=normals)
        return out, B, S

    def forward(self, ht, zt):
        out, B, S = self._run(ht, zt, None)
        return out["mu"].view(B, S, -1), out["sigma"].view(B, S, -1)

    def act(self, ht, zt, deterministic=False, normals=None):
        """-> (action, mu, sigma)   [Agent.py:202-210]; tanh-Normal rsample == tanh(mu + sigma * eps)"""
        B, S, _ = ht.shape
        if deterministic:
            mu, sigma = self.forward(ht, zt)
            return torch.tanh(mu), mu, sigma
        if normals is None:
            normals = torch.randn(B * S, self.action_dim, device=ht.device)
        out, B, S = self._run(ht, zt, normals.reshape(B * S, -1))
        return out["action"].view(B, S, -1), out["mu"].view(B, S, -1), out["sigma"].view(B, S, -1)


class Critic(nn.Module):
    def __init__(self, latent_row_dim, latent_column_dim, hidden_state_dim, hidden_layer_num_nodes_1, hidden_layer_num_nodes_2,
                 num_buckets=255, device='cpu'):
        super().__init__()
        self.latent_row_dim = latent_row_dim
        self.latent_column_dim = latent_column_dim
        self.num_buckets = num_buckets
        self.flatten = nn.Flatten(start_dim=2)
        self.value_net = _mlp(latent_column_dim * latent_row_dim + hidden_state_dim, hidden_layer_num_nodes_1, hidden_layer_num_nodes_2, num_buckets, device)
        self.register_buffer('buckets_crit', torch.linspace(-20, 20, num_buckets, device=device))
        self._pk = _Packed(self, "agent.critic.", latent_row_dim, latent_column_dim, hidden_state_dim)

    def _run(self, ht, zt, want_logits):
        B, S, _ = ht.shape
        ws = self._pk.workspace(B * S)
        return ws.heads(ht.reshape(B * S, -1), zt.reshape(B * S, -1), L.HEAD_CRITIC, want_logits=want_logits), B, S

    def forward(self, ht, zt):
        out, B, S = self._run(ht, zt, True)
        return out["value_logits"].view(B, S, -1)

    def value(self, ht, zt):
        out, B, S = self._run(ht, zt, False)
        return out["value"].view(B, S, 1)


# --------------------------------------------------------------------------------------------
# VariationalAutoEncoder.py
# --------------------------------------------------------------------------------------------
class _VaeMixin:
    """Encoder and Decoder each pack the shared drm_vae handle lazily from whatever half they own; the
    WorldModel (which owns both) installs the combined handle via ``_bind``."""
    _vae_owner = None

    def _bind(self, owner):
        object.__setattr__(self, "_vae_owner", owner)


class Encoder(nn.Module, _VaeMixin):
    def __init__(self, observation_dims, hidden_state_dim, latent_num_rows, latent_num_columns, num_filters_1, num_filters_2,
                 hidden_layer_nodes, device='cpu'):
        super().__init__()
        self.latent_size = latent_num_rows * latent_num_columns
        self.latent_num_rows = latent_num_rows
        self.latent_num_columns = latent_num_columns
        self.final_height = observation_dims[0] // 16
        self.final_width = observation_dims[1] // 16
        if self.final_height < 1 or self.final_width < 1:
            raise ValueError(f"Input image {observation_dims} is too small for 4 layers of downsampling.")
        f1, f2 = num_filters_1, num_filters_2
        self.feature_extractor = nn.Sequential(
            nn.Conv2d(3, f1, kernel_size=4, stride=2, padding=1, device=device), nn.SiLU(),
            nn.Conv2d(f1, f2, kernel_size=4, stride=2, padding=1, device=device), nn.SiLU(),
            nn.Conv2d(f2, f2 * 2, kernel_size=4, stride=2, padding=1, device=device), nn.SiLU(),
            nn.Conv2d(f2 * 2, f2 * 4, kernel_size=4, stride=2, padding=1, device=device), nn.SiLU())
        total_in = f2 * 4 * self.final_height * self.final_width + hidden_state_dim
        self.flatten = nn.Flatten(start_dim=2)
        self.latent_mapper = nn.Sequential(nn.Linear(total_in, hidden_layer_nodes, device=device), nn.LayerNorm(hidden_layer_nodes, device=device),
                                           nn.SiLU(), nn.Linear(hidden_layer_nodes, self.latent_size, device=device))
        self.observation_dims = tuple(observation_dims)
        self.hidden_state_dim = hidden_state_dim

    def forward(self, hidden, observation):
        """-> flat logits (B, S, R*C)   [VariationalAutoEncoder.py:57-75]"""
        B, S = hidden.shape[:2]
        out = _vae_engine(self).observe(B * S).encode(_flat2(hidden), _flat2(observation))
        return out["logits"].view(B, S, self.latent_size)

    def encode(self, hidden_state, observation, uniforms=None):
        """-> (straight-through latent (B,S,R,C), logits (B,S,R,C))   [VariationalAutoEncoder.py:77-99]"""
        B, S = hidden_state.shape[:2]
        if uniforms is None:
            uniforms = torch.rand(B * S, self.latent_num_rows, device=hidden_state.device)
        out = _vae_engine(self).observe(B * S).encode(_flat2(hidden_state), _flat2(observation), uniforms.reshape(B * S, -1))
        shp = (B, S, self.latent_num_rows, self.latent_num_columns)
        return out["z"].view(shp), out["logits"].view(shp)


class Decoder(nn.Module, _VaeMixin):
    def __init__(self, latent_num_rows, latent_num_columns, observation_dim, hidden_state_dim, num_filters_1, num_filters_2,
                 hidden_layer_nodes, device='cpu'):
        super().__init__()
        self.start_height = observation_dim[0] // 16
        self.start_width = observation_dim[1] // 16
        self.num_filters_start = num_filters_2 * 4
        self.hidden_dim = hidden_state_dim
        self.latent_row_dim = latent_num_rows
        self.latent_col_dim = latent_num_columns
        self.flatten = nn.Flatten(start_dim=1)
        f1, f2 = num_filters_1, num_filters_2
        self.upscaler = nn.Sequential(
            nn.Linear(latent_num_rows * latent_num_columns + hidden_state_dim, hidden_layer_nodes, device=device),
            nn.LayerNorm(hidden_layer_nodes, device=device), nn.SiLU(),
            nn.Linear(hidden_layer_nodes, self.num_filters_start * self.start_height * self.start_width, device=device), nn.SiLU())
        self.image_builder = nn.Sequential(
            nn.ConvTranspose2d(self.num_filters_start, f2 * 2, kernel_size=4, stride=2, padding=1, device=device), nn.SiLU(),
            nn.ConvTranspose2d(f2 * 2, f2, kernel_size=4, stride=2, padding=1, device=device), nn.SiLU(),
            nn.ConvTranspose2d(f2, f1, kernel_size=4, stride=2, padding=1, device=device), nn.SiLU(),
            nn.ConvTranspose2d(f1, 3, kernel_size=4, stride=2, padding=1, device=device), nn.Tanh())
        self.observation_dims = tuple(observation_dim)

    def forward(self, hidden, latent):
        """-> mu (B, S, 3, H, W)   [VariationalAutoEncoder.py:139-161]"""
        B, S, _ = hidden.shape
        mu = _vae_engine(self).observe(B * S).decode(_flat2(hidden), latent.reshape(B * S, -1))
        return mu.view(B, S, 3, *self.observation_dims)

    def decode(self, hidden_state, latent_state):
        return self.forward(hidden_state, latent_state)


class _VaeEngine:
    """Packed drm_rssm + drm_vae + observe workspaces for an (encoder, decoder, [world model]) group."""

    def __init__(self, get_sd, R, C, D, A, obs_hw):
        self.get_sd, self.R, self.C, self.D, self.A, self.obs_hw = get_sd, R, C, D, A, obs_hw
        self.model = self.vae = None
        self.versions = None
        self.obs_ws: Dict[tuple, ops.Observe] = {}
        self.roll_ws: Dict[tuple, ops.Rollout] = {}

    def __deepcopy__(self, memo):
        raise RuntimeError("_VaeEngine holds native handles bound to one WorldModel's parameters: deep-copy the WorldModel's "
                           "state_dict into a new WorldModel instead")

    def refresh(self):
        sd = self.get_sd()
        vers = tuple((v.data_ptr(), v._version) for v in sd.values())
        if self.model is None:
            for k, v in sd.items():
                L.require_cuda(v, k)
            self.model = ops.PackedRssm.from_state_dict(sd, self.R, self.C, D=self.D, A=self.A)
            self.vae = ops.PackedVae.from_state_dict(self.model, sd, self.obs_hw)
        elif vers != self.versions:
            self.model.pack(sd)
            self.vae.pack(sd)
        self.versions = vers
        return self

    def observe(self, B, T=1) -> ops.Observe:
        self.refresh()
        if T == 1:
            B = max(128, 1 << (max(B, 1) - 1).bit_length())
        if (B, T) not in self.obs_ws:
            self.obs_ws[(B, T)] = ops.Observe(self.vae, B, T)
        return self.obs_ws[(B, T)]

    def rollout(self, B, H) -> ops.Rollout:
        self.refresh()
        if (B, H) not in self.roll_ws:
            self.roll_ws[(B, H)] = ops.Rollout(self.model, B, H)
        return self.roll_ws[(B, H)]


def _vae_engine(mod) -> _VaeEngine:
    owner = mod._vae_owner
    if owner is not None:
        return owner._engine
    eng = mod.__dict__.get("_own_engine")
    if eng is None:
        # stand-alone Encoder / Decoder: the missing half is packed from zero tensors of the mirrored shapes
        dev = next(mod.parameters()).device
        if isinstance(mod, Encoder):
            D, R, C = mod.hidden_state_dim, mod.latent_num_rows, mod.latent_num_columns
            f1 = mod.feature_extractor[0].weight.shape[0]; f2 = mod.feature_extractor[2].weight.shape[0]
            hn = mod.latent_mapper[0].weight.shape[0]
            other = Decoder(R, C, mod.observation_dims, D, f1, f2, hn, device=dev)
            get = lambda: {**{"world_model.encoder." + k: v for k, v in mod.state_dict(keep_vars=True).items()},
                           **{"world_model.decoder." + k: v for k, v in other.state_dict(keep_vars=True).items()}}
        else:
            D, R, C = mod.hidden_dim, mod.latent_row_dim, mod.latent_col_dim
            f2 = mod.image_builder[2].weight.shape[1]; f1 = mod.image_builder[4].weight.shape[1]
            hn = mod.upscaler[0].weight.shape[0]
            other = Encoder(mod.observation_dims, D, R, C, f1, f2, hn, device=dev)
            get = lambda: {**{"world_model.decoder." + k: v for k, v in mod.state_dict(keep_vars=True).items()},
                           **{"world_model.encoder." + k: v for k, v in other.state_dict(keep_vars=True).items()}}
        eng = _VaeEngine(get, R, C, D, 3, mod.observation_dims)
        mod.__dict__["_own_engine"] = eng
        mod.__dict__["_other_half"] = other
    return eng


# --------------------------------------------------------------------------------------------
# Buffer.py -- HBM-resident replay ring
# --------------------------------------------------------------------------------------------
class Buffer:
    """Buffer.py:5-63 with the ring held in device memory: obs uint8 (cap,3,H,W), act f32 (cap,A), reward (symlog'd) and
    continue f32 (cap,1).  ``sample_sequences`` draws the start indices exactly like the reference (global ``np.random``,
    one re-draw for windows that straddle the write head) and gathers on the device."""

    def __init__(self, buffer_size, sequence_length, action_size, observation_dims, device='cpu'):
        dev = torch.device(device)
        # the ring lives in HBM; a CPU device is accepted only so that the reference's Dreamer(config, "cpu") constructor can be
        # exercised without a GPU -- every data call below then fails loudly in the C-ABI wrappers (there is no CPU path)
        self.observation_buffer = torch.zeros((buffer_size, 3, *observation_dims), dtype=torch.uint8, device=dev)
        self.action_buffer = torch.zeros((buffer_size, action_size), dtype=torch.float32, device=dev)
        self.reward_buffer = torch.zeros((buffer_size, 1), dtype=torch.float32, device=dev)
        self.continue_buffer = torch.zeros((buffer_size, 1), dtype=torch.float32, device=dev)
        self.capacity = buffer_size
        self.sequence_length = sequence_length
        self.device = dev
        self.next_idx = 0
        self.size = 0

    def add_to_buffer(self, observation, action, reward, continue_):
        """One transition (Buffer.py:19-30); the reward is stored symlog'd by the insert kernel."""
        self.add_batch(np.asarray(observation, dtype=np.uint8)[None], np.asarray(action, dtype=np.float32)[None],
                       np.asarray([reward], dtype=np.float32), np.asarray([continue_], dtype=np.float32))

    def add_batch(self, observations, actions, rewards, continues):
        """n consecutive transitions (host arrays or device tensors)."""
        dev = self.device
        obs = torch.as_tensor(observations, dtype=torch.uint8).to(dev, non_blocking=True)
        n = obs.shape[0]
        ops.replay_insert(self.observation_buffer, self.action_buffer, self.reward_buffer, self.continue_buffer, obs,
                          torch.as_tensor(actions, dtype=torch.float32).to(dev).reshape(n, -1),
                          torch.as_tensor(rewards, dtype=torch.float32).to(dev).reshape(n),
                          torch.as_tensor(continues, dtype=torch.float32).to(dev).reshape(n), self.next_idx)
        self.next_idx = (self.next_idx + n) % self.capacity
        self.size = min(self.size + n, self.capacity)

    def draw_starts(self, batch_size):
        if self.size < self.sequence_length:
            raise ValueError("Not enough data in buffer to sample a full sequence")
        valid = self.size - self.sequence_length + 1
        starts = np.random.randint(0, valid, size=batch_size)
        if self.size == self.capacity:
            fixed = []
            for s in starts:
                fixed.append(np.random.randint(0, valid) if s < self.next_idx < s + self.sequence_length else s)
            starts = np.array(fixed)
        return np.asarray(starts, dtype=np.int64)

    def sample_sequences(self, batch_size, normalise=False):
        starts = torch.from_numpy(self.draw_starts(batch_size))
        o, a, r, c = ops.replay_gather(self.observation_buffer, self.action_buffer, self.reward_buffer, self.continue_buffer,
                                       starts, self.sequence_length, normalise=normalise)
        return o, a, r, c, self.sequence_length
