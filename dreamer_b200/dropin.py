"""Register the mirrored classes under the reference's flat module names.

    import dreamer_b200.dropin as dropin
    dropin.install()              # before `from Dreamer import Dreamer`
    from Dreamer import Dreamer   # the reference's unchanged orchestrator now builds the sm_100a modules

After ``install()``, ``Dreamer.py:6-8`` (``from WorldModel import WorldModel`` ...) resolve to this package.
``patch_dreamer(DreamerClass)`` additionally swaps ``Dreamer.dream_episodes`` for the fused rollout.
"""
from __future__ import annotations

import sys
import types

import numpy as np
import torch

from . import learners, modules, rollout

_NAMES = {
    "SequenceModel": dict(SequenceModel=modules.SequenceModel),
    "DynamicsPredictors": dict(DynamicsPredictor=modules.DynamicsPredictor, RewardPredictor=modules.RewardPredictor,
                               ContinuePredictor=modules.ContinuePredictor),
    "VariationalAutoEncoder": dict(Encoder=modules.Encoder, Decoder=modules.Decoder),
    "WorldModel": dict(WorldModel=learners.WorldModel),
    "Agent": dict(Agent=learners.Agent, Actor=modules.Actor, Critic=modules.Critic),
    "Buffer": dict(Buffer=modules.Buffer),
}


def _sanitize_for_save(data_list):
    """DreamerUtils.py:52-63."""
    clean = []
    for item in data_list:
        if isinstance(item, torch.Tensor):
            clean.append(item.detach().cpu().item())
        elif isinstance(item, list):
            clean.append([x.detach().cpu().item() if isinstance(x, torch.Tensor) else x for x in item])
        else:
            clean.append(item)
    return np.array(clean)


def install():
    for name, attrs in _NAMES.items():
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        mod.__dict__["__dreamer_b200__"] = True
        sys.modules[name] = mod
    utils = types.ModuleType("DreamerUtils")
    utils.__dict__.update(symlog=modules.symlog, symexp=modules.symexp, symlog_np=modules.symlog_np, to_twohot=modules.to_twohot,
                          _sanitize_for_save=_sanitize_for_save, __dreamer_b200__=True)
    sys.modules["DreamerUtils"] = utils


def uninstall():
    for name in list(_NAMES) + ["DreamerUtils"]:
        if getattr(sys.modules.get(name), "__dreamer_b200__", False):
            del sys.modules[name]


def patch_dreamer(dreamer_cls):
    """Replace the per-step Python loop of Dreamer.dream_episodes (Dreamer.py:143-175) by the fused rollout."""
    def dream_episodes(self, starting_latent_state_batch, starting_hidden_state_batch):
        self.agent.attach_world_model(self.world_model)       # Agent.train_step: actor gradient through the imagined states
        return rollout.dream_episodes_modules(self.world_model, self.agent, starting_latent_state_batch, starting_hidden_state_batch, self.horizon)
    dreamer_cls.dream_episodes = dream_episodes
    return dreamer_cls
