"""Register the mirrored classes under the reference's flat module names.

    import dreamer_b200.dropin as dropin
    dropin.install()              # before `from Dreamer import Dreamer`
    from Dreamer import Dreamer   # the reference's unchanged orchestrator now builds the sm_100a modules

After ``install()``, ``Dreamer.py:6-8`` (``from WorldModel import WorldModel`` ...) resolve to this package.
``patch_dreamer(DreamerClass)`` additionally swaps ``Dreamer.dream_episodes / warm_start_generator / rollout_policy /
evaluate_agent`` for their fused forms.
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


def _acting(dreamer, deterministic):
    """One cached acting.ActingPath per policy mode on the Dreamer instance (stochastic: writes the replay ring)."""
    from .acting import ActingPath
    cache = dreamer.__dict__.setdefault("_b200_acting", {})
    if deterministic not in cache:
        cache[deterministic] = ActingPath(dreamer.world_model, dreamer.agent, None if deterministic else dreamer.buffer,
                                          deterministic=deterministic)
    return cache[deterministic]


def _chw(observation):
    return np.ascontiguousarray(np.asarray(observation).transpose(2, 0, 1)).astype(np.uint8)


def patch_dreamer(dreamer_cls):
    """Swap the hot call sites of the reference's ``Dreamer`` for their fused forms (same signatures, same side effects):

      dream_episodes        Dreamer.py:143-175   one fused rollout instead of horizon x (Actor.act -> imagine_step)
      warm_start_generator  Dreamer.py:244-262   one fused posterior scan instead of sequence_length // 2 observe_steps
      rollout_policy        Dreamer.py:177-226   acting.ActingPath: device-resident agent state, transitions written into the HBM
      evaluate_agent        Dreamer.py:295-322   ring from device tensors, observe_step + act replayed as one CUDA graph
    """
    def dream_episodes(self, starting_latent_state_batch, starting_hidden_state_batch):
        self.agent.attach_world_model(self.world_model)       # Agent.train_step: actor gradient through the imagined states
        return rollout.dream_episodes_modules(self.world_model, self.agent, starting_latent_state_batch, starting_hidden_state_batch, self.horizon)

    def warm_start_generator(self, observation_seq_batch, action_seq_batch, sequence_length):
        wm = self.world_model
        B, Wn = observation_seq_batch.shape[0], sequence_length // 2
        obs = (observation_seq_batch[:, :Wn].float() / 255.0) - 0.5
        u = torch.rand(Wn, B, wm.latent_num_rows, device=obs.device)
        sc = wm._engine.observe(B, Wn).scan(obs, action_seq_batch[:, :Wn], u, warm_start=True, want_logits=False, want_idx=False)
        return sc["latent"][:, -1:].contiguous(), sc["hidden"][:, -1:].contiguous()

    def rollout_policy(self, env, random_policy=False):
        ap = _acting(self, False)
        # an episode continues across training phases (Dreamer.py:340-343 trains between two calls): act with the CURRENT weights,
        # as the reference does -- a cheap version check, the re-pack writes in place so the captured graphs pick it up
        ap.sync_weights()
        if self.agent_obs is None:                              # first call: start an episode (Dreamer.py:184-191)
            observation, _ = env.reset(seed=self.seed)
            ap.reset(_chw(observation))
            self.agent_obs = True
        for _ in range(self.sequence_length):
            if random_policy:
                action_np = np.asarray(env.action_space.sample(), dtype=np.float32)
                ap.set_action(action_np)
            else:
                action_np = ap.act()
            observation_, reward, terminated, truncated, _ = env.step(action_np)
            done = bool(terminated or truncated)
            ap.record(reward, 1 - done)                         # buffer.add_to_buffer(obs_t, a_t, r_t, c_t) (Dreamer.py:211-212)
            if done:
                self.seed += 1
                observation, _ = env.reset(seed=self.seed)
                ap.reset(_chw(observation))
            else:
                ap.observe(_chw(observation_))

    def evaluate_agent(self, env, eval_episodes):
        ap = _acting(self, True)
        ap.sync_weights()
        totals = []
        for _ in range(eval_episodes):
            self.seed += 1
            observation, _ = env.reset(seed=self.seed)
            ap.reset(_chw(observation))
            action_np, total, done = ap.act(), 0.0, False
            while not done:
                observation_, reward, terminated, truncated, _ = env.step(action_np)
                total += float(reward)
                done = bool(terminated or truncated)
                if not done:
                    action_np = ap.step(_chw(observation_), reward, 1.0)     # observe_step + act (no ring: evaluation)
            totals.append(total)
        return torch.tensor(totals, dtype=torch.float32, device=next(self.world_model.parameters()).device).mean()

    dreamer_cls.dream_episodes = dream_episodes
    dreamer_cls.warm_start_generator = warm_start_generator
    dreamer_cls.rollout_policy = rollout_policy
    dreamer_cls.evaluate_agent = evaluate_agent
    return dreamer_cls
