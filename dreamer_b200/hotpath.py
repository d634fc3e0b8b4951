"""The hot call sites of the reference orchestrator (Dreamer.py:143-175, 228-287) on the mirrored modules.

``HotPath`` owns a WorldModel, an Agent and a Buffer built from the reference's config dict exactly as ``Dreamer.__init__``
does (Dreamer.py:71-125) and re-implements only the four methods on the hot path:

    train_world_model     Dreamer.py:228-242   sample_sequences -> WorldModel.training_step, WM_epochs times
    warm_start_generator  Dreamer.py:244-262   ONE fused posterior scan (mode 1) instead of sequence_length // 2 observe_steps
    dream_episodes        Dreamer.py:143-175   ONE fused rollout instead of horizon x (Actor.act -> imagine_step)
    train_Agent           Dreamer.py:264-287   sample -> warm start -> dream -> Agent.train_step, AC_epochs times

Environment stepping, evaluation, checkpoints and logging stay in the reference's ``Dreamer.py`` (out of scope).  Under
``torch.distributed`` every rank draws its own replay windows / start states (data parallel).
"""
from __future__ import annotations

import torch

from . import rollout


class HotPath:
    def __init__(self, config: dict, device):
        self.cfg = dict(config)
        self.device = torch.device(device)
        from .modules import Buffer
        self.world_model, self.agent = _build(self.cfg, self.device)
        self.buffer = Buffer(config["buffer_size"], config["sequence_length"], config["action_dims"], tuple(config["observation_dims"]),
                             device=self.device)
        self.horizon = config["horizon"]
        self.batch_size = config["batch_size"]
        self.sequence_length = config["sequence_length"]
        self.WM_epochs = config["WM_epochs"]
        self.AC_epochs = config["AC_epochs"]
        self.hidden_state_dims = config["hidden_state_dims"]

    # Dreamer.py:228-242
    def train_world_model(self):
        losses = []
        for _ in range(self.WM_epochs):
            obs, act, rew, cont, _ = self.buffer.sample_sequences(batch_size=self.batch_size)
            losses.append(self.world_model.training_step(obs, act, rew, cont))
        return losses

    # Dreamer.py:244-262
    def warm_start_generator(self, observation_seq_batch, action_seq_batch, sequence_length, uniforms=None):
        wm = self.world_model
        B = observation_seq_batch.shape[0]
        W = sequence_length // 2
        obs = (observation_seq_batch[:, :W].float() / 255.0) - 0.5
        if uniforms is None:
            uniforms = torch.rand(W, B, wm.latent_num_rows, device=obs.device)
        sc = wm._engine.observe(B, W).scan(obs, action_seq_batch[:, :W], uniforms, warm_start=True, want_logits=False, want_idx=False)
        return sc["latent"][:, -1:].contiguous(), sc["hidden"][:, -1:].contiguous()

    # Dreamer.py:143-175
    def dream_episodes(self, starting_latent_state_batch, starting_hidden_state_batch, uniforms=None, normals=None):
        return rollout.dream_episodes_modules(self.world_model, self.agent, starting_latent_state_batch, starting_hidden_state_batch,
                                              self.horizon, uniforms, normals)

    # Dreamer.py:264-287
    def train_Agent(self):
        la, lc = [], []
        for _ in range(self.AC_epochs):
            obs, act, _, _, L = self.buffer.sample_sequences(batch_size=self.batch_size)
            z0, h0 = self.warm_start_generator(obs, act, L)
            z, h, a, r, c, mu, sg = self.dream_episodes(z0, h0)
            loss_actor, loss_critic = self.agent.train_step(z, h, r, c, a, mu, sg)
            la.append(loss_actor); lc.append(loss_critic)
        return torch.stack(la).mean(dim=0), torch.stack(lc).mean(dim=0)


def _build(cfg, device):
    """Default-initialised WorldModel + Agent (the reference's init: torch defaults, zero mu head)."""
    from . import learners
    wm = learners.WorldModel(cfg["hidden_state_dims"], tuple(cfg["latent_state_dims"]), tuple(cfg["observation_dims"]), cfg["action_dims"],
                             cfg["horizon"], cfg["batch_size"], cfg["world_model_lr"], tuple(cfg["world_model_betas"]), cfg["world_model_eps"],
                             cfg["beta_prediction"], cfg["beta_dynamics"], cfg["beta_representation"], cfg["encoder_filter_num_1"],
                             cfg["encoder_filter_num_2"], cfg["encoder_hidden_layer_nodes"], cfg["decoder_filter_num_1"], cfg["decoder_filter_num_2"],
                             cfg["decoder_hidden_layer_nodes"], cfg["dyn_pred_hidden_num_nodes_1"], cfg["dyn_pred_hidden_num_nodes_2"],
                             cfg["rew_pred_hidden_num_nodes_1"], cfg["rew_pred_hidden_num_nodes_2"], cfg["critic_reward_buckets"],
                             cfg["cont_pred_hidden_num_nodes_1"], cfg["cont_pred_hidden_num_nodes_2"], device=device)
    ag = learners.Agent(cfg["action_dims"], tuple(cfg["latent_state_dims"]), cfg["hidden_state_dims"], cfg["hidden_layer_actor_1_size"],
                        cfg["hidden_layer_actor_2_size"], cfg["hidden_layer_critic_1_size"], cfg["hidden_layer_critic_2_size"],
                        cfg["critic_reward_buckets"], cfg["actor_lr"], tuple(cfg["actor_betas"]), cfg["actor_eps"], cfg["critic_lr"],
                        tuple(cfg["critic_betas"]), cfg["critic_eps"], cfg["nu"], cfg["lambda_"], cfg["gamma"], device=device)
    return wm, ag
