"""The hot call sites of the reference orchestrator (Dreamer.py:143-175, 228-287) on the mirrored modules.

``HotPath`` owns a WorldModel, an Agent and a Buffer built from the reference's config dict exactly as ``Dreamer.__init__``
does (Dreamer.py:71-125) and re-implements only the four methods on the hot path:

    train_world_model     Dreamer.py:228-242   sample_sequences -> WorldModel.training_step, WM_epochs times
    warm_start_generator  Dreamer.py:244-262   ONE fused posterior scan (mode 1) instead of sequence_length // 2 observe_steps
    dream_episodes        Dreamer.py:143-175   ONE fused rollout instead of horizon x (Actor.act -> imagine_step)
    train_Agent           Dreamer.py:264-287   sample -> warm start -> dream -> Agent.train_step, AC_epochs times

plus the two file formats either side of it:

    save_trained_Dreamer / load_pretrained_dreamer   Dreamer.py:289-293   the reference's ``.pth``: state_dict with its 97 keys
    save_training_logs                               Dreamer.py:356-364   the reference's ``training_logs.npz`` (same four keys)
    save_training_state / load_training_state        (the reference saves none of this)  optimisers, return scale S, RNG, ring

and ``acting()`` -- the B = 1 acting path of Dreamer.rollout_policy / evaluate_agent (acting.ActingPath, 8f rank 3).
Environment stepping, evaluation and logging stay in the reference's ``Dreamer.py`` (out of scope).  Under
``torch.distributed`` every rank draws its own replay windows / start states (data parallel).
"""
from __future__ import annotations

import torch

from . import rollout


class HotPath:
    def __init__(self, config: dict, device, cuda_graphs: bool = False):
        """`cuda_graphs=True` replays WorldModel.training_step and Agent.train_step as CUDA graphs after three eager steps per
        input shape (learners.enable_cuda_graphs; needs a CUDA device)."""
        self.cfg = dict(config)
        self.device = torch.device(device)
        from .modules import Buffer
        self.world_model, self.agent = _build(self.cfg, self.device)
        self.agent.attach_world_model(self.world_model)      # actor gradient through the imagined states (bptt.actor_backward)
        self.cuda_graphs = bool(cuda_graphs)
        if cuda_graphs:
            self.world_model.enable_cuda_graphs()
            self.agent.enable_cuda_graphs()
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
    def dream_episodes(self, starting_latent_state_batch, starting_hidden_state_batch, uniforms=None, normals=None, graphed=False):
        return rollout.dream_episodes_modules(self.world_model, self.agent, starting_latent_state_batch, starting_hidden_state_batch,
                                              self.horizon, uniforms, normals, graphed=graphed)

    # Dreamer.py:264-287
    def train_Agent(self):
        la, lc = [], []
        for _ in range(self.AC_epochs):
            obs, act, _, _, L = self.buffer.sample_sequences(batch_size=self.batch_size)
            z0, h0 = self.warm_start_generator(obs, act, L)
            # with cuda_graphs the rollout is one graph replay; its static outputs are consumed by train_step before the next call
            z, h, a, r, c, mu, sg = self.dream_episodes(z0, h0, graphed=self.cuda_graphs)
            loss_actor, loss_critic = self.agent.train_step(z, h, r, c, a, mu, sg)
            la.append(loss_actor); lc.append(loss_critic)
        return torch.stack(la).mean(dim=0), torch.stack(lc).mean(dim=0)


    # ---- 8f rank 3: the acting path -------------------------------------------------------------------------------------
    def acting(self, deterministic: bool = False, use_graphs: bool = True):
        """ActingPath on this HotPath's models and replay ring (one per policy mode; cached)."""
        from .acting import ActingPath
        key = (bool(deterministic), bool(use_graphs))
        cache = self.__dict__.setdefault("_acting", {})
        if key not in cache:
            cache[key] = ActingPath(self.world_model, self.agent, self.buffer, deterministic=deterministic, use_graphs=use_graphs)
        return cache[key]

    # ---- 8f rank 4: checkpoint formats ---------------------------------------------------------------------------------
    def state_dict(self):
        """The reference Dreamer's state_dict: ``world_model.*`` then ``agent.*`` (97 keys at the reference config)."""
        sd = {"world_model." + k: v for k, v in self.world_model.state_dict().items()}
        sd.update({"agent." + k: v for k, v in self.agent.state_dict().items()})
        return sd

    def load_state_dict(self, sd, strict: bool = True):
        self.world_model.load_state_dict({k[len("world_model."):]: v for k, v in sd.items() if k.startswith("world_model.")}, strict=strict)
        self.agent.load_state_dict({k[len("agent."):]: v for k, v in sd.items() if k.startswith("agent.")}, strict=strict)
        if strict:
            extra = [k for k in sd if not (k.startswith("world_model.") or k.startswith("agent."))]
            if extra:
                raise RuntimeError(f"unexpected keys in checkpoint: {extra[:5]}")

    def save_trained_Dreamer(self, save_path):                       # Dreamer.py:292-293
        torch.save({k: v.detach().cpu() for k, v in self.state_dict().items()}, save_path)

    def load_pretrained_dreamer(self, path):                         # Dreamer.py:289-290
        self.load_state_dict(torch.load(path, weights_only=True, map_location=self.device))

    @staticmethod
    def save_training_logs(path, world_model_loss, actor_loss, critic_loss, rewards):
        """The reference's ``training_logs.npz`` (Dreamer.py:356-364): the same four keys, tensors reduced to floats."""
        import numpy as np
        from .dropin import _sanitize_for_save
        np.savez(path, world_model_loss=_sanitize_for_save(world_model_loss), actor_loss=_sanitize_for_save(actor_loss),
                 critic_loss=_sanitize_for_save(critic_loss), rewards=_sanitize_for_save(rewards))

    def save_training_state(self, path, include_buffer: bool = False):
        """Everything a bit-exact resume needs and the reference omits: optimiser moments and step counts, the return scale S,
        torch / numpy RNG states, replay cursor (and, optionally, the ring contents)."""
        import numpy as np
        buf = self.buffer
        state = dict(model={k: v.detach().cpu() for k, v in self.state_dict().items()},
                     optim=dict(world_model=self.world_model.optimiser.state_dict(), actor=self.agent.actor_optimiser.state_dict(),
                                critic=self.agent.critic_optimiser.state_dict()),
                     S=float(self.agent.S), torch_rng=torch.get_rng_state(),
                     cuda_rng=torch.cuda.get_rng_state(self.device) if self.device.type == "cuda" else None,
                     numpy_rng=np.random.get_state(), buffer=dict(size=buf.size, next_idx=buf.next_idx, capacity=buf.capacity))
        if include_buffer:
            n = buf.size
            state["buffer"].update(observation=buf.observation_buffer[:n].cpu(), action=buf.action_buffer[:n].cpu(),
                                   reward=buf.reward_buffer[:n].cpu(), cont=buf.continue_buffer[:n].cpu())
        torch.save(state, path)

    def load_training_state(self, path):
        import numpy as np
        state = torch.load(path, weights_only=False, map_location="cpu")
        self.load_state_dict({k: v.to(self.device) for k, v in state["model"].items()})
        self.world_model.optimiser.load_state_dict(state["optim"]["world_model"])
        self.agent.actor_optimiser.load_state_dict(state["optim"]["actor"])
        self.agent.critic_optimiser.load_state_dict(state["optim"]["critic"])
        if isinstance(self.agent.S, torch.Tensor):
            self.agent.S.fill_(state["S"])
        else:
            self.agent.S = state["S"]
        torch.set_rng_state(state["torch_rng"])
        if state.get("cuda_rng") is not None and self.device.type == "cuda":
            torch.cuda.set_rng_state(state["cuda_rng"], self.device)
        np.random.set_state(state["numpy_rng"])
        b = state["buffer"]
        if "observation" in b:
            n = b["size"]
            buf = self.buffer
            buf.observation_buffer[:n].copy_(b["observation"]); buf.action_buffer[:n].copy_(b["action"])
            buf.reward_buffer[:n].copy_(b["reward"]); buf.continue_buffer[:n].copy_(b["cont"])
            buf.size, buf.next_idx = b["size"], b["next_idx"]


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
