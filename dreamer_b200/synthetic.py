"""Deterministic synthetic weights / inputs for the hot path (used by bench.py, the tests and the oracle's fixture generator).

Nothing here computes the path: it only manufactures random-init weights of the reference architecture and synthetic inputs
of the shapes SURVEY.md section 8d prescribes, from numpy PCG64 streams so that every machine regenerates the same bits.

``make_state_dict`` builds the reference's 97-key ``state_dict`` (names and shapes as produced by
``Dreamer(config)``; Dreamer.py:71-118) from a numpy PCG64 stream, so the same weights can be
regenerated on the GPU box (where /root/reference does not exist) and loaded into the reference
here with ``load_state_dict(strict=True)`` -- which also proves the key/shape inventory.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

REF_CONFIG = dict(  # car_racer_config.yaml:5-52
    hidden_state_dims=600, latent_state_dims=[32, 32], action_dims=3, observation_dims=[64, 64],
    encoder_filter_num_1=32, encoder_filter_num_2=64, encoder_hidden_layer_nodes=200,
    decoder_filter_num_1=32, decoder_filter_num_2=64, decoder_hidden_layer_nodes=200,
    dyn_pred_hidden_num_nodes_1=200, dyn_pred_hidden_num_nodes_2=200,
    rew_pred_hidden_num_nodes_1=200, rew_pred_hidden_num_nodes_2=200,
    cont_pred_hidden_num_nodes_1=200, cont_pred_hidden_num_nodes_2=200,
    hidden_layer_actor_1_size=200, hidden_layer_actor_2_size=200,
    hidden_layer_critic_1_size=200, hidden_layer_critic_2_size=200,
    device="cpu", horizon=30, batch_size=50, nu=0.0003, lambda_=0.95, gamma=0.99,
    buffer_size=200000, sequence_length=50, seed=42, training_iterations=10000, random_iterations=500,
    actor_lr=0.00008, actor_betas=[0.9, 0.999], actor_eps=0.00001,
    critic_lr=0.0001, critic_betas=[0.9, 0.999], critic_eps=0.00001, AC_epochs=2,
    world_model_lr=0.0001, world_model_betas=[0.9, 0.999], world_model_eps=0.00001, WM_epochs=2,
    beta_prediction=1.0, beta_dynamics=0.5, beta_representation=0.1, critic_reward_buckets=255,
    env_id="CarRacing-v3",
)


def small_config(**over):
    """A shrunken config (same topology) whose fixtures are small enough to commit."""
    cfg = dict(REF_CONFIG)
    cfg.update(hidden_state_dims=96, encoder_filter_num_1=8, encoder_filter_num_2=16,
               decoder_filter_num_1=8, decoder_filter_num_2=16,
               encoder_hidden_layer_nodes=72, decoder_hidden_layer_nodes=72,
               dyn_pred_hidden_num_nodes_1=72, dyn_pred_hidden_num_nodes_2=72,
               rew_pred_hidden_num_nodes_1=72, rew_pred_hidden_num_nodes_2=72,
               cont_pred_hidden_num_nodes_1=72, cont_pred_hidden_num_nodes_2=72,
               hidden_layer_actor_1_size=72, hidden_layer_actor_2_size=72,
               hidden_layer_critic_1_size=72, hidden_layer_critic_2_size=72,
               horizon=6, batch_size=4, sequence_length=10, buffer_size=64)
    cfg.update(over)
    return cfg


def _mlp_shapes(prefix, d_in, h1, h2, d_out=None):
    s = [(f"{prefix}.0.weight", (h1, d_in)), (f"{prefix}.0.bias", (h1,)),
         (f"{prefix}.1.weight", (h1,)), (f"{prefix}.1.bias", (h1,)),
         (f"{prefix}.3.weight", (h2, h1)), (f"{prefix}.3.bias", (h2,)),
         (f"{prefix}.4.weight", (h2,)), (f"{prefix}.4.bias", (h2,))]
    if d_out is not None:
        s += [(f"{prefix}.6.weight", (d_out, h2)), (f"{prefix}.6.bias", (d_out,))]
    return s


def state_dict_shapes(cfg):
    """Ordered (key, shape) list mirroring ``Dreamer(config).state_dict()``."""
    D = cfg["hidden_state_dims"]; R, C = cfg["latent_state_dims"]; Z = R * C; A = cfg["action_dims"]
    oh, ow = cfg["observation_dims"]
    e1, e2 = cfg["encoder_filter_num_1"], cfg["encoder_filter_num_2"]
    d1, d2 = cfg["decoder_filter_num_1"], cfg["decoder_filter_num_2"]
    eh, dh = cfg["encoder_hidden_layer_nodes"], cfg["decoder_hidden_layer_nodes"]
    NB = cfg["critic_reward_buckets"]
    feat = e2 * 4 * (oh // 16) * (ow // 16)
    dfeat = d2 * 4 * (oh // 16) * (ow // 16)
    w = "world_model."
    out = []
    fe = w + "encoder.feature_extractor"
    for i, (ci, co) in zip((0, 2, 4, 6), ((3, e1), (e1, e2), (e2, 2 * e2), (2 * e2, 4 * e2))):
        out += [(f"{fe}.{i}.weight", (co, ci, 4, 4)), (f"{fe}.{i}.bias", (co,))]
    lm = w + "encoder.latent_mapper"
    out += [(f"{lm}.0.weight", (eh, feat + D)), (f"{lm}.0.bias", (eh,)), (f"{lm}.1.weight", (eh,)),
            (f"{lm}.1.bias", (eh,)), (f"{lm}.3.weight", (Z, eh)), (f"{lm}.3.bias", (Z,))]
    g = w + "sequence_model.GRU"
    out += [(f"{g}.weight_ih", (3 * D, Z + A)), (f"{g}.weight_hh", (3 * D, D)),
            (f"{g}.bias_ih", (3 * D,)), (f"{g}.bias_hh", (3 * D,))]
    out += _mlp_shapes(w + "dynamics_predictor.logit_net", D, cfg["dyn_pred_hidden_num_nodes_1"],
                       cfg["dyn_pred_hidden_num_nodes_2"], Z)
    out += [(w + "reward_predictor.buckets_rew", (NB,))]
    out += _mlp_shapes(w + "reward_predictor.logit_net", D + Z, cfg["rew_pred_hidden_num_nodes_1"],
                       cfg["rew_pred_hidden_num_nodes_2"], NB)
    out += _mlp_shapes(w + "continue_predictor.logit_generator", D + Z, cfg["cont_pred_hidden_num_nodes_1"],
                       cfg["cont_pred_hidden_num_nodes_2"], 1)
    up = w + "decoder.upscaler"
    out += [(f"{up}.0.weight", (dh, Z + D)), (f"{up}.0.bias", (dh,)), (f"{up}.1.weight", (dh,)),
            (f"{up}.1.bias", (dh,)), (f"{up}.3.weight", (dfeat, dh)), (f"{up}.3.bias", (dfeat,))]
    ib = w + "decoder.image_builder"
    for i, (ci, co) in zip((0, 2, 4, 6), ((4 * d2, 2 * d2), (2 * d2, d2), (d2, d1), (d1, 3))):
        out += [(f"{ib}.{i}.weight", (ci, co, 4, 4)), (f"{ib}.{i}.bias", (co,))]
    a = "agent."
    out += _mlp_shapes(a + "actor.base_net", D + Z, cfg["hidden_layer_actor_1_size"], cfg["hidden_layer_actor_2_size"])
    out += [(a + "actor.mu_head.weight", (A, cfg["hidden_layer_actor_2_size"])), (a + "actor.mu_head.bias", (A,)),
            (a + "actor.log_sig_head.weight", (A, cfg["hidden_layer_actor_2_size"])), (a + "actor.log_sig_head.bias", (A,))]
    for which in ("critic", "target_critic"):
        out += [(a + which + ".buckets_crit", (NB,))]
        out += _mlp_shapes(a + which + ".value_net", D + Z, cfg["hidden_layer_critic_1_size"],
                           cfg["hidden_layer_critic_2_size"], NB)
    return out


def make_state_dict(cfg, seed=0, actor_mu_zero=False, scale=1.0):
    """Random weights with PyTorch-default-like scales (U(-1/sqrt(fan_in), 1/sqrt(fan_in))).

    LayerNorm gains are drawn around 1 and biases around 0 (not exactly 1/0) so that parity tests
    exercise them.  ``actor_mu_zero`` reproduces the reference's zero-initialised mu head
    (Agent.py:188-189).  Bucket buffers are ``torch.linspace(-20, 20, NB)`` as in the reference
    (DynamicsPredictors.py:61; Agent.py:228) -- note buckets[NB//2] is 7.45e-8, not 0.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = OrderedDict()
    for key, shape in state_dict_shapes(cfg):
        if "buckets" in key:
            sd[key] = torch.linspace(-20.0, 20.0, shape[0])
            continue
        leaf = key.rsplit(".", 1)[1]
        is_ln = len(shape) == 1 and leaf == "weight"
        if is_ln:
            v = 1.0 + 0.1 * rng.standard_normal(shape)
        else:
            if len(shape) == 4:
                # conv: (co,ci,4,4) fan_in = ci*16 ; convT: (ci,co,4,4) torch uses size(1)*16
                fan_in = shape[1] * 16
            elif len(shape) == 2:
                fan_in = shape[1]
            else:
                fan_in = max(shape[0], 1)
            k = scale / np.sqrt(fan_in)
            v = rng.uniform(-k, k, size=shape)
        sd[key] = torch.from_numpy(np.asarray(v, dtype=np.float32))
    if actor_mu_zero:
        sd["agent.actor.mu_head.weight"].zero_()
        sd["agent.actor.mu_head.bias"].zero_()
    # the target critic starts as a copy of the critic (Agent.py:50)
    for k in list(sd):
        if k.startswith("agent.critic."):
            sd[k.replace("agent.critic.", "agent.target_critic.")] = sd[k].clone()
    return sd


def rollout_inputs(cfg, B, H, seed=1234):
    """z0 one-hot (B,1,R,C), h0 = tanh(N(0,1)) (B,1,D), uniforms (H,B,R), normals (H,B,A) -- SURVEY 8d."""
    rng = np.random.Generator(np.random.PCG64(seed))
    R, C = cfg["latent_state_dims"]; D = cfg["hidden_state_dims"]; A = cfg["action_dims"]
    idx = rng.integers(0, C, size=(B, 1, R))
    z0 = torch.nn.functional.one_hot(torch.from_numpy(idx), C).float()
    h0 = torch.tanh(torch.from_numpy(rng.standard_normal((B, 1, D)).astype(np.float32)))
    u = torch.from_numpy(rng.random((H, B, R), dtype=np.float32))
    n = torch.from_numpy(rng.standard_normal((H, B, A)).astype(np.float32))
    return z0, h0, u, n


def sequence_inputs(cfg, B, T, seed=4321):
    """obs raw 0..255 fp32 (B,T,3,H,W), act U(-1,1), rew symlog(N(0,1)), cont 1 with one 0 per
    sequence, posterior uniforms (T,B,R) -- SURVEY 8d config C3."""
    rng = np.random.Generator(np.random.PCG64(seed))
    R, _ = cfg["latent_state_dims"]; A = cfg["action_dims"]; oh, ow = cfg["observation_dims"]
    obs = torch.from_numpy(rng.integers(0, 256, size=(B, T, 3, oh, ow)).astype(np.float32))
    act = torch.from_numpy(rng.uniform(-1, 1, size=(B, T, A)).astype(np.float32))
    r = rng.standard_normal((B, T, 1)).astype(np.float32)
    rew = torch.from_numpy(np.sign(r) * np.log1p(np.abs(r)))
    cont = torch.ones(B, T, 1)
    for b in range(B):
        cont[b, int(rng.integers(0, T)), 0] = 0.0
    u = torch.from_numpy(rng.random((T, B, R), dtype=np.float32))
    return obs, act, rew, cont, u


def build_learners(cfg, sd, device):
    """WorldModel + Agent mirrors constructed exactly as ``Dreamer.__init__`` does (Dreamer.py:71-118) with weights `sd`."""
    from . import learners
    dev = torch.device(device)
    wm = learners.WorldModel(cfg["hidden_state_dims"], tuple(cfg["latent_state_dims"]), tuple(cfg["observation_dims"]), cfg["action_dims"],
                             cfg["horizon"], cfg["batch_size"], cfg["world_model_lr"], tuple(cfg["world_model_betas"]), cfg["world_model_eps"],
                             cfg["beta_prediction"], cfg["beta_dynamics"], cfg["beta_representation"], cfg["encoder_filter_num_1"],
                             cfg["encoder_filter_num_2"], cfg["encoder_hidden_layer_nodes"], cfg["decoder_filter_num_1"], cfg["decoder_filter_num_2"],
                             cfg["decoder_hidden_layer_nodes"], cfg["dyn_pred_hidden_num_nodes_1"], cfg["dyn_pred_hidden_num_nodes_2"],
                             cfg["rew_pred_hidden_num_nodes_1"], cfg["rew_pred_hidden_num_nodes_2"], cfg["critic_reward_buckets"],
                             cfg["cont_pred_hidden_num_nodes_1"], cfg["cont_pred_hidden_num_nodes_2"], device=dev)
    ag = learners.Agent(cfg["action_dims"], tuple(cfg["latent_state_dims"]), cfg["hidden_state_dims"], cfg["hidden_layer_actor_1_size"],
                        cfg["hidden_layer_actor_2_size"], cfg["hidden_layer_critic_1_size"], cfg["hidden_layer_critic_2_size"],
                        cfg["critic_reward_buckets"], cfg["actor_lr"], tuple(cfg["actor_betas"]), cfg["actor_eps"], cfg["critic_lr"],
                        tuple(cfg["critic_betas"]), cfg["critic_eps"], cfg["nu"], cfg["lambda_"], cfg["gamma"], device=dev)
    wm.load_state_dict({k[len("world_model."):]: v for k, v in sd.items() if k.startswith("world_model.")}, strict=True)
    ag.load_state_dict({k[len("agent."):]: v for k, v in sd.items() if k.startswith("agent.")}, strict=True)
    return wm, ag
