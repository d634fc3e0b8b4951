"""Functional fp32 CPU restatement of the reference RSSM hot path (TEST INFRASTRUCTURE).

Every function takes the reference ``state_dict`` (97 keys) and plain
tensors.  Randomness is *always* host-supplied: categorical draws consume uniforms through an
inverse-CDF rule, actor draws consume standard normals.  The reference itself draws from the
global torch RNG (``Categorical.sample`` = an exponential race, DynamicsPredictors.py:36-37),
which no "same uniforms" contract can reproduce, so ``oracle/make_golden.py`` patches exactly
those draw sites when it runs the reference and this file restates the result.

Sampling contract (shared with the CUDA kernels, include/dreamer_b200.h):
    p    = 0.99 * softmax(logits) + 0.01 / C          (fp32)
    cdf  = inclusive prefix sum of p, left to right    (fp32)
    idx  = min(C - 1, #{k : cdf[k] <= u})
    z    = (onehot(idx) + p) - p                       (the reference's straight-through value)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
WM = "world_model."
AG = "agent."


# --------------------------------------------------------------------------------------
# DreamerUtils.py:29-50
# --------------------------------------------------------------------------------------
def symlog(x: torch.Tensor) -> torch.Tensor:
    """DreamerUtils.py:29-30."""
    return torch.sign(x) * torch.log(1.0 + torch.abs(x))


def symexp(x: torch.Tensor) -> torch.Tensor:
    """DreamerUtils.py:35-37 (clamp to +-20, exp in fp32)."""
    x = x.clamp(-20.0, 20.0)
    return torch.sign(x) * (torch.exp(x.abs().float()) - 1.0)


def twohot_index_weight(value: torch.Tensor, buckets: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Lower bucket index and upper-bucket weight of DreamerUtils.py:39-47.

    value (..., 1) -> idx (..., 1) int64, w (..., 1) fp32 with twohot[idx] = 1-w, twohot[idx+1] = w.
    """
    v = value.clamp(min=float(buckets.min()), max=float(buckets.max()))
    idx = torch.searchsorted(buckets, v.contiguous(), right=True) - 1
    idx = idx.clamp(max=buckets.numel() - 2)
    lo = buckets[idx]
    hi = buckets[idx + 1]
    w = (v - lo) / (hi - lo + 1e-8)
    return idx, w


def to_twohot(value: torch.Tensor, buckets: torch.Tensor) -> torch.Tensor:
    """DreamerUtils.py:39-50."""
    idx, w = twohot_index_weight(value, buckets)
    out = torch.zeros(value.shape[:-1] + (buckets.numel(),), dtype=torch.float32)
    out = out.scatter(-1, idx, 1.0 - w)
    out = out.scatter(-1, idx + 1, w)
    return out


def twohot_ce(logits: torch.Tensor, value: torch.Tensor, buckets: torch.Tensor) -> torch.Tensor:
    """sum(twohot(value) * log_softmax(logits)) as in WorldModel.py:137-138 / Agent.py:133-134.

    logits (..., NB), value (..., 1) -> (..., 1) log-likelihood (NOT negated).
    """
    return (to_twohot(value, buckets) * F.log_softmax(logits.float(), dim=-1)).sum(-1, keepdim=True)


# --------------------------------------------------------------------------------------
# categorical 32x32 head: DynamicsPredictors.py:31-40, VariationalAutoEncoder.py:85-99
# --------------------------------------------------------------------------------------
def unimix_probs(logits: torch.Tensor) -> torch.Tensor:
    C = logits.shape[-1]
    return 0.99 * torch.softmax(logits.float(), dim=-1) + 0.01 * (1.0 / C)


def cumsum_f32(p: torch.Tensor) -> torch.Tensor:
    """Inclusive left-to-right prefix sum with every partial sum rounded to fp32 (the contract's
    cdf; torch.cumsum on CPU may accumulate in a wider type)."""
    out = torch.empty_like(p)
    acc = torch.zeros_like(p[..., 0])
    for k in range(p.shape[-1]):
        acc = acc + p[..., k]
        out[..., k] = acc
    return out


def interior_uniforms(p: torch.Tensor, u: torch.Tensor, margin_frac: float = 0.0, delta: float = 1e-5) -> torch.Tensor:
    """Move each uniform into the interior of the bin the contract selects for it.

    The selected index is unchanged; the draw merely becomes insensitive to ULP-level (delta) or
    bf16-level (margin_frac of the bin width) perturbations of the CDF (SURVEY.md section 7 hard part b).
    """
    cdf = cumsum_f32(p)
    C = p.shape[-1]
    idx = (cdf <= u.unsqueeze(-1)).sum(-1).clamp(max=C - 1)
    hi = cdf.gather(-1, idx.unsqueeze(-1)).squeeze(-1)
    lo = torch.where(idx > 0, cdf.gather(-1, (idx - 1).clamp(min=0).unsqueeze(-1)).squeeze(-1), torch.zeros_like(hi))
    hi = torch.where(idx == C - 1, torch.ones_like(hi), hi)
    pad = torch.maximum(torch.full_like(hi, delta), margin_frac * (hi - lo))
    pad = torch.minimum(pad, 0.5 * (hi - lo))
    out = torch.minimum(torch.maximum(u, lo + pad), hi - pad)
    return out.clamp(0.0, 1.0 - 1e-7)


def categorical_st(logits: torch.Tensor, u: Optional[torch.Tensor]):
    """logits (..., R, C), u (..., R) in [0,1) -> (z_st, idx, p).

    Restates softmax -> unimix -> sample -> one_hot -> straight-through
    (DynamicsPredictors.py:33-39; VariationalAutoEncoder.py:88-98) with the inverse-CDF rule.
    ``u=None`` draws with the stock sampler exactly as the reference does
    (``Categorical(probs=p).sample()``); only bench.py's CPU-baseline timing uses that mode, so the
    timed CPU work has the reference's cost structure (the sampler dominates it).
    """
    p = unimix_probs(logits)
    if u is None:
        idx = torch.distributions.Categorical(probs=p).sample()
        oh = F.one_hot(idx, p.shape[-1]).float()
        return (oh + p) - p, idx, p
    cdf = cumsum_f32(p)
    C = p.shape[-1]
    idx = (cdf <= u.unsqueeze(-1)).sum(-1).clamp(max=C - 1)
    oh = F.one_hot(idx, C).float()
    z = (oh + p) - p
    return z, idx, p


def categorical_kl_terms(post_logits: torch.Tensor, prior_logits: torch.Tensor) -> torch.Tensor:
    """sum over the R rows of KL(Cat(post) || Cat(prior)) per (b, t): WorldModel.py:175-181.

    Both Dkl_dyn and Dkl_rep have this forward value (they differ only in which side is detached).
    """
    lp = F.log_softmax(post_logits.float(), dim=-1)
    lq = F.log_softmax(prior_logits.float(), dim=-1)
    return (lp.exp() * (lp - lq)).sum(-1).sum(-1)


# --------------------------------------------------------------------------------------
# small building blocks
# --------------------------------------------------------------------------------------
def _mlp(sd: SD, prefix: str, x: torch.Tensor, n_hidden: int = 2) -> torch.Tensor:
    """Linear-LN-SiLU x n_hidden, then the final Linear at index 3*n_hidden.

    Layout of the reference ``nn.Sequential`` heads (DynamicsPredictors.py:15-23, 52-60, 85-93;
    Agent.py:219-227): indices 0,1,(2) / 3,4,(5) / 6.
    """
    for i in range(n_hidden):
        k = 3 * i
        x = F.linear(x, sd[f"{prefix}.{k}.weight"], sd[f"{prefix}.{k}.bias"])
        x = F.layer_norm(x, (x.shape[-1],), sd[f"{prefix}.{k + 1}.weight"], sd[f"{prefix}.{k + 1}.bias"], 1e-5)
        x = F.silu(x)
    k = 3 * n_hidden
    if f"{prefix}.{k}.weight" in sd:
        x = F.linear(x, sd[f"{prefix}.{k}.weight"], sd[f"{prefix}.{k}.bias"])
    return x


def gru_step(sd: SD, z: torch.Tensor, h: torch.Tensor, a: torch.Tensor) -> torch.Tensor:
    """SequenceModel.py:19-24 (nn.GRUCell, gate order [r; z; n], no LayerNorm).

    z (B, R, C) or (B, R*C); h (B, D); a (B, A) -> h' (B, D).
    """
    p = WM + "sequence_model.GRU."
    x = torch.cat([z.reshape(z.shape[0], -1), a], dim=-1)
    gi = F.linear(x, sd[p + "weight_ih"], sd[p + "bias_ih"])
    gh = F.linear(h, sd[p + "weight_hh"], sd[p + "bias_hh"])
    D = h.shape[-1]
    r = torch.sigmoid(gi[:, :D] + gh[:, :D])
    u = torch.sigmoid(gi[:, D:2 * D] + gh[:, D:2 * D])
    n = torch.tanh(gi[:, 2 * D:] + r * gh[:, 2 * D:])
    return (1.0 - u) * n + u * h


def prior_logits(sd: SD, h: torch.Tensor, R: int = 32, C: int = 32) -> torch.Tensor:
    """DynamicsPredictors.py:25-29.  h (N, D) -> (N, R, C)."""
    return _mlp(sd, WM + "dynamics_predictor.logit_net", h).reshape(h.shape[0], R, C)


def reward_logits(sd: SD, h: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """DynamicsPredictors.py:64-68.  input order [h, z]."""
    return _mlp(sd, WM + "reward_predictor.logit_net", torch.cat([h, z.reshape(z.shape[0], -1)], -1))


def reward_predict(sd: SD, h: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """DynamicsPredictors.py:70-74."""
    p = torch.softmax(reward_logits(sd, h, z), dim=-1)
    return symexp((p * sd[WM + "reward_predictor.buckets_rew"]).sum(-1, keepdim=True))


def continue_logit(sd: SD, h: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """DynamicsPredictors.py:95-100 (the logit; probability = sigmoid)."""
    return _mlp(sd, WM + "continue_predictor.logit_generator", torch.cat([h, z.reshape(z.shape[0], -1)], -1))


def actor_forward(sd: SD, h: torch.Tensor, z: torch.Tensor):
    """Agent.py:191-200.  -> (mu, sigma) each (N, A)."""
    s = torch.cat([h, z.reshape(z.shape[0], -1)], -1)
    base = _mlp(sd, AG + "actor.base_net", s)  # two LN-SiLU blocks, no final linear in base_net
    mu = F.linear(base, sd[AG + "actor.mu_head.weight"], sd[AG + "actor.mu_head.bias"])
    ls = F.linear(base, sd[AG + "actor.log_sig_head.weight"], sd[AG + "actor.log_sig_head.bias"])
    sigma = F.softplus(ls.clamp(-5.0, 2.0)) + 1e-3
    return mu, sigma


def actor_act(sd: SD, h: torch.Tensor, z: torch.Tensor, eps: torch.Tensor):
    """Agent.py:202-210: tanh-Normal rsample == tanh(mu + sigma * eps)."""
    mu, sigma = actor_forward(sd, h, z)
    return torch.tanh(mu + sigma * eps), mu, sigma


def critic_logits(sd: SD, h: torch.Tensor, z: torch.Tensor, which: str = "critic") -> torch.Tensor:
    """Agent.py:231-235."""
    return _mlp(sd, AG + which + ".value_net", torch.cat([h, z.reshape(z.shape[0], -1)], -1))


def critic_value(sd: SD, h: torch.Tensor, z: torch.Tensor, which: str = "critic") -> torch.Tensor:
    """Agent.py:237-241."""
    p = torch.softmax(critic_logits(sd, h, z, which), dim=-1)
    return symexp((p * sd[AG + which + ".buckets_crit"]).sum(-1, keepdim=True))


# --------------------------------------------------------------------------------------
# imagination: WorldModel.py:72-77 + Dreamer.py:143-175
# --------------------------------------------------------------------------------------
def imagine_step(sd: SD, h: torch.Tensor, z: torch.Tensor, a: torch.Tensor, u: torch.Tensor,
                 margin_frac: float = 0.0, delta: float = 0.0):
    """One WorldModel.imagine_step.  h (B,D), z (B,R,C), a (B,A), u (B,R).

    Returns h', z', reward (B,1), continue prob (B,1), prior logits, idx, uniforms actually used.
    """
    h2 = gru_step(sd, z, h, a)
    logits = prior_logits(sd, h2, z.shape[-2], z.shape[-1])
    if u is not None and (margin_frac > 0.0 or delta > 0.0):
        u = interior_uniforms(unimix_probs(logits), u, margin_frac, delta)
    z2, idx, _ = categorical_st(logits, u)
    r = reward_predict(sd, h2, z2)
    c = torch.sigmoid(continue_logit(sd, h2, z2))
    return h2, z2, r, c, logits, idx, u


def dream_episodes(sd: SD, z0: torch.Tensor, h0: torch.Tensor, uniforms: torch.Tensor, normals: torch.Tensor,
                   margin_frac: float = 0.0, delta: float = 0.0):
    """Dreamer.dream_episodes (Dreamer.py:143-175) with host-supplied randomness.

    z0 (B,1,R,C), h0 (B,1,D), uniforms (H,B,R) or None (stock sampler, timing only), normals (H,B,A).
    Returns the reference 7-tuple (latent (B,H+1,R,C), hidden (B,H+1,D), actions, rewards,
    continues, mu, sigma) followed by extras: idx (B,H,R) int64, prior logits (B,H,R,C), and the
    uniforms actually consumed (H,B,R).
    """
    H = normals.shape[0]
    h = h0[:, 0]
    z = z0[:, 0]
    Z, Hs, A, Rw, Cn, MU, SG, IDX, LG, U = [z], [h], [], [], [], [], [], [], [], []
    for t in range(H):
        a, mu, sg = actor_act(sd, h, z, normals[t])
        h, z, r, c, lg, idx, u = imagine_step(sd, h, z, a, None if uniforms is None else uniforms[t], margin_frac, delta)
        Z.append(z); Hs.append(h); A.append(a); Rw.append(r); Cn.append(c); MU.append(mu); SG.append(sg)
        IDX.append(idx); LG.append(lg); U.append(u)
    st = lambda xs: torch.stack(xs, dim=1)
    used = None if uniforms is None else torch.stack(U, 0)
    return (st(Z), st(Hs), st(A), st(Rw), st(Cn), st(MU), st(SG), st(IDX), st(LG), used)


# --------------------------------------------------------------------------------------
# VAE: VariationalAutoEncoder.py
# --------------------------------------------------------------------------------------
def encoder_features(sd: SD, obs: torch.Tensor) -> torch.Tensor:
    """VariationalAutoEncoder.py:33-42,65: 4x [Conv2d k4 s2 p1 + SiLU].  obs (N,3,H,W) -> (N, 4096)."""
    p = WM + "encoder.feature_extractor."
    x = obs
    for i in (0, 2, 4, 6):
        x = F.silu(F.conv2d(x, sd[p + f"{i}.weight"], sd[p + f"{i}.bias"], stride=2, padding=1))
    return x.flatten(1)


def encoder_logits(sd: SD, h: torch.Tensor, obs: torch.Tensor, R: int = 32, C: int = 32) -> torch.Tensor:
    """VariationalAutoEncoder.py:57-75: input order [features, h].  h (N,D), obs (N,3,H,W) -> (N,R,C)."""
    p = WM + "encoder.latent_mapper."
    x = torch.cat([encoder_features(sd, obs), h], dim=-1)
    x = F.linear(x, sd[p + "0.weight"], sd[p + "0.bias"])
    x = F.silu(F.layer_norm(x, (x.shape[-1],), sd[p + "1.weight"], sd[p + "1.bias"], 1e-5))
    x = F.linear(x, sd[p + "3.weight"], sd[p + "3.bias"])
    return x.reshape(-1, R, C)


def decoder_forward(sd: SD, h: torch.Tensor, z: torch.Tensor, hw: Tuple[int, int] = (64, 64)) -> torch.Tensor:
    """VariationalAutoEncoder.py:139-161: input order [h, z].  -> (N,3,H,W) in (-1,1)."""
    p = WM + "decoder."
    x = torch.cat([h, z.reshape(z.shape[0], -1)], dim=-1)
    x = F.linear(x, sd[p + "upscaler.0.weight"], sd[p + "upscaler.0.bias"])
    x = F.silu(F.layer_norm(x, (x.shape[-1],), sd[p + "upscaler.1.weight"], sd[p + "upscaler.1.bias"], 1e-5))
    x = F.silu(F.linear(x, sd[p + "upscaler.3.weight"], sd[p + "upscaler.3.bias"]))
    c0 = sd[p + "image_builder.0.weight"].shape[0]
    x = x.reshape(-1, c0, hw[0] // 16, hw[1] // 16)
    for i in (0, 2, 4):
        x = F.silu(F.conv_transpose2d(x, sd[p + f"image_builder.{i}.weight"], sd[p + f"image_builder.{i}.bias"], stride=2, padding=1))
    x = torch.tanh(F.conv_transpose2d(x, sd[p + "image_builder.6.weight"], sd[p + "image_builder.6.bias"], stride=2, padding=1))
    return x


# --------------------------------------------------------------------------------------
# observe: WorldModel.py:79-146, Dreamer.py:244-262
# --------------------------------------------------------------------------------------
def observe_step(sd: SD, z: torch.Tensor, h: torch.Tensor, a: torch.Tensor, obs: torch.Tensor, u: torch.Tensor,
                 margin_frac: float = 0.0, delta: float = 0.0):
    """WorldModel.observe_step (WorldModel.py:79-82).  obs (B,3,H,W) already in [-0.5, 0.5]."""
    h2 = gru_step(sd, z, h, a)
    logits = encoder_logits(sd, h2, obs, z.shape[-2], z.shape[-1])
    if margin_frac > 0.0 or delta > 0.0:
        u = interior_uniforms(unimix_probs(logits), u, margin_frac, delta)
    z2, idx, _ = categorical_st(logits, u)
    return z2, h2, logits, idx, u


def observe_scan(sd: SD, obs: torch.Tensor, act: torch.Tensor, uniforms: torch.Tensor,
                 margin_frac: float = 0.0, delta: float = 0.0):
    """The posterior scan of WorldModel.unroll_model (WorldModel.py:92-111).

    obs (B,T,3,H,W) normalised, act (B,T,A), uniforms (T,B,R).  t = 0 runs a GRU step on all-zero
    (h, z, a) before encoding frame 0, exactly like the reference.
    Returns latent (B,T,R,C), hidden (B,T,D), post logits (B,T,R,C), idx (B,T,R), uniforms used.
    """
    B, T = obs.shape[:2]
    R = uniforms.shape[-1]
    D = sd[WM + "sequence_model.GRU.weight_hh"].shape[1]
    C = sd[WM + "encoder.latent_mapper.3.weight"].shape[0] // R
    A = act.shape[-1]
    h = torch.zeros(B, D)
    z = torch.zeros(B, R, C)
    Zs, Hs, Ls, Is, Us = [], [], [], [], []
    for t in range(T):
        a = act[:, t - 1] if t > 0 else torch.zeros(B, A)
        z, h, lg, idx, u = observe_step(sd, z, h, a, obs[:, t], uniforms[t], margin_frac, delta)
        Zs.append(z); Hs.append(h); Ls.append(lg); Is.append(idx); Us.append(u)
    st = lambda xs: torch.stack(xs, dim=1)
    return st(Zs), st(Hs), st(Ls), st(Is), torch.stack(Us, 0)


def warm_start(sd: SD, obs: torch.Tensor, act: torch.Tensor, uniforms: torch.Tensor, warmup: int,
               margin_frac: float = 0.0, delta: float = 0.0):
    """Dreamer.warm_start_generator (Dreamer.py:244-262): frame 0 is encoded with h = 0 and NO GRU
    step, then ``warmup - 1`` observe steps.  obs (B,T,3,H,W) raw 0..255, act (B,T,A),
    uniforms (warmup,B,R).  Returns (z (B,1,R,C), h (B,1,D), uniforms used).
    """
    obs = obs.float() / 255.0 - 0.5
    B = obs.shape[0]
    R = uniforms.shape[-1]
    D = sd[WM + "sequence_model.GRU.weight_hh"].shape[1]
    C = sd[WM + "encoder.latent_mapper.3.weight"].shape[0] // R
    h = torch.zeros(B, D)
    lg = encoder_logits(sd, h, obs[:, 0], R, C)
    u0 = uniforms[0]
    if margin_frac > 0.0 or delta > 0.0:
        u0 = interior_uniforms(unimix_probs(lg), u0, margin_frac, delta)
    z, _, _ = categorical_st(lg, u0)
    Us = [u0]
    for t in range(1, warmup):
        z, h, _, _, u = observe_step(sd, z, h, act[:, t - 1], obs[:, t], uniforms[t], margin_frac, delta)
        Us.append(u)
    return z.unsqueeze(1), h.unsqueeze(1), torch.stack(Us, 0)


def unroll_model(sd: SD, obs: torch.Tensor, act: torch.Tensor, rew: torch.Tensor, cont: torch.Tensor,
                 uniforms: torch.Tensor, margin_frac: float = 0.0, delta: float = 0.0):
    """WorldModel.unroll_model (WorldModel.py:84-146).  obs normalised (B,T,3,H,W); T = horizon.

    Returns (prior_logits[:,1:], post_logits[:,1:], obs_ll[:,1:], rew_ll, cont_bce) plus extras
    (latent, hidden, idx, uniforms used).
    """
    B, T = obs.shape[:2]
    z, h, post, idx, used = observe_scan(sd, obs, act, uniforms, margin_frac, delta)
    R, C = z.shape[-2:]
    hf = h.reshape(B * T, -1)
    zf = z.reshape(B * T, R, C)
    prior = prior_logits(sd, hf, R, C).reshape(B, T, R, C)
    dec = decoder_forward(sd, hf, zf, obs.shape[-2:]).reshape(obs.shape)
    h1 = h[:, 1:].reshape(B * (T - 1), -1)
    z1 = z[:, 1:].reshape(B * (T - 1), R, C)
    rl = reward_logits(sd, h1, z1).reshape(B, T - 1, -1)
    cl = continue_logit(sd, h1, z1).reshape(B, T - 1, 1)
    obs_ll = -((dec - obs) ** 2).sum(dim=[-3, -2, -1])
    cont_bce = F.binary_cross_entropy_with_logits(cl, cont[:, :T - 1], reduction="none")
    rew_ll = twohot_ce(rl, rew[:, :T - 1], sd[WM + "reward_predictor.buckets_rew"])
    return (prior[:, 1:], post[:, 1:], obs_ll[:, 1:], rew_ll, cont_bce), (z, h, idx, used, dec)


def world_model_loss(sd: SD, obs_raw: torch.Tensor, act: torch.Tensor, rew: torch.Tensor, cont: torch.Tensor,
                     uniforms: torch.Tensor, horizon: int, betas=(1.0, 0.5, 0.1),
                     margin_frac: float = 0.0, delta: float = 0.0):
    """Forward value of WorldModel.training_step's loss (WorldModel.py:148-188), fp32 throughout.

    obs_raw (B,L,3,H,W) in 0..255.  Returns (total, dict of parts, extras).
    """
    obs = obs_raw.float() / 255.0 - 0.5
    T = horizon
    (prior, post, obs_ll, rew_ll, cont_bce), extras = unroll_model(
        sd, obs[:, :T], act[:, :T], rew[:, :T], cont[:, :T], uniforms, margin_frac, delta)
    mask = cont[:, :T - 1]
    obs_ll = obs_ll * mask.squeeze(-1)
    rew_ll = rew_ll * mask
    cont_bce = cont_bce * mask
    kl = categorical_kl_terms(post, prior)
    kl_mean = (kl * mask.squeeze(-1)).mean()
    denom = mask.sum() + 1e-5
    loss_pred = (-obs_ll.sum() - rew_ll.sum() + cont_bce.sum()) / denom
    loss_dyn = torch.maximum(torch.tensor(1.0), kl_mean)
    loss_rep = torch.maximum(torch.tensor(1.0), kl_mean)
    total = betas[0] * loss_pred + betas[1] * loss_dyn + betas[2] * loss_rep
    parts = dict(loss_pred=loss_pred, kl_mean=kl_mean, obs_ll_sum=obs_ll.sum(), rew_ll_sum=rew_ll.sum(),
                 cont_bce_sum=cont_bce.sum(), denom=denom)
    return total, parts, extras


# --------------------------------------------------------------------------------------
# Agent: Agent.py:78-172
# --------------------------------------------------------------------------------------
def lambda_returns(rew: torch.Tensor, cont: torch.Tensor, value: torch.Tensor, gamma: float, lam: float) -> torch.Tensor:
    """Agent.compute_batched_R_lambda_returns' reverse scan (Agent.py:158-171).

    rew, cont (B,H,1); value (B,H+1,1) -> (B,H,1).
    """
    H = rew.shape[1]
    nxt = rew[:, -1] + gamma * cont[:, -1] * value[:, -1]
    out = [nxt]
    for t in reversed(range(H - 1)):
        nxt = rew[:, t] + gamma * cont[:, t] * ((1 - lam) * value[:, t + 1] + lam * nxt)
        out.insert(0, nxt)
    return torch.stack(out, dim=1)


def tanh_normal_log_prob(a: torch.Tensor, mu: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
    """log-prob of a tanh-squashed Normal, summed over the action dim (Agent.py:110-115).

    Equals Normal(mu, sigma).log_prob(atanh a) - 2 (log 2 - y - softplus(-2y)), y = atanh a.
    """
    a = a.clamp(-1.0 + 1e-6, 1.0 - 1e-6)
    y = torch.atanh(a)
    base = -((y - mu) ** 2) / (2 * sigma ** 2) - torch.log(sigma) - 0.5 * math.log(2 * math.pi)
    ldj = 2.0 * (math.log(2.0) - y - F.softplus(-2.0 * y))
    return (base - ldj).sum(-1)


def quantile_linear(x: torch.Tensor, q: float) -> torch.Tensor:
    """torch.quantile's default linear interpolation on a flat tensor (Agent.py:83-84)."""
    s, _ = torch.sort(x.flatten())
    pos = q * (s.numel() - 1)
    lo = int(math.floor(pos))
    hi = min(lo + 1, s.numel() - 1)
    return s[lo] + (s[hi] - s[lo]) * (pos - lo)


def agent_losses(sd: SD, z: torch.Tensor, h: torch.Tensor, rew: torch.Tensor, cont: torch.Tensor,
                 act: torch.Tensor, mu: torch.Tensor, sigma: torch.Tensor, S: float,
                 gamma: float = 0.99, lam: float = 0.95, nu: float = 3e-4, smoothing: float = 0.99):
    """Forward values of Agent.train_step (Agent.py:96-135).

    z (B,H+1,R,C), h (B,H+1,D), rew/cont (B,H,1), act/mu/sigma (B,H,A).
    Returns dict(loss_actor, loss_critic, returns, S_new, values, log_prob).
    """
    B, H1 = h.shape[:2]
    hf = h.reshape(B * H1, -1)
    zf = z.reshape(B * H1, *z.shape[-2:])
    v_tgt = critic_value(sd, hf, zf, "target_critic").reshape(B, H1, 1)
    R = lambda_returns(rew, cont, v_tgt, gamma, lam)
    v = critic_value(sd, hf, zf, "critic").reshape(B, H1, 1)
    adv = (R - v[:, :-1]).squeeze(-1)
    logp = tanh_normal_log_prob(act, mu, sigma)
    rng = torch.maximum(quantile_linear(R, 0.95) - quantile_linear(R, 0.05), torch.tensor(1.0))
    S_new = smoothing * S + (1.0 - smoothing) * rng
    norm = torch.maximum(torch.as_tensor(S_new, dtype=torch.float32), torch.tensor(1.0))
    loss_actor = (-(logp * (adv / norm)) - nu * (-logp)).mean()
    cl = critic_logits(sd, hf, zf, "critic").reshape(B, H1, -1)[:, :-1]
    ce = -twohot_ce(cl, symlog(R), sd[AG + "critic.buckets_crit"]).squeeze(-1)
    return dict(loss_actor=loss_actor, loss_critic=ce.mean(), returns=R, S_new=S_new, values=v,
                target_values=v_tgt, log_prob=logp, advantage=adv)
