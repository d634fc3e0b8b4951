"""WorldModel (WorldModel.py) and Agent (Agent.py) mirrors: same constructors, attributes and methods.

Forward quantities (states, logits, log-likelihoods, returns, losses) come from the sm_100a kernels.
The world model's gradient comes from bptt.world_model_backward: a hand-scheduled back-propagation through time on the
trajectory the scan kernels produced (batched heads / decoder MLP / KL part differentiated by hand, 7 launches per time step for the
recurrence, batched weight-gradient GEMMs -- all on the library's drm_gemm_tf32 and elementwise backward kernels; the conv stacks'
backward is torch autograd); the Agent's from bptt.actor_backward / bptt.critic_backward.  ``_tail_world_model`` below -- a torch autograd graph over the Python scan, teacher-forced on the
sampled classes -- is the reference implementation of that gradient (``grad_mode = "autograd"``) and the Agent's tail.  Clipping, AdamW, the target-critic EMA and gradient zeroing run as the fused
flat-bucket kernels of optim.FlatAdamW (section 8f rank 2).
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import bptt
from . import dist as D
from . import ops
from .graphs import StepGraph
from .optim import FlatAdamW, make_adamw
from .modules import (Actor, ContinuePredictor, Critic, Decoder, DynamicsPredictor, Encoder, RewardPredictor, SequenceModel,
                      _Packed, _VaeEngine, symlog)


class WorldModel(nn.Module):
    def __init__(self, hidden_dims, latent_dims, observation_dims, action_dims, training_horizon, batch_size, WM_lr, WM_betas,
                 WM_eps, beta_pred, beta_dyn, beta_rep, num_encoder_filters_1, num_encoder_filters_2, encoder_hidden_layer_nodes,
                 num_decoder_filters_1, num_decoder_filters_2, decoder_hidden_layer_nodes, dyn_pred_hidden_num_nodes_1,
                 dyn_pred_hidden_num_nodes_2, rew_pred_hidden_num_nodes_1, rew_pred_hidden_num_nodes_2, reward_buckets,
                 cont_pred_hidden_num_nodes_1, cont_pred_hidden_num_nodes_2, device='cpu'):
        super().__init__()
        self.latent_num_rows, self.latent_num_columns = latent_dims
        self.hidden_dims = hidden_dims
        self.action_dims = action_dims
        self.observation_dim_x, self.observation_dim_y = observation_dims
        self.horizon = training_horizon
        self.buckets = reward_buckets
        self.beta_pred, self.beta_dyn, self.beta_rep = beta_pred, beta_dyn, beta_rep
        self.batch_size = batch_size
        R, C = latent_dims
        self.encoder = Encoder(observation_dims, hidden_dims, R, C, num_encoder_filters_1, num_encoder_filters_2, encoder_hidden_layer_nodes, device=device)
        self.sequence_model = SequenceModel(R, C, hidden_dims, action_dims, num_layers=1, device=device)
        self.dynamics_predictor = DynamicsPredictor(R, C, hidden_dims, dyn_pred_hidden_num_nodes_1, dyn_pred_hidden_num_nodes_2, device)
        self.reward_predictor = RewardPredictor(R, C, hidden_dims, rew_pred_hidden_num_nodes_1, rew_pred_hidden_num_nodes_2, reward_buckets, device=device)
        self.continue_predictor = ContinuePredictor(R, C, hidden_dims, cont_pred_hidden_num_nodes_1, cont_pred_hidden_num_nodes_2, device=device)
        self.decoder = Decoder(R, C, observation_dims, hidden_dims, num_decoder_filters_1, num_decoder_filters_2, decoder_hidden_layer_nodes, device=device)
        self.device = device
        # clip(100) + AdamW(wd 1e-6) fused on one flat 31 MB bucket (WorldModel.py:46,198-199)
        self.optimiser = make_adamw(self.parameters(), WM_lr, (WM_betas[0], WM_betas[1]), WM_eps, weight_decay=1e-6, max_norm=100.0)
        self.scalar = torch.amp.GradScaler(enabled=False)   # kept for API compatibility: the tail runs in fp32/TF32, no loss scaling
        object.__setattr__(self, "_actor", None)
        object.__setattr__(self, "_engine", _VaeEngine(self._engine_sd, R, C, hidden_dims, action_dims, tuple(observation_dims)))
        self.encoder._bind(self)
        self.decoder._bind(self)
        self.last = {}

    # ---- engine plumbing -------------------------------------------------------------------
    def _engine_sd(self):
        sd = {"world_model." + k: v for k, v in self.state_dict(keep_vars=True).items()}
        if self._actor is not None:
            sd.update({"agent.actor." + k: v for k, v in self._actor.state_dict(keep_vars=True).items()})
        return sd

    def attach_actor(self, actor):
        """Pack the policy next to the world model so Dreamer.dream_episodes runs as one fused rollout."""
        if self._actor is not actor:
            object.__setattr__(self, "_actor", actor)
            self._engine.versions = None

    # ---- WorldModel.py:72-82 ---------------------------------------------------------------
    def imagine_step(self, hidden_state, latent_state, action, uniforms=None):
        B = hidden_state.shape[0]
        ws = self._engine.rollout(max(128, 1 << (B - 1).bit_length()), 1)
        h2 = ws.gru_step(latent_state.reshape(B, -1), hidden_state.reshape(B, -1), action.reshape(B, -1))
        if uniforms is None:
            uniforms = torch.rand(B, self.latent_num_rows, device=h2.device)
        pr = ws.prior(h2, uniforms.reshape(B, -1), want_logits=False)
        hd = ws.heads(h2, pr["z"], L.HEAD_REWARD | L.HEAD_CONT)
        return h2.unsqueeze(1), pr["z"].unsqueeze(1), hd["reward"].unsqueeze(1), hd["cont_prob"].unsqueeze(1)

    def observe_step(self, last_latent, last_hidden, last_action, observation, uniforms=None):
        hidden_state = self.sequence_model.forward(last_latent, last_hidden, last_action)
        latent_state, latent_logits = self.encoder.encode(hidden_state, observation, uniforms)
        return latent_state, hidden_state, latent_logits

    # ---- WorldModel.py:84-146 --------------------------------------------------------------
    def unroll_model(self, observation_sequence_batch, action_sequence_batch, reward_sequence_batch, continue_sequence_batch, uniforms=None):
        B = continue_sequence_batch.shape[0]
        T = self.horizon
        obs = L.f32c(observation_sequence_batch[:, :T])
        dev = obs.device
        if uniforms is None:
            uniforms = torch.rand(T, B, self.latent_num_rows, device=dev)
        ws = self._engine.observe(B, T)
        sc = ws.scan(obs, action_sequence_batch[:, :T], uniforms)
        hd = ws.heads()
        obs_ll = ops.neg_sse_rows(hd["dec_mu"], obs)
        rew_ll = ops.twohot_ce(hd["reward_logits"], L.f32c(reward_sequence_batch[:, :T - 1]), self.reward_predictor.buckets_rew)
        x, y = hd["cont_logit"], continue_sequence_batch[:, :T - 1]
        cont_bce = torch.clamp(x, min=0) - x * y + torch.log1p(torch.exp(-torch.abs(x)))   # BCE-with-logits, elementwise on (B,T-1,1)
        self.last = dict(scan=sc, heads=hd, uniforms=uniforms)
        return hd["prior_logits"][:, 1:], sc["logits"][:, 1:], obs_ll[:, 1:], rew_ll, cont_bce

    def loss_forward(self, observation_sequences, action_sequences, reward_sequences, continue_sequences, uniforms=None):
        """The loss of WorldModel.training_step (WorldModel.py:156-188) on the kernels; returns (total, parts)."""
        T = self.horizon
        obs = (L.f32c(observation_sequences)[:, :T] / 255.0) - 0.5
        prior, post, obs_ll, rew_ll, cont_ll = self.unroll_model(obs, action_sequences[:, :T], reward_sequences[:, :T], continue_sequences[:, :T], uniforms)
        mask = continue_sequences[:, :T - 1]
        m1 = mask.squeeze(-1)
        kl = ops.categorical32_kl(post, prior)                      # forward value of both KL terms (they differ only in stop-gradients)
        # per-rank sums -> one packed all-reduce -> the GLOBAL loss on every rank (a no-op for a single process)
        local = torch.stack([(obs_ll * m1).sum(), (rew_ll * mask).sum(), (cont_ll * mask).sum(), mask.sum(), (kl * m1).sum(),
                             torch.full((), float(kl.numel()), device=kl.device)])
        total, parts = D.world_model_loss_from_sums(local, (self.beta_pred, self.beta_dyn, self.beta_rep))
        parts["obs_norm"] = obs
        return total, parts

    # ---- WorldModel.py:148-202 -------------------------------------------------------------
    def training_step(self, observation_sequences, action_sequences, reward_sequences, continue_sequences, uniforms=None):
        graphs = self.__dict__.get("_graphs")
        if graphs is not None:                 # enable_cuda_graphs(): eager warm-up calls, then one captured step replayed
            if uniforms is None:
                uniforms = torch.rand(self.horizon, continue_sequences.shape[0], self.latent_num_rows, device=continue_sequences.device)
            return graphs(L.f32c(observation_sequences), L.f32c(action_sequences), L.f32c(reward_sequences),
                          L.f32c(continue_sequences), uniforms).clone()
        if self._fused_step_ok():
            return self._scan_bptt_step(observation_sequences, action_sequences, reward_sequences, continue_sequences, uniforms, host_check=True)
        total, parts = self.loss_forward(observation_sequences, action_sequences, reward_sequences, continue_sequences, uniforms)
        if D.any_rank_flag(bool(torch.isnan(total) or torch.isinf(total)), total.device):
            print("World Model loss is nan or inf, skipping update.")
            return total
        return self._backward_and_step(total, parts, action_sequences, reward_sequences, continue_sequences)

    def _fused_step_ok(self):
        return (self.__dict__.get("grad_mode", "bptt") == "bptt" and self.latent_num_columns == 32 and isinstance(self.optimiser, FlatAdamW))

    def _scan_bptt_step(self, observation_sequences, action_sequences, reward_sequences, continue_sequences, uniforms, host_check):
        """The training step without evaluating anything twice: the posterior scan runs on the kernels (it fixes the trajectory:
        hidden states and sampled classes), then bptt.world_model_backward evaluates decoder / heads / KL ONCE in the batched graph
        it differentiates, reduces the loss sums over ranks itself and returns the global loss (WorldModel.py:156-202)."""
        T = self.horizon
        B = continue_sequences.shape[0]
        obs = (L.f32c(observation_sequences)[:, :T] / 255.0) - 0.5
        if uniforms is None:
            uniforms = torch.rand(T, B, self.latent_num_rows, device=obs.device)
        sc = self._engine.observe(B, T).scan(obs, action_sequences[:, :T], uniforms)
        self.last = dict(scan=sc, uniforms=uniforms)
        self.optimiser.zero_grad()
        on_loss = None
        if host_check:
            def on_loss(total):
                bad = D.any_rank_flag(bool(torch.isnan(total) or torch.isinf(total)), total.device)
                if bad:
                    print("World Model loss is nan or inf, skipping update.")
                return not bad
        total, done = bptt.world_model_backward(self, obs, L.f32c(action_sequences)[:, :T], L.f32c(reward_sequences)[:, :T],
                                                L.f32c(continue_sequences)[:, :T], sc["idx"], sc["hidden"], parts="reduce", on_loss=on_loss)
        if done:
            self.optimiser.step(all_reduce=True)
        self.last["tail_loss"] = total.detach()
        return total

    def _backward_and_step(self, total, parts, action_sequences, reward_sequences, continue_sequences):
        T = self.horizon
        self.optimiser.zero_grad()
        if self.__dict__.get("grad_mode", "bptt") == "bptt" and self.latent_num_columns == 32:
            # hand-scheduled BPTT on the scan's own trajectory: batched non-recurrent graph + 7 launches per time step (bptt.py)
            tail = bptt.world_model_backward(self, parts["obs_norm"], action_sequences[:, :T], reward_sequences[:, :T],
                                             continue_sequences[:, :T], self.last["scan"]["idx"], self.last["scan"]["hidden"], parts)
        else:                # reference implementation of the gradient: autograd over the Python scan (~70 launches per time step)
            tail = _tail_world_model(self, parts["obs_norm"], action_sequences[:, :T], reward_sequences[:, :T], continue_sequences[:, :T],
                                     self.last["scan"]["idx"], parts)
            tail.backward()
        if isinstance(self.optimiser, FlatAdamW):
            # gradients of the global loss = SUM over ranks of the per-rank tails: all-reduce of the flat 31 MB bucket the
            # gradients were accumulated into, then clip + AdamW + zeroing in three launches
            self.optimiser.step(all_reduce=True)
        else:                # parameters moved to the GPU after construction: stock torch optimiser
            if D.is_dist():
                if self.__dict__.get("_bucket") is None:
                    self.__dict__["_bucket"] = D.FlatBucket(self.parameters())
                self.__dict__["_bucket"].all_reduce()
            nn.utils.clip_grad_norm_(self.parameters(), 100.0)
            self.optimiser.step()
        self.last["tail_loss"] = tail.detach()
        return total

    def _step_body(self, obs, act, rew, cont, uniforms):
        """One whole training step with no host synchronisation (capturable).  The reference's host-side NaN/Inf check
        (WorldModel.py:191) becomes the optimiser's device-side skip: a non-finite loss gives a non-finite gradient norm."""
        if self._fused_step_ok():
            return self._scan_bptt_step(obs, act, rew, cont, uniforms, host_check=False)
        total, parts = self.loss_forward(obs, act, rew, cont, uniforms)
        return self._backward_and_step(total, parts, act, rew, cont)

    def enable_cuda_graphs(self, warmup: int = 3):
        """Replay training_step as one CUDA graph per input shape after `warmup` eager steps (graphs.StepGraph)."""
        if not isinstance(self.optimiser, FlatAdamW):
            raise RuntimeError("enable_cuda_graphs needs the flat fused optimiser (construct the WorldModel on the GPU)")

        def before_capture():
            self._engine.versions = None           # the captured step must contain the weight re-pack

        self.__dict__["_graphs"] = StepGraph(self._step_body, warmup, before_capture, self.optimiser.mark_updated)
        return self


def _st_latent(logits, idx, C):
    """straight-through one-hot on given indices (DynamicsPredictors.py:33-39 with the kernel's draw)"""
    if C == 32:
        return ops.straight_through(logits, idx)           # one kernel each way
    p = 0.99 * torch.softmax(logits.float(), dim=-1) + 0.01 / C
    return F.one_hot(idx.long(), C).float() + p - p.detach()


def _cat_kl(lp_logits, lq_logits):
    lp = F.log_softmax(lp_logits.float(), -1)
    lq = F.log_softmax(lq_logits.float(), -1)
    return (lp.exp() * (lp - lq)).sum(-1).sum(-1)


def _tail_world_model(wm: WorldModel, obs, act, rew, cont, idx, parts=None):
    """Differentiable restatement of WorldModel.training_step's loss on the kernels' sampled indices (gradients only).

    With `parts` (the globally reduced denominators) the value is this rank's additive share of the global loss, so the
    SUM of the per-rank gradients is the gradient of the global loss; on one process it equals the loss itself."""
    B, T = obs.shape[:2]
    R, C, Dh = wm.latent_num_rows, wm.latent_num_columns, wm.hidden_dims
    dev = obs.device
    feats = wm.encoder.feature_extractor(obs.reshape(B * T, *obs.shape[2:])).flatten(1).view(B, T, -1)
    h = torch.zeros(B, Dh, device=dev)
    z = torch.zeros(B, R * C, device=dev)
    hs, zs, post = [], [], []
    for t in range(T):
        a = act[:, t - 1] if t > 0 else torch.zeros(B, wm.action_dims, device=dev)
        h = wm.sequence_model.GRU(torch.cat([z, a], -1), h)
        lg = wm.encoder.latent_mapper(torch.cat([feats[:, t], h], -1)).view(B, R, C)
        z = _st_latent(lg, idx[:, t], C).reshape(B, R * C)
        hs.append(h); zs.append(z); post.append(lg)
    hseq, zseq, post = torch.stack(hs, 1), torch.stack(zs, 1), torch.stack(post, 1)
    prior = wm.dynamics_predictor.logit_net(hseq).view(B, T, R, C)
    hz = torch.cat([hseq, zseq], -1)
    x = wm.decoder.upscaler(hz.reshape(B * T, -1)).view(B * T, wm.decoder.num_filters_start, wm.decoder.start_height, wm.decoder.start_width)
    dec = wm.decoder.image_builder(x).view(obs.shape)
    rl = wm.reward_predictor.logit_net(hz[:, 1:])
    cl = wm.continue_predictor.logit_generator(hz[:, 1:])
    mask = cont[:, :T - 1]
    m1 = mask.squeeze(-1)
    obs_ll = -((dec.float() - obs) ** 2).sum(dim=[-3, -2, -1])[:, 1:] * m1
    b = wm.reward_predictor.buckets_rew
    v = torch.maximum(torch.minimum(rew[:, :T - 1], b[-1]), b[0])       # clamp to the (sorted) bucket range, no host read
    lo = torch.clamp(torch.searchsorted(b, v.contiguous(), right=True) - 1, max=len(b) - 2)
    w = (v - b[lo]) / (b[lo + 1] - b[lo] + 1e-8)
    lsm = F.log_softmax(rl, -1)
    rew_ll = ((1 - w) * lsm.gather(-1, lo) + w * lsm.gather(-1, lo + 1)) * mask
    cont_ll = F.binary_cross_entropy_with_logits(cl, mask, reduction='none') * mask
    dyn_sum = (_cat_kl(post[:, 1:].detach(), prior[:, 1:]) * m1).sum()
    rep_sum = (_cat_kl(post[:, 1:], prior[:, 1:].detach()) * m1).sum()
    if parts is None:
        denom, n_el, kl_mean = mask.sum() + 1e-5, float(m1.numel()), None
    else:
        denom, n_el, kl_mean = parts["denom"], parts["n_elements"], parts["kl_mean"]
    loss_pred = (-obs_ll.sum() - rew_ll.sum() + cont_ll.sum()) / denom
    dyn, rep = dyn_sum / n_el, rep_sum / n_el
    if kl_mean is None:
        one = torch.ones((), device=dev)
        return wm.beta_pred * loss_pred + wm.beta_dyn * torch.maximum(one, dyn) + wm.beta_rep * torch.maximum(one, rep)
    # free bits on the GLOBAL mean: below 1 the KL terms are the constant 1 (no gradient), above it they are linear in the sums
    live = (kl_mean > 1.0).to(loss_pred.dtype)               # device-side: no host read of the global mean
    const = (1.0 - live) * ((wm.beta_dyn + wm.beta_rep) / D.world())
    return wm.beta_pred * loss_pred + live * (wm.beta_dyn * dyn + wm.beta_rep * rep) + const


class Agent(nn.Module):
    def __init__(self, action_dim, latent_dims, hidden_state_dim, HL_A1, HL_A2, HL_C1, HL_C2, critic_buckets, A_lr, A_betas, A_eps,
                 C_lr, C_betas, C_eps, nu, lambda_, gamma, *, device='cpu'):
        super().__init__()
        self.device = device
        self.actor = Actor(action_dim, latent_dims[0], latent_dims[1], hidden_state_dim, HL_A1, HL_A2, device=device)
        self.critic = Critic(latent_dims[0], latent_dims[1], hidden_state_dim, HL_C1, HL_C2, critic_buckets, device=device)
        self.target_critic = copy.deepcopy(self.critic)
        for p in self.target_critic.parameters():
            p.requires_grad = False
        self.nu, self.lambda_, self.gamma = nu, lambda_, gamma
        self.buckets = critic_buckets
        self.S = 1.0
        self.smoothing_factor = 0.99
        # Agent.py:30-31,147-153: clip(100) + AdamW per group; the critic's pass also applies the target EMA (Agent.py:90-94)
        self.actor_optimiser = make_adamw(self.actor.parameters(), A_lr, (A_betas[0], A_betas[1]), A_eps, weight_decay=1e-6, max_norm=100.0)
        self.critic_optimiser = make_adamw(self.critic.parameters(), C_lr, (C_betas[0], C_betas[1]), C_eps, weight_decay=1e-6, max_norm=100.0,
                                           ema_params=self.target_critic.parameters(), tau=0.02)
        object.__setattr__(self, "_pk", _Packed(self, "agent.", latent_dims[0], latent_dims[1], hidden_state_dim, action_dim))
        self.last = {}

    def _heads(self, h, z, heads, want_logits=False):
        B, S = h.shape[:2]
        ws = self._pk.workspace(B * S)
        return ws.heads(h.reshape(B * S, -1), z.reshape(B * S, -1), heads, want_logits=want_logits)

    def update_S(self, lambda_returns):
        """Agent.py:78-88 (percentiles by sort + linear interpolation, the definition torch.quantile uses).  S lives on the
        device and is updated in place; a non-finite return set leaves it unchanged (Agent.py:80-81) -- decided on the device."""
        flat = D.all_gather_cat(lambda_returns.detach().flatten())     # percentiles over the GLOBAL return set
        if flat.is_cuda:
            q = ops.percentile_pair(flat, 0.05, 0.95)                  # radix select, one launch (drm_percentile_pair)
            ok = q[2] > 0.5
            rng = torch.maximum(q[1] - q[0], torch.ones((), device=flat.device))
        else:
            ok = torch.isfinite(flat).all()
            s, _ = torch.sort(flat)
            n = s.numel()

            def q(p):
                pos = p * (n - 1)
                lo = int(pos)
                hi = min(lo + 1, n - 1)
                return s[lo] + (s[hi] - s[lo]) * (pos - lo)

            rng = torch.maximum(q(0.95) - q(0.05), torch.ones((), device=flat.device))
        alpha = 1.0 - self.smoothing_factor
        if not isinstance(self.S, torch.Tensor):
            self.S = torch.full((), float(self.S), dtype=torch.float32, device=flat.device)
        self.S.copy_(torch.where(ok, (1.0 - alpha) * self.S + alpha * rng, self.S))

    def soft_update_target(self, tau=0.02):
        with torch.no_grad():
            for pc, pt in zip(self.critic.parameters(), self.target_critic.parameters()):
                pt.data.mul_(1.0 - tau)
                pt.data.add_(tau * pc.data)

    def compute_batched_R_lambda_returns(self, hidden_state_batched_seq, latent_state_batched_seq, reward_batched_seq, continue_batched_seq, seq_length):
        """Agent.py:156-172: target-critic values on the H+1 states, then the reverse lambda scan (one kernel each)."""
        B, H1 = hidden_state_batched_seq.shape[:2]
        v = self._heads(hidden_state_batched_seq, latent_state_batched_seq, L.HEAD_TARGET_CRITIC)["target_value"].view(B, H1, 1)
        return ops.lambda_return(reward_batched_seq, continue_batched_seq, v, self.gamma, self.lambda_)

    def losses_forward(self, z, h, rew, cont, act, mu, sigma):
        """Forward values of Agent.train_step (Agent.py:96-135) on the kernels."""
        B, H1 = h.shape[:2]
        hd = self._heads(h, z, L.HEAD_CRITIC | L.HEAD_TARGET_CRITIC, want_logits=True)
        v_t = hd["target_value"].view(B, H1, 1)
        v = hd["value"].view(B, H1, 1)
        R = ops.lambda_return(rew, cont, v_t, self.gamma, self.lambda_)
        adv = (R - v[:, :-1]).squeeze(-1)
        logp = ops.tanh_normal_logp(act, mu, sigma) if act.is_cuda else _tanh_normal_log_prob(act, mu, sigma)
        self.update_S(R)
        norm = torch.maximum(self.S, torch.ones((), device=R.device))
        ce = -ops.twohot_ce(hd["value_logits"].view(B, H1, -1)[:, :-1], R, self.critic.buckets_crit, apply_symlog=True)
        sums = D.all_reduce_sum_(torch.stack([(-(logp * (adv / norm)) - self.nu * (-logp)).sum(), ce.sum(),
                                              torch.full((), float(logp.numel()), device=R.device)]))
        n_glob = sums[2]
        return dict(loss_actor=sums[0] / n_glob, loss_critic=sums[1] / n_glob, returns=R, values=v, target_values=v_t, advantage=adv,
                    log_prob=logp, norm=norm, n_global=n_glob)

    def train_step(self, z_batch_seq, h_batch_seq, reward_batch_seq, continue_batch_seq, action_batch_seq, a_mu_batch_seq, a_sigma_batch_seq):
        graphs = self.__dict__.get("_graphs")
        if graphs is not None:                 # enable_cuda_graphs(): eager warm-up calls, then one captured step replayed
            la, lc = graphs(*(L.f32c(t) for t in (z_batch_seq, h_batch_seq, reward_batch_seq, continue_batch_seq, action_batch_seq,
                                                  a_mu_batch_seq, a_sigma_batch_seq)))
            return la.clone(), lc.clone()
        f = self.losses_forward(z_batch_seq, h_batch_seq, reward_batch_seq, continue_batch_seq, action_batch_seq, a_mu_batch_seq, a_sigma_batch_seq)
        la, lc = f["loss_actor"], f["loss_critic"]
        self.last = f
        if D.any_rank_flag(bool(torch.isnan(la) or torch.isinf(la) or torch.isnan(lc) or torch.isinf(lc)), la.device):
            print("Agent loss is nan or inf, skipping update.")
            return la, lc
        return self._backward_and_step(f, z_batch_seq, h_batch_seq, action_batch_seq, a_mu_batch_seq, a_sigma_batch_seq)

    def _step_body(self, z, h, rew, cont, act, mu, sigma):
        """One whole Agent.train_step with no host synchronisation (capturable); non-finite losses are skipped by the
        optimisers' device-side check instead of Agent.py:137-139."""
        f = self.losses_forward(z, h, rew, cont, act, mu, sigma)
        self.last = f
        return self._backward_and_step(f, z, h, act, mu, sigma)

    def enable_cuda_graphs(self, warmup: int = 3):
        """Replay train_step as one CUDA graph per input shape after `warmup` eager steps (graphs.StepGraph)."""
        if not (isinstance(self.critic_optimiser, FlatAdamW) and isinstance(self.actor_optimiser, FlatAdamW)):
            raise RuntimeError("enable_cuda_graphs needs the flat fused optimisers (construct the Agent on the GPU)")

        def before_capture():
            self._pk.invalidate()                  # the captured step must contain the weight re-pack

        def after_replay():
            self.critic_optimiser.mark_updated()
            self.actor_optimiser.mark_updated()

        self.__dict__["_graphs"] = StepGraph(self._step_body, warmup, before_capture, after_replay)
        return self

    def attach_world_model(self, world_model):
        """Give train_step access to the world model the rollout came from, so the actor gradient includes the path through
        the imagined states (bptt.actor_backward).  HotPath and dropin.patch_dreamer call this; without it the gradient is the
        detached-state policy gradient."""
        object.__setattr__(self, "_world_model", world_model)
        return self

    def _backward_and_step(self, f, z_batch_seq, h_batch_seq, action_batch_seq, a_mu_batch_seq, a_sigma_batch_seq):
        la, lc = f["loss_actor"], f["loss_critic"]
        # critic: two-hot CE on the kernel's lambda-returns (torch autograd on the batched MLP)
        B, H1 = h_batch_seq.shape[:2]
        hz = torch.cat([h_batch_seq.detach(), z_batch_seq.detach().reshape(B, H1, -1)], -1)
        self.critic_optimiser.zero_grad()
        if self.__dict__.get("grad_mode", "bptt") == "bptt" and h_batch_seq.is_cuda:
            # hand-scheduled batched MLP backward on the library's kernels (two-hot CE backward, LayerNorm-SiLU backward, drm_gemm_tf32)
            bptt.critic_backward(self, z_batch_seq, h_batch_seq, f["returns"], f["n_global"])
        else:       # grad_mode = "autograd": torch autograd on the batched MLP (the implementation critic_backward is tested against)
            lsm = F.log_softmax(self.critic.value_net(hz[:, :-1]), -1)
            b = self.critic.buckets_crit
            tv = torch.maximum(torch.minimum(symlog(f["returns"]), b[-1]), b[0])   # clamp to the (sorted) bucket range, no host read
            lo = torch.clamp(torch.searchsorted(b, tv.contiguous(), right=True) - 1, max=len(b) - 2)
            w = (tv - b[lo]) / (b[lo + 1] - b[lo] + 1e-8)
            ((-((1 - w) * lsm.gather(-1, lo) + w * lsm.gather(-1, lo + 1))).sum() / f["n_global"]).backward()
        self.actor_optimiser.zero_grad()
        wm = self.__dict__.get("_world_model")
        if wm is not None and self.__dict__.get("grad_mode", "bptt") == "bptt" and z_batch_seq.shape[-1] == 32:
            # the reference's actor gradient: through mu_t, sigma_t directly AND through the imagined states' dependence on the
            # earlier reparameterised actions (the states are not detached in Agent.py:110) -- hand-scheduled BPTT over the rollout
            coef = (-(f["advantage"] / f["norm"]) + self.nu) / f["n_global"]
            bptt.actor_backward(self, wm, z_batch_seq, h_batch_seq, action_batch_seq, a_mu_batch_seq, a_sigma_batch_seq, coef)
        else:
            # no world model attached: policy gradient through mu, sigma recomputed on the detached imagined states (omits the
            # through-the-world-model term, about 3 % of the gradient at initialisation)
            base = self.actor.base_net(hz[:, :-1])
            mu_t = self.actor.mu_head(base)
            sg_t = F.softplus(torch.clamp(self.actor.log_sig_head(base), -5.0, 2.0)) + 1e-3
            logp = _tanh_normal_log_prob(action_batch_seq.detach(), mu_t, sg_t)
            ((-(logp * (f["advantage"] / f["norm"])) - self.nu * (-logp)).sum() / f["n_global"]).backward()
        if isinstance(self.critic_optimiser, FlatAdamW) and isinstance(self.actor_optimiser, FlatAdamW):
            # SUM of per-rank shares = gradient of the global means (flat 1.67 MB and 1.47 MB buckets); the critic's pass also
            # moves the target critic: target = 0.98 target + 0.02 updated critic
            self.critic_optimiser.step(all_reduce=True)
            self.actor_optimiser.step(all_reduce=True)
        else:
            if D.is_dist():
                if self.__dict__.get("_buckets") is None:
                    self.__dict__["_buckets"] = (D.FlatBucket(self.critic.parameters()), D.FlatBucket(self.actor.parameters()))
                for bkt in self.__dict__["_buckets"]:
                    bkt.all_reduce()
            torch.nn.utils.clip_grad_norm_(self.critic.parameters(), 100.0)
            torch.nn.utils.clip_grad_norm_(self.actor.parameters(), 100.0)
            self.critic_optimiser.step()
            self.actor_optimiser.step()
            self.soft_update_target()
        return la, lc


def _tanh_normal_log_prob(a, mu, sigma):
    """Agent.py:110-115: log-prob of the clamped action under tanh(Normal(mu, sigma)), summed over the action dim."""
    a = torch.clamp(a, -1.0 + 1e-6, 1.0 - 1e-6)
    y = torch.atanh(a)
    base = -((y - mu) ** 2) / (2 * sigma ** 2) - torch.log(sigma) - 0.9189385332046727
    return (base - 2.0 * (0.6931471805599453 - y - F.softplus(-2.0 * y))).sum(-1)
