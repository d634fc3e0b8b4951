"""Pin the oracle against the reference's own modules and write tests/golden/*.npz.

Run in the BUILD container only (it imports /root/reference, which does not exist on the GPU
box):  ``python -m oracle.make_golden``.

The reference is executed unmodified except for its two RNG draw sites, which are redirected to
host-supplied streams:
  * ``torch.distributions.Categorical.sample``  -> inverse-CDF on queued uniforms
  * ``torch.distributions.Normal.rsample``      -> loc + scale * queued standard normals
and ``torch.autocast`` is made a no-op so the comparison is fp32-to-fp32 (on CPU the reference's
fp16 autocast only changes rounding, WorldModel.py:162).

For every fixture the script first checks  reference == oracle  (indices exactly, floats to 2e-5)
and aborts otherwise; fixtures hold the *reference's* outputs.
"""
from __future__ import annotations

import contextlib
import copy
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

from oracle import rssm as O           # noqa: E402
from oracle import weights as W        # noqa: E402
from oracle.replay import ReplayOracle  # noqa: E402


class Streams:
    """FIFO of uniforms / normals consumed by the patched draw sites."""

    def __init__(self):
        self.u, self.n = [], []


@contextlib.contextmanager
def patched_reference(streams: Streams):
    import torch.distributions as Dst
    orig_cs, orig_nr, orig_ac = Dst.Categorical.sample, Dst.Normal.rsample, torch.autocast

    def cat_sample(self, sample_shape=torch.Size()):
        u = streams.u.pop(0).reshape(self.probs.shape[:-1])
        cdf = torch.cumsum(self.probs, dim=-1)
        return (cdf <= u.unsqueeze(-1)).sum(-1).clamp(max=self.probs.shape[-1] - 1)

    def normal_rsample(self, sample_shape=torch.Size()):
        eps = streams.n.pop(0).reshape(self.loc.shape)
        return self.loc + self.scale * eps

    class NoAutocast(contextlib.nullcontext):
        def __init__(self, *a, **k):
            super().__init__()

    Dst.Categorical.sample, Dst.Normal.rsample, torch.autocast = cat_sample, normal_rsample, NoAutocast
    try:
        yield
    finally:
        Dst.Categorical.sample, Dst.Normal.rsample, torch.autocast = orig_cs, orig_nr, orig_ac


def build_reference(cfg, sd):
    sys.path.insert(0, REF)
    from Dreamer import Dreamer  # reference orchestrator (Dreamer.py:13)
    d = Dreamer(dict(cfg), torch.device("cpu"))
    d.load_state_dict(sd, strict=True)   # proves the 97-key inventory of oracle/weights.py
    return d


def check(name, ref, ora, tol=2e-5, exact=False):
    ref, ora = torch.as_tensor(ref), torch.as_tensor(ora)
    assert ref.shape == ora.shape, (name, ref.shape, ora.shape)
    if exact:
        assert torch.equal(ref, ora), f"{name}: integer mismatch ({(ref != ora).sum().item()} of {ref.numel()})"
        return 0.0
    err = (ref.float() - ora.float()).abs().max().item()
    scale = max(1.0, ref.float().abs().max().item())
    assert err <= tol * scale, f"{name}: max abs err {err} (scale {scale})"
    return err


def golden_rollout(cfg, tag, B, H, seed, margin=0.25, save_full=True):
    sd = W.make_state_dict(cfg, seed=seed)
    ref = build_reference(cfg, sd)
    ref.horizon = H
    z0, h0, u, n = W.rollout_inputs(cfg, B, H, seed=seed + 1)
    with torch.no_grad():
        out = O.dream_episodes(sd, z0, h0, u, n, margin_frac=margin, delta=1e-5)
    used = out[9]
    st = Streams()
    st.u = [used[t] for t in range(H)]
    st.n = [n[t] for t in range(H)]
    with patched_reference(st), torch.no_grad():
        r = ref.dream_episodes(z0, h0)      # Dreamer.py:143-175, unmodified
    assert not st.u and not st.n
    names = ["latent", "hidden", "actions", "rewards", "continues", "mu", "sigma"]
    errs = {k: check(f"{tag}.{k}", r[i], out[i]) for i, k in enumerate(names)}
    idx_ref = r[0][:, 1:].argmax(-1)
    check(f"{tag}.idx", idx_ref, out[7], exact=True)
    if save_full:
        np.savez_compressed(os.path.join(GOLD, f"{tag}.npz"),
                            cfg=json.dumps(cfg), seed=seed, B=B, H=H, margin=margin,
                            uniforms_used=used.numpy(), idx=idx_ref.numpy().astype(np.uint8),
                            latent_last=r[0][:, -1].numpy(), hidden=r[1].numpy(), actions=r[2].numpy(),
                            rewards=r[3].numpy(), continues=r[4].numpy(), mu=r[5].numpy(), sigma=r[6].numpy())
    return errs, r, out


def golden_observe(cfg, tag, B, T, seed, margin=0.25):
    sd = W.make_state_dict(cfg, seed=seed)
    cfg = dict(cfg, horizon=T, batch_size=B)
    ref = build_reference(cfg, sd)
    obs, act, rew, cont, u = W.sequence_inputs(cfg, B, T, seed=seed + 2)
    with torch.no_grad():
        total, parts, extras = O.world_model_loss(sd, obs, act, rew, cont, u, T, margin_frac=margin, delta=1e-5)
        (prior, post, obs_ll, rew_ll, cont_bce), _ = O.unroll_model(sd, obs / 255.0 - 0.5, act, rew, cont, extras[3])
    used = extras[3]
    st = Streams(); st.u = [used[t] for t in range(T)]
    with patched_reference(st), torch.no_grad():
        r = ref.world_model.unroll_model(obs / 255.0 - 0.5, act, rew, cont)   # WorldModel.py:84-146
    assert not st.u
    errs = {}
    for k, a, b in zip(["prior_logits", "post_logits", "obs_ll", "rew_ll", "cont_bce"], r, (prior, post, obs_ll, rew_ll, cont_bce)):
        errs[k] = check(f"{tag}.{k}", a, b, tol=1e-4)
    # the loss through the reference's own training_step (fp32; it also takes an optimiser step)
    wm = copy.deepcopy(ref.world_model)
    st = Streams(); st.u = [used[t] for t in range(T)]
    with patched_reference(st):
        ref_total = wm.training_step(obs, act, rew, cont).detach()            # WorldModel.py:148-202
    errs["total_loss"] = check(f"{tag}.total_loss", ref_total, total, tol=1e-4)
    # warm start (Dreamer.py:244-262)
    wlen = T // 2
    uw = W.sequence_inputs(cfg, B, T, seed=seed + 3)[4][:wlen]
    with torch.no_grad():
        zo, ho, uw_used = O.warm_start(sd, obs, act, uw, wlen, margin_frac=margin, delta=1e-5)
    st = Streams(); st.u = [uw_used[t] for t in range(wlen)]
    with patched_reference(st), torch.no_grad():
        zr, hr = ref.warm_start_generator(obs, act, T)
    errs["warm_h"] = check(f"{tag}.warm_h", hr, ho, tol=1e-4)
    check(f"{tag}.warm_idx", zr.argmax(-1), zo.argmax(-1), exact=True)
    np.savez_compressed(os.path.join(GOLD, f"{tag}.npz"), cfg=json.dumps(cfg), seed=seed, B=B, T=T, margin=margin,
                        uniforms_used=used.numpy(), idx=extras[2].numpy().astype(np.uint8),
                        hidden=extras[1].numpy(), prior_logits=r[0].numpy(), post_logits=r[1].numpy(),
                        obs_ll=r[2].numpy(), rew_ll=r[3].numpy(), cont_bce=r[4].numpy(),
                        total_loss=ref_total.numpy(), kl_mean=parts["kl_mean"].numpy(),
                        warm_uniforms_used=uw_used.numpy(), warm_idx=zr.argmax(-1).numpy().astype(np.uint8),
                        warm_hidden=hr.numpy())
    return errs


def golden_agent(cfg, tag, B, H, seed):
    sd = W.make_state_dict(cfg, seed=seed)
    ref = build_reference(cfg, sd)
    ref.horizon = H
    z0, h0, u, n = W.rollout_inputs(cfg, B, H, seed=seed + 1)
    with torch.no_grad():
        out = O.dream_episodes(sd, z0, h0, u, n, margin_frac=0.25, delta=1e-5)
        res = O.agent_losses(sd, out[0], out[1], out[3], out[4], out[2], out[5], out[6], S=1.0,
                             gamma=cfg["gamma"], lam=cfg["lambda_"], nu=cfg["nu"])
    ag = copy.deepcopy(ref.agent)
    with torch.no_grad():
        R_ref = ag.compute_batched_R_lambda_returns(out[1], out[0], out[3], out[4], H)   # Agent.py:156-172
        v_ref = ag.critic.value(out[1], out[0])
    # mu/sigma must carry a graph for the reference's loss_actor.backward() (Agent.py:145)
    la, lc = ag.train_step(out[0], out[1], out[3], out[4], out[2],
                           out[5].clone().requires_grad_(), out[6].clone().requires_grad_())   # Agent.py:96-154
    errs = dict(returns=check(f"{tag}.returns", R_ref, res["returns"], tol=1e-4),
                values=check(f"{tag}.values", v_ref, res["values"], tol=1e-4),
                loss_actor=check(f"{tag}.loss_actor", la.detach(), res["loss_actor"], tol=1e-4),
                loss_critic=check(f"{tag}.loss_critic", lc.detach(), res["loss_critic"], tol=1e-4),
                S=check(f"{tag}.S", torch.as_tensor(ag.S), torch.as_tensor(res["S_new"]), tol=1e-5))
    # to_twohot edge cases: below, above, exactly on a bucket, the 7.45e-8 centre bucket
    sys.path.insert(0, REF)
    from DreamerUtils import to_twohot as ref_twohot, symlog as ref_symlog, symexp as ref_symexp
    b = sd["agent.critic.buckets_crit"]
    vals = torch.tensor([[-25.0], [-20.0], [-19.99], [0.0], [7.45e-8], [1e-3], [19.9999], [20.0], [31.0], [float(b[10])], [float(b[200])]])
    check(f"{tag}.twohot_edges", ref_twohot(vals, b), O.to_twohot(vals, b), tol=1e-6)
    x = torch.linspace(-30, 30, 241)
    check(f"{tag}.symlog", ref_symlog(x), O.symlog(x), tol=1e-6)
    check(f"{tag}.symexp", ref_symexp(x), O.symexp(x), tol=1e-6)
    np.savez_compressed(os.path.join(GOLD, f"{tag}.npz"), cfg=json.dumps(cfg), seed=seed, B=B, H=H,
                        uniforms_used=out[9].numpy(), returns=R_ref.numpy(), values=v_ref.numpy(),
                        loss_actor=la.detach().numpy(), loss_critic=lc.detach().numpy(), S=np.float32(ag.S),
                        twohot_vals=vals.numpy(), twohot=ref_twohot(vals, b).numpy())
    return errs


def grad_digest(named, prefix, out, n=1024):
    """Compact, elementwise-checkable digest of a set of tensors: every element of small tensors, a fixed random sample of
    `n` elements of large ones (indices from a crc32(key)-seeded stream), plus each tensor's L2 norm."""
    import zlib
    for key, t in named:
        t = t.detach().reshape(-1).double()
        if t.numel() <= n:
            ix = np.arange(t.numel())
        else:
            ix = np.sort(np.random.RandomState(zlib.crc32(key.encode()) & 0x7FFFFFFF).choice(t.numel(), n, replace=False))
        out[f"{prefix}::{key}::idx"] = ix.astype(np.int32)
        out[f"{prefix}::{key}::val"] = t.numpy()[ix].astype(np.float32)
        out[f"{prefix}::{key}::norm"] = np.float64(t.norm().item())


@contextlib.contextmanager
def capture_preclip_grads(store):
    """The reference clips right after backward (WorldModel.py:198, Agent.py:147-148): snapshot every parameter's gradient on
    entry to clip_grad_norm_, i.e. the raw autograd result."""
    orig = torch.nn.utils.clip_grad_norm_

    def hook(parameters, max_norm, *a, **k):
        params = list(parameters)
        store.append([None if p.grad is None else p.grad.detach().clone() for p in params])
        return orig(params, max_norm, *a, **k)

    torch.nn.utils.clip_grad_norm_ = hook
    try:
        yield
    finally:
        torch.nn.utils.clip_grad_norm_ = orig


def golden_wm_grads(cfg, tag, B, T, seed, margin=0.25):
    """Per-parameter gradients and post-step weights of the REFERENCE's WorldModel.training_step (WorldModel.py:148-202,
    autocast off = fp32) on the trajectory fixed by `uniforms_used`."""
    sd = W.make_state_dict(cfg, seed=seed)
    cfg = dict(cfg, horizon=T, batch_size=B, sequence_length=T)
    ref = build_reference(cfg, sd)
    obs, act, rew, cont, u = W.sequence_inputs(cfg, B, T, seed=seed + 2)
    with torch.no_grad():
        total, parts, extras = O.world_model_loss(sd, obs, act, rew, cont, u, T, margin_frac=margin, delta=1e-5)
    used = extras[3]
    wm = ref.world_model
    before = {k: p.detach().clone() for k, p in wm.named_parameters()}
    st = Streams(); st.u = [used[t] for t in range(T)]
    snaps = []
    with patched_reference(st), capture_preclip_grads(snaps):
        ref_total = wm.training_step(obs, act, rew, cont).detach()
    assert len(snaps) == 1 and not st.u
    names = [k for k, _ in wm.named_parameters()]
    assert len(names) == len(snaps[0])
    pack = dict(cfg=json.dumps(cfg), seed=seed, B=B, T=T, margin=margin, uniforms_used=used.numpy(), total_loss=ref_total.numpy(),
                idx=extras[2].numpy().astype(np.uint8), lr=np.float64(cfg["world_model_lr"]))
    grad_digest([(k, g) for k, g in zip(names, snaps[0]) if g is not None], "grad", pack)
    grad_digest([(k, p.detach() - before[k]) for k, p in wm.named_parameters()], "dw", pack)
    np.savez_compressed(os.path.join(GOLD, f"{tag}.npz"), **pack)
    err = check(f"{tag}.total_loss", ref_total, total, tol=1e-4)
    gnorm = float(torch.sqrt(sum((g.double() ** 2).sum() for g in snaps[0] if g is not None)))
    return dict(total_loss=err, grad_norm=gnorm, n_params=len(names))


def golden_agent_grads(cfg, tag, B, H, seed):
    """Actor and critic gradients / post-step weights of the REFERENCE's Agent.train_step (Agent.py:96-154) on a rollout made with
    gradients enabled, exactly as Dreamer.train_Agent does (Dreamer.py:264-287): the actor gradient includes the path through
    the imagined states."""
    sd = W.make_state_dict(cfg, seed=seed)
    ref = build_reference(cfg, sd)
    ref.horizon = H
    z0, h0, u, n = W.rollout_inputs(cfg, B, H, seed=seed + 1)
    with torch.no_grad():
        out = O.dream_episodes(sd, z0, h0, u, n, margin_frac=0.25, delta=1e-5)
    used = out[9]
    st = Streams(); st.u = [used[t] for t in range(H)]; st.n = [n[t] for t in range(H)]
    ag = ref.agent
    before = {k: p.detach().clone() for k, p in ag.named_parameters()}
    snaps = []
    with patched_reference(st), capture_preclip_grads(snaps):
        r = ref.dream_episodes(z0, h0)                                   # with autograd (Dreamer.py:274)
        la, lc = ag.train_step(r[0], r[1], r[3], r[4], r[2], r[5], r[6])   # Agent.py:96-154
    assert len(snaps) == 2                                                # critic, then actor (Agent.py:147-148)
    pack = dict(cfg=json.dumps(cfg), seed=seed, B=B, H=H, uniforms_used=used.numpy(), loss_actor=la.detach().numpy(),
                loss_critic=lc.detach().numpy(), S=np.float32(ag.S))
    grad_digest([("critic." + k, g) for (k, _), g in zip(ag.critic.named_parameters(), snaps[0])], "grad", pack)
    grad_digest([("actor." + k, g) for (k, _), g in zip(ag.actor.named_parameters(), snaps[1])], "grad", pack)
    grad_digest([(k, p.detach() - before[k]) for k, p in ag.named_parameters() if not k.startswith("target_critic")], "dw", pack)
    np.savez_compressed(os.path.join(GOLD, f"{tag}.npz"), **pack)
    return dict(loss_actor=float(la), loss_critic=float(lc),
                actor_grad_norm=float(torch.sqrt(sum((g.double() ** 2).sum() for g in snaps[1]))),
                critic_grad_norm=float(torch.sqrt(sum((g.double() ** 2).sum() for g in snaps[0]))))


def golden_acting(cfg, tag, steps, episode_len, seed, margin=0.25):
    """The B = 1 acting loop: the REFERENCE's unmodified Dreamer.rollout_policy (Dreamer.py:177-226) on a fake environment, draws
    redirected to streams.  The oracle walks the same loop first to place every posterior uniform inside its CDF bin; the
    reference then has to reproduce the oracle's classes and actions, and ITS actions / classes / ring contents are stored."""
    from oracle.fake_env import FakeEnv
    sd = W.make_state_dict(cfg, seed=seed)
    cfg = dict(cfg, sequence_length=steps, buffer_size=64)
    ref = build_reference(cfg, sd)
    rng = np.random.Generator(np.random.PCG64(seed + 5))
    u_raw = torch.from_numpy(rng.random((steps + 1, 1, 32), dtype=np.float32))
    normals = torch.from_numpy(rng.standard_normal((steps, 1, 3)).astype(np.float32))
    D = cfg["hidden_state_dims"]

    def norm(frame_hwc):
        return torch.from_numpy(frame_hwc.transpose(2, 0, 1).astype(np.float32) / 255.0 - 0.5).unsqueeze(0)

    def encode0(obs, u):
        lg = O.encoder_logits(sd, torch.zeros(1, D), obs)
        uu = O.interior_uniforms(O.unimix_probs(lg), u, margin, 1e-5)
        z, idx, _ = O.categorical_st(lg, uu)
        return z, idx, uu

    # --- the oracle's walk over the loop (fixes the uniforms) ---
    env = FakeEnv(episode_len, seed=seed + 6)
    obs, _ = env.reset(seed=7)
    h = torch.zeros(1, D)
    used, idxs, acts, hs = [], [], [], []
    z, idx, uu = encode0(norm(obs), u_raw[0]); used.append(uu); idxs.append(idx)
    k = 1
    with torch.no_grad():
        for i in range(steps):
            a, _, _ = O.actor_act(sd, h, z, normals[i])
            acts.append(a)
            obs_, reward, term, trunc, _ = env.step(a.numpy().reshape(-1))
            if term or trunc:
                obs, _ = env.reset(seed=0)
                h = torch.zeros(1, D)
                z, idx, uu = encode0(norm(obs), u_raw[k])
            else:
                z, h, _, idx, uu = O.observe_step(sd, z, h, a, norm(obs_), u_raw[k], margin_frac=margin, delta=1e-5)
            used.append(uu); idxs.append(idx); hs.append(h.clone()); k += 1
    # --- the reference, unmodified, on an identical environment ---
    env_r = FakeEnv(episode_len, seed=seed + 6)
    st = Streams(); st.u = list(used); st.n = [normals[i].reshape(1, 1, 3) for i in range(steps)]
    ref.seed = 7
    with patched_reference(st):
        ref.rollout_policy(env_r, random_policy=False)
    assert not st.u and not st.n
    a_ref = np.stack(env_r.actions)
    err = check(f"{tag}.actions", torch.from_numpy(a_ref), torch.cat(acts).reshape(steps, 3), tol=1e-5)
    check(f"{tag}.hidden_last", ref.agent_hidden.reshape(1, D), hs[-1], tol=1e-5)
    check(f"{tag}.idx_last", ref.agent_latent.argmax(-1).reshape(1, 32), idxs[-1], exact=True)
    n = ref.buffer.size
    np.savez_compressed(os.path.join(GOLD, f"{tag}.npz"), cfg=json.dumps(cfg), seed=seed, steps=steps, episode_len=episode_len,
                        env_seed=seed + 6, uniforms_used=torch.stack(used).numpy(), normals=normals.numpy(), actions=a_ref,
                        idx=torch.stack(idxs).numpy().astype(np.uint8).reshape(steps + 1, 32),
                        hidden_last=ref.agent_hidden.numpy().reshape(D),
                        ring_obs_sum=ref.buffer.observation_buffer[:n].reshape(n, -1).sum(-1).numpy().astype(np.float64) if hasattr(ref.buffer.observation_buffer, "numpy") else np.asarray(ref.buffer.observation_buffer[:n]).reshape(n, -1).sum(-1).astype(np.float64),
                        ring_act=np.asarray(ref.buffer.action_buffer[:n]), ring_rew=np.asarray(ref.buffer.reward_buffer[:n]),
                        ring_cont=np.asarray(ref.buffer.continue_buffer[:n]))
    return dict(actions=err, transitions=int(n))


def golden_replay(tag):
    sys.path.insert(0, REF)
    from Buffer import Buffer  # Buffer.py:5
    cap, L, B = 37, 8, 16
    errs = {}
    packs = {}
    for fill, name in ((20, "partial"), (37 + 11, "wrapped")):
        rb = Buffer(cap, L, 3, (64, 64), device="cpu")
        ro = ReplayOracle(cap, L, 3, (64, 64))
        rng = np.random.Generator(np.random.PCG64(7))
        for i in range(fill):
            o = rng.integers(0, 256, size=(3, 64, 64)).astype(np.uint8)
            a = rng.uniform(-1, 1, 3).astype(np.float32)
            r = float(rng.standard_normal() * 5)
            c = float(i % 9 != 8)
            rb.add_to_buffer(o, a, r, c); ro.add(o, a, r, c)
        np.random.seed(123)
        ob, ab, rbw, cb, Lr = rb.sample_sequences(B)                    # Buffer.py:32-63
        starts = ro.draw_starts(B, rng=_SeededGlobal(123))
        oo, ao, rwo, co, idx = ro.gather(starts)
        check(f"{tag}.{name}.obs", ob, oo, exact=True)
        check(f"{tag}.{name}.act", ab, ao, exact=True)
        check(f"{tag}.{name}.rew", rbw, rwo, exact=True)
        check(f"{tag}.{name}.cont", cb, co, exact=True)
        packs[name + "_starts"] = starts
        packs[name + "_fill"] = fill
        packs[name + "_obs_sum"] = ob.numpy().astype(np.float64).sum(axis=(2, 3, 4))
        packs[name + "_act"] = ab.numpy(); packs[name + "_rew"] = rbw.numpy(); packs[name + "_cont"] = cb.numpy()
        errs[name] = 0.0
    np.savez_compressed(os.path.join(GOLD, f"{tag}.npz"), cap=cap, L=L, B=B, **packs)
    return errs


class _SeededGlobal:
    """np.random-module lookalike seeded like the reference run (legacy MT19937 stream)."""

    def __init__(self, seed):
        self.rs = np.random.RandomState(seed)

    def randint(self, lo, hi, size=None):
        return self.rs.randint(lo, hi, size=size)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    report = {}
    small = W.small_config()
    report["rollout_small"], _, _ = golden_rollout(small, "rollout_small", B=6, H=5, seed=11)
    report["observe_small"] = golden_observe(small, "observe_small", B=3, T=6, seed=21)
    report["agent_small"] = golden_agent(small, "agent_small", B=6, H=5, seed=31)
    report["replay_small"] = golden_replay("replay_small")
    report["wm_grads_small"] = golden_wm_grads(small, "wm_grads_small", B=3, T=6, seed=61)
    report["agent_grads_small"] = golden_agent_grads(small, "agent_grads_small", B=6, H=5, seed=71)
    report["acting_small"] = golden_acting(small, "acting_small", steps=12, episode_len=5, seed=81)
    # full reference sizes: compared here, only a digest is committed
    full = dict(W.REF_CONFIG, horizon=15)
    errs, r, out = golden_rollout(full, "rollout_ref", B=32, H=15, seed=41, save_full=False)
    report["rollout_ref_sizes"] = errs
    np.savez_compressed(os.path.join(GOLD, "rollout_ref_digest.npz"), cfg=json.dumps(full), seed=41, B=32, H=15,
                        margin=0.25, uniforms_used=out[9].numpy(), idx=r[0][:, 1:].argmax(-1).numpy().astype(np.uint8),
                        hidden_last=r[1][:, -1].numpy(), rewards=r[3].numpy(), continues=r[4].numpy(),
                        actions=r[2].numpy(), mu=r[5].numpy(), sigma=r[6].numpy())
    report["observe_ref_sizes"] = golden_observe(dict(W.REF_CONFIG), "observe_ref_digest", B=2, T=4, seed=51)
    with open(os.path.join(GOLD, "PINNING_REPORT.json"), "w") as f:
        json.dump(dict(note="max |reference - oracle| per quantity, measured by oracle/make_golden.py in the build "
                            "container against /root/reference (torch %s, CPU fp32)" % torch.__version__,
                       errors=report), f, indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
