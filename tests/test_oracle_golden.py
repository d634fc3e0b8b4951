"""The oracle replayed against the committed fixtures (which hold the REFERENCE's outputs,
written by oracle/make_golden.py in the build container).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import rssm as O
from oracle import weights as W
from oracle.replay import ReplayOracle


def _load(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    return g, json.loads(str(g["cfg"])) if "cfg" in g.files else None


def _close(a, b, tol):
    a, b = torch.as_tensor(a).float(), torch.as_tensor(b).float()
    assert a.shape == b.shape
    assert (a - b).abs().max().item() <= tol * max(1.0, a.abs().max().item())


@pytest.mark.parametrize("name", ["rollout_small.npz", "rollout_ref_digest.npz"])
def test_rollout_matches_reference_fixture(golden_dir, name):
    g, cfg = _load(golden_dir, name)
    B, H, seed = int(g["B"]), int(g["H"]), int(g["seed"])
    sd = W.make_state_dict(cfg, seed=seed)
    z0, h0, _, n = W.rollout_inputs(cfg, B, H, seed=seed + 1)
    with torch.no_grad():
        out = O.dream_episodes(sd, z0, h0, torch.from_numpy(g["uniforms_used"]), n)
    assert np.array_equal(out[7].numpy().astype(np.uint8), g["idx"])          # sampled indices: exact
    for key, i in (("actions", 2), ("rewards", 3), ("continues", 4), ("mu", 5), ("sigma", 6)):
        _close(g[key], out[i], 2e-5)
    if "hidden" in g.files:
        _close(g["hidden"], out[1], 2e-5)
        _close(g["latent_last"], out[0][:, -1], 1e-6)
    else:
        _close(g["hidden_last"], out[1][:, -1], 2e-5)


@pytest.mark.parametrize("name", ["observe_small.npz", "observe_ref_digest.npz"])
def test_observe_and_wm_loss_match_reference_fixture(golden_dir, name):
    g, cfg = _load(golden_dir, name)
    B, T, seed = int(g["B"]), int(g["T"]), int(g["seed"])
    sd = W.make_state_dict(cfg, seed=seed)
    obs, act, rew, cont, _ = W.sequence_inputs(cfg, B, T, seed=seed + 2)
    with torch.no_grad():
        total, parts, extras = O.world_model_loss(sd, obs, act, rew, cont, torch.from_numpy(g["uniforms_used"]), T)
        (prior, post, obs_ll, rew_ll, cont_bce), _ = O.unroll_model(
            sd, obs / 255.0 - 0.5, act, rew, cont, torch.from_numpy(g["uniforms_used"]))
    assert np.array_equal(extras[2].numpy().astype(np.uint8), g["idx"])
    _close(g["prior_logits"], prior, 1e-4)
    _close(g["post_logits"], post, 1e-4)
    _close(g["obs_ll"], obs_ll, 1e-4)
    _close(g["rew_ll"], rew_ll, 1e-4)
    _close(g["cont_bce"], cont_bce, 1e-4)
    _close(g["total_loss"], total, 1e-4)
    wlen = T // 2
    with torch.no_grad():
        z, h, _ = O.warm_start(sd, obs, act, torch.from_numpy(g["warm_uniforms_used"]), wlen)
    assert np.array_equal(z.argmax(-1).numpy().astype(np.uint8), g["warm_idx"])
    _close(g["warm_hidden"], h, 1e-4)


def test_agent_losses_match_reference_fixture(golden_dir):
    g, cfg = _load(golden_dir, "agent_small.npz")
    B, H, seed = int(g["B"]), int(g["H"]), int(g["seed"])
    sd = W.make_state_dict(cfg, seed=seed)
    z0, h0, _, n = W.rollout_inputs(cfg, B, H, seed=seed + 1)
    with torch.no_grad():
        out = O.dream_episodes(sd, z0, h0, torch.from_numpy(g["uniforms_used"]), n)
        res = O.agent_losses(sd, out[0], out[1], out[3], out[4], out[2], out[5], out[6], S=1.0,
                             gamma=cfg["gamma"], lam=cfg["lambda_"], nu=cfg["nu"])
    _close(g["returns"], res["returns"], 1e-4)
    _close(g["values"], res["values"], 1e-4)
    _close(g["loss_actor"], res["loss_actor"], 1e-4)
    _close(g["loss_critic"], res["loss_critic"], 1e-4)
    _close(g["S"], torch.as_tensor(res["S_new"]), 1e-5)
    _close(g["twohot"], O.to_twohot(torch.from_numpy(g["twohot_vals"]), sd["agent.critic.buckets_crit"]), 1e-6)


def test_twohot_edges():
    """v <= -20 -> bucket 0 weight 1; v >= 20 -> buckets 253/254 weights 0/1."""
    b = torch.linspace(-20.0, 20.0, 255)
    th = O.to_twohot(torch.tensor([[-50.0], [50.0]]), b)
    assert th[0, 0] == 1.0 and th[0].sum() == 1.0
    assert th[1, 254] == pytest.approx(1.0, abs=1e-6) and th[1, 253] == pytest.approx(0.0, abs=1e-6)
    idx, w = O.twohot_index_weight(torch.tensor([[0.3]]), b)
    assert th.shape == (2, 255) and 0 <= float(w) <= 1 and b[idx] <= 0.3 < b[idx + 1]


class _Legacy:
    def __init__(self, seed):
        self.rs = np.random.RandomState(seed)

    def randint(self, lo, hi, size=None):
        return self.rs.randint(lo, hi, size=size)


def test_replay_matches_reference_fixture(golden_dir):
    g, _ = _load(golden_dir, "replay_small.npz")
    cap, L, B = int(g["cap"]), int(g["L"]), int(g["B"])
    for name in ("partial", "wrapped"):
        ro = ReplayOracle(cap, L, 3, (64, 64))
        rng = np.random.Generator(np.random.PCG64(7))
        for i in range(int(g[name + "_fill"])):
            o = rng.integers(0, 256, size=(3, 64, 64)).astype(np.uint8)
            a = rng.uniform(-1, 1, 3).astype(np.float32)
            r = float(rng.standard_normal() * 5)
            ro.add(o, a, r, float(i % 9 != 8))
        starts = ro.draw_starts(B, rng=_Legacy(123))
        assert np.array_equal(starts, g[name + "_starts"])                    # indices: exact
        o, a, r, c, idx = ro.gather(starts)
        assert np.array_equal(o.astype(np.float64).sum(axis=(2, 3, 4)), g[name + "_obs_sum"])
        assert np.array_equal(a, g[name + "_act"]) and np.array_equal(r, g[name + "_rew"]) and np.array_equal(c, g[name + "_cont"])
        assert idx.max() < cap


def test_replay_too_short_raises():
    ro = ReplayOracle(8, 5, 3, (64, 64))
    with pytest.raises(ValueError):
        ro.draw_starts(2)


def test_sharded_rollout_concatenates():
    """rank shards with sliced uniforms/normals concatenate to the full-batch result."""
    cfg = W.small_config()
    sd = W.make_state_dict(cfg, seed=5)
    z0, h0, u, n = W.rollout_inputs(cfg, 8, 4, seed=6)
    with torch.no_grad():
        full = O.dream_episodes(sd, z0, h0, u, n)
        parts = [O.dream_episodes(sd, z0[s], h0[s], u[:, s], n[:, s]) for s in (slice(0, 4), slice(4, 8))]
    assert torch.equal(torch.cat([p[7] for p in parts]), full[7])
    assert torch.allclose(torch.cat([p[1] for p in parts]), full[1], atol=1e-6)


def test_optim_oracle_matches_torch():
    """oracle/optim.py (clip_grad_norm_ + AdamW + soft target update) against the installed torch on CPU."""
    import torch
    from oracle import optim as O
    torch.manual_seed(0)
    ps = [torch.randn(33, 7), torch.randn(19)]
    ref = [torch.nn.Parameter(p.clone()) for p in ps]
    opt = torch.optim.AdamW(ref, lr=4e-3, betas=(0.9, 0.95), eps=1e-7, weight_decay=1e-2)
    p = [x.numpy().copy() for x in ps]
    m, v, step = [np.zeros_like(x) for x in p], [np.zeros_like(x) for x in p], 0
    for it in range(6):
        gs = [torch.randn_like(x) * (50 if it % 2 else 1) for x in ps]     # odd steps: the clip at 100 is active
        for r, g in zip(ref, gs):
            r.grad = g.clone()
        tn = torch.nn.utils.clip_grad_norm_(ref, 100.0)
        opt.step()
        step, total = O.adamw_step(p, [g.numpy() for g in gs], m, v, step, 4e-3, (0.9, 0.95), 1e-7, 1e-2, 100.0)
        assert abs(float(tn) - float(total)) <= 1e-6 * float(tn)
        for a, b in zip(p, ref):
            assert np.allclose(a, b.detach().numpy(), rtol=1e-6, atol=1e-7)
    tgt = [np.ones_like(x) for x in p]
    O.soft_update(tgt, p, 0.02)
    assert np.allclose(tgt[0], 0.98 + 0.02 * p[0])
