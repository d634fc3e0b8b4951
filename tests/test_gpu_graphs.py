"""CUDA-graph replay of whole training steps (graphs.StepGraph): the captured step must do what the eager step does."""
import numpy as np
import pytest
import torch

from dreamer_b200 import synthetic as W

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _pair(cfg, seed=3):
    sd = W.make_state_dict(cfg, seed=seed)
    return W.build_learners(cfg, sd, DEV), W.build_learners(cfg, sd, DEV)


def _close(a, b, what):
    for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.allclose(x, y, rtol=2e-3, atol=2e-5), (what, k, float((x - y).abs().max()))


def test_world_model_training_step_graph_matches_eager():
    cfg = W.small_config(batch_size=4, sequence_length=8, horizon=8)
    (wm_e, _), (wm_g, _) = _pair(cfg)
    wm_g.enable_cuda_graphs(warmup=2)
    B, T = cfg["batch_size"], cfg["horizon"]
    g = torch.Generator().manual_seed(0)
    for it in range(6):
        obs = torch.randint(0, 256, (B, T, 3, 64, 64), generator=g).float().to(DEV)
        act = (torch.rand(B, T, 3, generator=g) * 2 - 1).to(DEV)
        rew = torch.randn(B, T, 1, generator=g).to(DEV)
        cont = (torch.rand(B, T, 1, generator=g) > 0.1).float().to(DEV)
        u = torch.rand(T, B, 32, generator=g).to(DEV)
        le = wm_e.training_step(obs, act, rew, cont, uniforms=u)
        lg = wm_g.training_step(obs, act, rew, cont, uniforms=u)
        assert torch.allclose(le, lg, rtol=1e-4, atol=1e-5), it
        assert wm_g._graphs.captured(obs, act, rew, cont, u) == (it >= 2)
    assert float(wm_g.optimiser.opt_state[0]) == 6 == float(wm_e.optimiser.opt_state[0])
    _close(wm_e, wm_g, "world model")
    # the replayed step re-packs the weights it trains: a fresh forward on the updated parameters agrees too
    le, _ = wm_e.loss_forward(obs, act, rew, cont, u)
    wm_g.__dict__["_graphs"] = None
    lg, _ = wm_g.loss_forward(obs, act, rew, cont, u)
    assert torch.allclose(le, lg, rtol=1e-4, atol=1e-5)


def test_agent_train_step_graph_matches_eager_and_skips_nonfinite_on_device():
    cfg = W.small_config(horizon=5)
    (_, ag_e), (_, ag_g) = _pair(cfg, seed=4)
    ag_g.enable_cuda_graphs(warmup=2)
    B, H, D = 48, cfg["horizon"], cfg["hidden_state_dims"]
    g = torch.Generator().manual_seed(1)

    def batch(poison=False):
        z = torch.nn.functional.one_hot(torch.randint(0, 32, (B, H + 1, 32), generator=g), 32).float().to(DEV)
        h = torch.tanh(torch.randn(B, H + 1, D, generator=g)).to(DEV)
        r = torch.randn(B, H, 1, generator=g).to(DEV)
        c = (torch.rand(B, H, 1, generator=g) > 0.05).float().to(DEV)
        mu = (torch.randn(B, H, 3, generator=g) * 0.3).to(DEV)
        sg = (torch.rand(B, H, 3, generator=g) * 0.5 + 0.1).to(DEV)
        a = torch.tanh(mu + sg * torch.randn(B, H, 3, generator=g).to(DEV))
        if poison:
            r[0, 0, 0] = float("nan")
        return z, h, r, c, a, mu, sg

    for it in range(6):
        b = batch()
        la_e, lc_e = ag_e.train_step(*b)
        la_g, lc_g = ag_g.train_step(*b)
        assert torch.allclose(la_e, la_g, rtol=1e-4, atol=1e-5) and torch.allclose(lc_e, lc_g, rtol=1e-4, atol=1e-5), it
    _close(ag_e, ag_g, "agent")
    assert torch.allclose(torch.as_tensor(ag_e.S).cpu(), torch.as_tensor(ag_g.S).cpu(), rtol=1e-5)
    # a poisoned batch inside the replayed graph: no host check exists there, the device-side skip must hold everything still
    before = {k: v.detach().clone() for k, v in ag_g.state_dict().items()}
    s_before = ag_g.S.clone()
    la, lc = ag_g.train_step(*batch(poison=True))
    assert not torch.isfinite(lc)
    assert bool(ag_g.critic_optimiser.last_step_skipped)
    for k, v in ag_g.state_dict().items():
        if k.startswith("critic") or k.startswith("target_critic"):
            assert torch.equal(v, before[k]), k
    assert torch.equal(ag_g.S, s_before)
    # and the next clean batch trains again
    ag_g.train_step(*batch())
    assert not bool(ag_g.critic_optimiser.last_step_skipped)


def test_graphed_rollout_follows_weight_updates():
    """Rollout.run_graphed: the captured rollout reads the packed weights in place, so after a parameter update (re-packed by
    the engine outside the graph) the replay must equal a fresh eager rollout bit for bit."""
    from dreamer_b200 import rollout as R
    cfg = W.small_config(horizon=6)
    wm, ag = W.build_learners(cfg, W.make_state_dict(cfg, seed=5), DEV)
    B, H = 40, cfg["horizon"]
    for it in range(5):
        z0, h0, u, n = (t.to(DEV) for t in W.rollout_inputs(cfg, B, H, seed=60 + it))
        got = [t.clone() for t in R.dream_episodes_modules(wm, ag, z0, h0, H, u, n, graphed=True)]
        ref = R.dream_episodes_modules(wm, ag, z0, h0, H, u, n)
        for a, b in zip(got, ref):
            assert torch.equal(a, b), it
        with torch.no_grad():                       # "training": every weight moves, versions bump -> the engine re-packs
            for p in list(wm.parameters()) + list(ag.actor.parameters()):
                p.mul_(1.0 + 0.01 * (it + 1))
    ro = wm._engine.rollout(B, H)
    assert ro._graphs[False].captured(z0.float(), h0.float(), u, n)
