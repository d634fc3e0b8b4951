"""One full synthetic Dreamer training iteration on the mirrored modules (BASELINE configs[4] at test size): replay fill ->
world-model updates -> warm start -> imagination -> actor-critic update.  Checks the callers of the hot path
(Dreamer.py:143-175, 228-287) against the oracle and that training actually moves the parameters."""
import numpy as np
import pytest
import torch

from dreamer_b200 import synthetic as W
from oracle import rssm as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _fill(buf, cfg, n, seed=0):
    rng = np.random.Generator(np.random.PCG64(seed))
    obs = rng.integers(0, 256, size=(n, 3, 64, 64)).astype(np.uint8)
    act = rng.uniform(-1, 1, (n, cfg["action_dims"])).astype(np.float32)
    rew = rng.standard_normal(n).astype(np.float32)
    con = (rng.random(n) > 0.05).astype(np.float32)
    buf.add_batch(obs, act, rew, con)
    return obs, act, rew, con


def test_warm_start_matches_oracle_and_iteration_trains():
    from dreamer_b200.hotpath import HotPath
    cfg = W.small_config(batch_size=5, sequence_length=10, horizon=4, buffer_size=96)
    torch.manual_seed(0)
    hp = HotPath(cfg, DEV)
    sd = {"world_model." + k: v.detach().cpu().clone() for k, v in hp.world_model.state_dict().items()}
    sd.update({"agent." + k: v.detach().cpu().clone() for k, v in hp.agent.state_dict().items()})
    _fill(hp.buffer, cfg, 80)
    assert hp.buffer.size == 80 and hp.buffer.next_idx == 80
    # warm start (fused scan, mode 1) vs the oracle's restatement of Dreamer.warm_start_generator
    np.random.seed(1)
    obs, act, _, _, L = hp.buffer.sample_sequences(cfg["batch_size"])
    Wn = L // 2
    u = torch.rand(Wn, cfg["batch_size"], 32)
    zo, ho, used = O.warm_start(sd, obs.cpu(), act.cpu(), u, Wn, margin_frac=0.25, delta=1e-5)
    z0, h0 = hp.warm_start_generator(obs, act, L, uniforms=used.to(DEV))
    assert z0.shape == (cfg["batch_size"], 1, 32, 32) and h0.shape == (cfg["batch_size"], 1, cfg["hidden_state_dims"])
    assert torch.equal(z0.cpu().argmax(-1), zo.argmax(-1))
    assert torch.allclose(h0.cpu(), ho, atol=2e-2, rtol=1e-2)
    # dream_episodes returns the reference's 7-tuple shapes
    out = hp.dream_episodes(z0, h0)
    H = cfg["horizon"]
    assert [tuple(t.shape[1:]) for t in out] == [(H + 1, 32, 32), (H + 1, cfg["hidden_state_dims"]), (H, 3), (H, 1), (H, 1), (H, 3), (H, 3)]
    # one full iteration: WM epochs then AC epochs
    before = {k: v.detach().clone() for k, v in list(hp.world_model.state_dict().items()) + list(hp.agent.state_dict().items())}
    wm_losses = hp.train_world_model()
    la, lc = hp.train_Agent()
    assert len(wm_losses) == cfg["WM_epochs"] and all(torch.isfinite(x) for x in wm_losses)
    assert torch.isfinite(la) and torch.isfinite(lc) and lc.item() > 0
    after = dict(list(hp.world_model.state_dict().items()) + list(hp.agent.state_dict().items()))
    moved = [k for k in before if before[k].dtype.is_floating_point and "buckets" not in k and not torch.equal(before[k], after[k])]
    assert any(k.startswith("sequence_model") for k in moved) and any(k.startswith("actor") for k in moved) and any(k.startswith("critic") for k in moved)
    assert any(k.startswith("target_critic") for k in moved)                 # EMA update, Agent.py:90-94
    # a second WM step on the same batch lowers the loss (the gradient path is live end to end)
    np.random.seed(2)
    o2, a2, r2, c2, _ = hp.buffer.sample_sequences(cfg["batch_size"])
    uu = torch.rand(cfg["horizon"], cfg["batch_size"], 32, device=DEV)
    l0 = hp.world_model.training_step(o2, a2, r2, c2, uniforms=uu).item()
    for _ in range(5):
        hp.world_model.training_step(o2, a2, r2, c2, uniforms=uu)
    l1 = hp.world_model.loss_forward(o2, a2, r2, c2, uniforms=uu)[0].item()
    assert l1 < l0, (l0, l1)


def test_training_state_resume_continues_the_run(tmp_path):
    """save_training_state / load_training_state (optimiser moments + step counts, S, RNG, ring): a run resumed from the file
    draws the same replay windows and latents, so its first loss is bit-identical to the original run's next loss, and it stays
    on the original run's trajectory (to the rounding of cuDNN's non-deterministic conv-gradient reductions)."""
    from dreamer_b200.hotpath import HotPath
    cfg = W.small_config(batch_size=4, sequence_length=8, horizon=4, buffer_size=64)
    torch.manual_seed(0)
    a = HotPath(cfg, DEV)
    _fill(a.buffer, cfg, 48)
    np.random.seed(3)
    a.train_world_model(); a.train_Agent()
    path = str(tmp_path / "state.pt")
    a.save_training_state(path, include_buffer=True)
    m_a, v_a = a.world_model.optimiser.exp_avg.clone(), a.agent.critic_optimiser.exp_avg_sq.clone()
    t_a = float(a.world_model.optimiser.opt_state[0])
    la, aa = a.train_world_model(), a.train_Agent()       # the original run goes on (the RNG streams are process-global) ...
    torch.manual_seed(1)
    b = HotPath(cfg, DEV)
    b.load_training_state(path)                           # ... then a fresh process state is rewound to the checkpoint
    assert b.buffer.size == a.buffer.size and torch.equal(b.buffer.observation_buffer[:48], a.buffer.observation_buffer[:48])
    assert float(b.world_model.optimiser.opt_state[0]) == t_a > 0
    assert torch.equal(m_a, b.world_model.optimiser.exp_avg) and torch.equal(v_a, b.agent.critic_optimiser.exp_avg_sq)
    lb, ab = b.train_world_model(), b.train_Agent()
    assert torch.equal(la[0], lb[0])
    assert all(torch.allclose(x, y, rtol=1e-3) for x, y in zip(la, lb))
    assert torch.allclose(aa[0], ab[0], rtol=1e-2, atol=1e-3) and torch.allclose(aa[1], ab[1], rtol=1e-3)
    for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.allclose(x, y, rtol=1e-3, atol=1e-4), k
    assert abs(float(a.agent.S) - float(b.agent.S)) < 1e-5
    assert float(a.world_model.optimiser.opt_state[0]) == float(b.world_model.optimiser.opt_state[0])
