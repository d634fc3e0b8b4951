"""The B = 1 acting path (dreamer_b200/acting.py) against the step-by-step module calls of Dreamer.rollout_policy."""
import numpy as np
import pytest
import torch

from dreamer_b200 import synthetic as W

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_acting_path_matches_module_calls_and_fills_the_ring():
    from dreamer_b200.acting import ActingPath
    from dreamer_b200.modules import Buffer
    cfg = W.small_config()
    wm, ag = W.build_learners(cfg, W.make_state_dict(cfg, seed=21), DEV)
    buf = Buffer(64, 8, cfg["action_dims"], tuple(cfg["observation_dims"]), device=DEV)
    acting = ActingPath(wm, ag, buf, warmup=2)
    rng = np.random.default_rng(0)
    g = torch.Generator(device="cuda").manual_seed(5)
    n = 9
    frames = rng.integers(0, 256, size=(n + 1, 3, 64, 64)).astype(np.uint8)
    rewards = rng.standard_normal(n).astype(np.float32)
    conts = (rng.random(n) > 0.2).astype(np.float32)
    us = torch.rand(n + 1, 1, 32, device=DEV, generator=g)
    ns = torch.randn(n + 1, 1, cfg["action_dims"], device=DEV, generator=g)
    # reference sequence of module calls (Dreamer.py:184-226)
    obs = lambda i: (torch.from_numpy(frames[i]).to(DEV).float() / 255.0 - 0.5).view(1, 1, 3, 64, 64)
    h = torch.zeros(1, 1, cfg["hidden_state_dims"], device=DEV)
    z, _ = wm.encoder.encode(h, obs(0), us[0])
    a, _, _ = ag.actor.act(h, z, normals=ns[0])
    ref_actions = [a.view(-1).cpu().numpy()]
    for i in range(n):
        z, h, _ = wm.observe_step(z, h, a, obs(i + 1), us[i + 1])
        a, _, _ = ag.actor.act(h, z, normals=ns[i + 1])
        ref_actions.append(a.view(-1).cpu().numpy())
    # the acting path: eager warm-up calls, then replayed graphs (n > warmup)
    acting.reset(frames[0], us[0])
    got = [acting.act(ns[0])]
    for i in range(n):
        got.append(acting.step(frames[i + 1], rewards[i], conts[i], us[i + 1], ns[i + 1]))
    assert acting._observe_act.captured(us[0], ns[0])
    for r, x in zip(ref_actions, got):
        assert np.allclose(r, x, atol=1e-6), (r, x)
    assert torch.allclose(acting.hidden, h) and torch.equal(acting.latent.argmax(-1), z.argmax(-1))
    # transitions landed in the HBM ring exactly as buffer.add_to_buffer(obs_t, a_t, r_t, c_t) would have put them
    assert buf.size == n and buf.next_idx == n
    assert np.array_equal(buf.observation_buffer[:n].cpu().numpy(), frames[:n])
    assert np.allclose(buf.action_buffer[:n].cpu().numpy(), np.stack(ref_actions[:n]), atol=1e-6)
    assert np.allclose(buf.continue_buffer[:n].cpu().numpy().reshape(-1), conts)
    sym = np.sign(rewards) * np.log1p(np.abs(rewards))
    assert np.allclose(buf.reward_buffer[:n].cpu().numpy().reshape(-1), sym, atol=1e-6)


def test_acting_path_sees_weight_updates_and_draws_on_device():
    """The captured graphs read packed weight buffers that sync_weights() refreshes in place; without explicit
    uniforms / normals the draws happen inside the graph and differ from step to step."""
    from dreamer_b200.acting import ActingPath
    cfg = W.small_config()
    wm, ag = W.build_learners(cfg, W.make_state_dict(cfg, seed=22), DEV)
    acting = ActingPath(wm, ag, None, deterministic=True, warmup=1)
    frame = np.random.default_rng(1).integers(0, 256, size=(3, 64, 64)).astype(np.uint8)
    u = torch.full((1, 32), 0.5, device=DEV)
    acts = []
    for _ in range(3):
        acting.reset(frame, u)
        acts.append(acting.act())
    assert np.array_equal(acts[0], acts[1]) and np.array_equal(acts[1], acts[2])      # deterministic policy, same state
    with torch.no_grad():
        ag.actor.mu_head.bias.add_(0.3)
    torch.autograd.graph.increment_version(list(ag.actor.parameters()))
    acting.reset(frame, u)
    moved = acting.act()
    assert not np.allclose(moved, acts[0])
    sto = ActingPath(wm, ag, None, deterministic=False, warmup=1)
    sto.reset(frame, u)
    draws = np.stack([sto.act() for _ in range(6)])
    assert len({tuple(np.round(d, 6)) for d in draws}) > 3
