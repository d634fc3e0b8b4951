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
    # (with the reference's fp32 round trip of the stored frame, Dreamer.py:186,209)
    assert np.array_equal(buf.observation_buffer[:n].cpu().numpy(), (((frames[:n].astype(np.float32) / 255.0) - 0.5 + 0.5) * 255.0).astype(np.uint8))
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


class _FakeEnv:
    """A gymnasium-shaped environment with HWC uint8 frames, fixed episode length and a counter-derived reward."""

    class _Space:
        def __init__(self, rng):
            self.rng = rng

        def sample(self):
            return self.rng.uniform(-1, 1, 3).astype(np.float32)

    def __init__(self, episode_len, seed=0):
        self.rng = np.random.default_rng(seed)
        self.action_space = self._Space(self.rng)
        self.episode_len, self.t, self.frames, self.actions = episode_len, 0, [], []

    def _frame(self):
        f = self.rng.integers(0, 256, size=(64, 64, 3)).astype(np.uint8)
        self.frames.append(f)
        return f

    def reset(self, seed=None):
        self.t = 0
        return self._frame(), {}

    def step(self, action):
        self.actions.append(np.asarray(action, dtype=np.float32).copy())
        self.t += 1
        done = self.t >= self.episode_len
        return self._frame(), 0.5 * self.t, done, False, {}


class _DreamerShell:
    """The attributes of the reference's Dreamer that rollout_policy / evaluate_agent / warm_start_generator touch
    (Dreamer.py:119-139); the methods themselves come from dropin.patch_dreamer."""

    def __init__(self, cfg):
        from dreamer_b200.modules import Buffer
        self.world_model, self.agent = W.build_learners(cfg, W.make_state_dict(cfg, seed=41), DEV)
        self.buffer = Buffer(256, cfg["sequence_length"], cfg["action_dims"], tuple(cfg["observation_dims"]), device=DEV)
        self.sequence_length, self.horizon, self.seed, self.agent_obs = cfg["sequence_length"], cfg["horizon"], 7, None


def test_patched_rollout_policy_and_evaluate_agent_follow_the_reference_loop():
    """dropin.patch_dreamer's rollout_policy / evaluate_agent (Dreamer.py:177-226, 295-322 on acting.ActingPath): every
    transition (obs_t, a_t, r_t, c_t) the environment saw lands in the ring in order, episodes restart on done, the random
    policy's actions are the environment's samples, and evaluation returns the mean episode return without touching the ring."""
    from dreamer_b200 import dropin
    cfg = W.small_config(sequence_length=6, horizon=4)
    Shell = dropin.patch_dreamer(type("Shell", (_DreamerShell,), {}))
    d = Shell(cfg)
    env = _FakeEnv(episode_len=4)
    d.rollout_policy(env, random_policy=True)
    d.rollout_policy(env, random_policy=False)
    n = 2 * cfg["sequence_length"]
    assert d.buffer.size == n and len(env.actions) == n
    # frames: env.frames holds reset frames and step frames in order; the transition stored at step i is the frame the action was chosen on
    stored = d.buffer.observation_buffer[:n].cpu().numpy()
    seen, k = [], 0
    cur = env.frames[0]
    idx = 1
    for i in range(n):
        seen.append(cur.transpose(2, 0, 1))
        nxt = env.frames[idx]; idx += 1
        if (i + 1) % 4 == 0:            # episode ended: the next stored frame is the reset frame that follows
            cur = env.frames[idx]; idx += 1
        else:
            cur = nxt
    roundtrip = lambda x: (((x.astype(np.float32) / 255.0) - 0.5 + 0.5) * 255.0).astype(np.uint8)      # what Dreamer.py:186,209 store
    assert np.array_equal(stored, roundtrip(np.stack(seen)))
    assert np.allclose(d.buffer.action_buffer[:n].cpu().numpy(), np.stack(env.actions), atol=1e-6)
    cont = d.buffer.continue_buffer[:n].cpu().numpy().reshape(-1)
    assert np.array_equal(cont, np.array([0.0 if (i + 1) % 4 == 0 else 1.0 for i in range(n)], dtype=np.float32))
    assert np.all(np.abs(np.stack(env.actions[cfg["sequence_length"]:])) <= 1.0) and d.seed == 7 + n // 4
    before = d.buffer.size
    mean_ret = d.evaluate_agent(_FakeEnv(episode_len=3, seed=1), eval_episodes=2)
    assert d.buffer.size == before and abs(float(mean_ret) - 0.5 * (1 + 2 + 3)) < 1e-6
    # warm start + imagination through the patched methods keep the reference's shapes
    np.random.seed(0)
    obs, act, _, _, L = d.buffer.sample_sequences(3)
    z0, h0 = d.warm_start_generator(obs, act, L)
    out = d.dream_episodes(z0, h0)
    assert z0.shape == (3, 1, 32, 32) and out[0].shape == (3, cfg["horizon"] + 1, 32, 32) and out[2].shape == (3, cfg["horizon"], 3)


def test_acting_path_matches_the_reference_rollout_policy_fixture(golden_dir):
    """The acting loop against the REFERENCE's unmodified Dreamer.rollout_policy (Dreamer.py:177-226) on the same fake environment
    and the same draws (tests/golden/acting_small.npz, oracle/make_golden.py: golden_acting): every posterior class of every
    step bit-exact, actions within the bf16 bound, the ring contents (frames, actions, symlog rewards, continue flags) as the
    reference's ring holds them, across an episode boundary."""
    import json
    import os
    from oracle.fake_env import FakeEnv
    from dreamer_b200.acting import ActingPath
    from dreamer_b200.modules import Buffer
    g = np.load(os.path.join(golden_dir, "acting_small.npz"))
    cfg = json.loads(str(g["cfg"]))
    steps, ep_len, seed = int(g["steps"]), int(g["episode_len"]), int(g["seed"])
    wm, ag = W.build_learners(cfg, W.make_state_dict(cfg, seed=seed), DEV)
    ring = Buffer(64, cfg["sequence_length"], cfg["action_dims"], tuple(cfg["observation_dims"]), device=DEV)
    ap = ActingPath(wm, ag, ring, deterministic=False, warmup=2)
    env = FakeEnv(ep_len, seed=int(g["env_seed"]))
    u = torch.from_numpy(g["uniforms_used"]).to(DEV)          # (steps + 1, 1, 32)
    nrm = torch.from_numpy(g["normals"]).to(DEV)              # (steps, 1, 3)
    chw = lambda f: np.ascontiguousarray(f.transpose(2, 0, 1))
    obs, _ = env.reset(seed=7)
    ap.reset(chw(obs), u[0])
    idx = [ap.latent.argmax(-1).view(-1).cpu().numpy()]
    k = 1
    for i in range(steps):
        a = ap.act(nrm[i])
        obs_, reward, term, trunc, _ = env.step(a)
        done = bool(term or trunc)
        ap.record(reward, 1 - done)
        if done:
            obs, _ = env.reset(seed=0)
            ap.reset(chw(obs), u[k])
        else:
            ap.observe(chw(obs_), u[k])
        idx.append(ap.latent.argmax(-1).view(-1).cpu().numpy())
        k += 1
    assert np.array_equal(np.stack(idx).astype(np.uint8), g["idx"])                        # every class of every step
    acts = np.stack(env.actions)
    assert np.abs(acts - g["actions"]).max() <= 1e-2 * np.abs(g["actions"]).max()
    hl = ap.hidden.view(-1).cpu().numpy()
    assert np.abs(hl - g["hidden_last"]).max() <= 1e-2 * np.abs(g["hidden_last"]).max()
    n = steps
    assert ring.size == n
    assert np.array_equal(ring.observation_buffer[:n].reshape(n, -1).sum(-1, dtype=torch.float64).cpu().numpy(), g["ring_obs_sum"])
    assert np.abs(ring.action_buffer[:n].cpu().numpy() - g["ring_act"]).max() <= 1e-2
    assert np.allclose(ring.reward_buffer[:n].cpu().numpy(), g["ring_rew"], atol=1e-6)
    assert np.array_equal(ring.continue_buffer[:n].cpu().numpy(), g["ring_cont"])


def test_patched_rollout_policy_acts_on_updated_weights_mid_episode():
    """An episode continues across training phases (Dreamer.py:340-343): the second rollout_policy call must act with the weights
    as they are NOW, although no reset happened in between (the captured graphs read packed weight caches)."""
    from dreamer_b200 import dropin
    cfg = W.small_config(sequence_length=5, horizon=4)
    Shell = dropin.patch_dreamer(type("Shell", (_DreamerShell,), {}))
    d = Shell(cfg)
    env = _FakeEnv(episode_len=1000)                     # no episode boundary in this test
    for _ in range(3):                                   # warm-up calls: the acting graphs get captured
        d.rollout_policy(env, random_policy=False)
    n0 = len(env.actions)
    with torch.no_grad():                                # "training": push the policy mean far positive
        d.agent.actor.mu_head.bias.add_(6.0)
    torch.autograd.graph.increment_version(list(d.agent.actor.parameters()))
    d.rollout_policy(env, random_policy=False)
    new = np.stack(env.actions[n0:])
    assert new.shape[0] == cfg["sequence_length"] and (new > 0.9).all(), new


def test_standalone_actor_with_non_square_latents():
    """Actor mirrors the reference's swapped constructor names (Agent.py:175); the packed engine must still get (rows, classes)."""
    from dreamer_b200 import modules as M
    torch.manual_seed(0)
    R, C, Dh = 16, 32, 64
    actor = M.Actor(3, R, C, Dh, 40, 40, device=DEV)
    with torch.no_grad():
        actor.mu_head.weight.normal_(0, 0.1)
    h = torch.tanh(torch.randn(5, 1, Dh, device=DEV))
    z = torch.nn.functional.one_hot(torch.randint(0, C, (5, 1, R), device=DEV), C).float()
    mu, sigma = actor(h, z)
    base = actor.base_net(torch.cat([h, z.flatten(2)], -1))
    mu_ref = actor.mu_head(base)
    assert (mu - mu_ref).abs().max() <= 1e-2 * max(mu_ref.abs().max().item(), 1e-3) + 1e-3
