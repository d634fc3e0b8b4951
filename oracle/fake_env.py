"""A gymnasium-shaped environment with HWC uint8 frames, fixed episode length and a counter-derived reward.

TEST INFRASTRUCTURE (the real CarRacing environment needs gymnasium / Box2D, which this image does not have): drives the reference's
unmodified Dreamer.rollout_policy in oracle/make_golden.py and the acting path's parity test with the same frame stream.
"""
import numpy as np


class FakeEnv:
    class _Space:
        def __init__(self, rng):
            self.rng = rng

        def sample(self):
            return self.rng.uniform(-1, 1, 3).astype(np.float32)

    def __init__(self, episode_len, seed=0, hw=(64, 64)):
        self.rng = np.random.Generator(np.random.PCG64(seed))
        self.action_space = self._Space(self.rng)
        self.episode_len, self.t, self.hw = episode_len, 0, hw
        self.frames, self.actions = [], []

    def _frame(self):
        f = self.rng.integers(0, 256, size=(self.hw[0], self.hw[1], 3)).astype(np.uint8)
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
