"""numpy restatement of the reference replay ring (Buffer.py) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import numpy as np


def symlog_np(x):
    """DreamerUtils.py:32-33."""
    return np.sign(x) * np.log(1.0 + np.abs(x))


class ReplayOracle:
    """Buffer.py:5-63.  Start indices are drawn from a caller-supplied ``np.random.Generator``-like
    object exposing ``randint(lo, hi, size=None)`` (the reference uses the global ``np.random``
    module, Buffer.py:37,44; pass ``np.random`` itself to mimic that)."""

    def __init__(self, capacity, sequence_length, action_size, observation_dims):
        self.obs = np.zeros((capacity, 3, *observation_dims), dtype=np.uint8)   # Buffer.py:7
        self.act = np.zeros((capacity, action_size), dtype=np.float32)          # Buffer.py:8
        self.rew = np.zeros((capacity, 1), dtype=np.float32)                    # Buffer.py:9
        self.con = np.zeros((capacity, 1), dtype=np.float32)                    # Buffer.py:10
        self.capacity = capacity
        self.sequence_length = sequence_length
        self.next_idx = 0
        self.size = 0

    def add(self, observation, action, reward, continue_):
        """Buffer.py:19-30 (reward stored symlog'd)."""
        i = self.next_idx
        self.obs[i] = np.array(observation, dtype=np.uint8)
        self.act[i] = np.array(action, dtype=np.float32)
        self.con[i] = np.array(continue_, dtype=np.float32)
        self.rew[i] = symlog_np(np.array(reward, dtype=np.float32))
        self.next_idx = (i + 1) % self.capacity
        self.size = min(self.size + 1, self.capacity)

    def draw_starts(self, batch_size, rng=np.random):
        """Buffer.py:36-48: uniform starts, ONE re-draw for windows that straddle the write head."""
        if self.size < self.sequence_length:
            raise ValueError("Not enough data in buffer to sample a full sequence")
        valid = self.size - self.sequence_length + 1
        starts = rng.randint(0, valid, size=batch_size)
        if self.size == self.capacity:
            fixed = []
            for s in starts:
                if s < self.next_idx < s + self.sequence_length:
                    fixed.append(rng.randint(0, valid))
                else:
                    fixed.append(s)
            starts = np.array(fixed)
        return np.asarray(starts, dtype=np.int64)

    def gather(self, starts):
        """Buffer.py:49-61: index grid modulo capacity, 4-array gather, obs as fp32 0..255."""
        idx = (np.asarray(starts)[:, None] + np.arange(self.sequence_length)[None, :]) % self.capacity
        return (self.obs[idx].astype(np.float32), self.act[idx], self.rew[idx], self.con[idx], idx)

    def sample_sequences(self, batch_size, rng=np.random):
        o, a, r, c, _ = self.gather(self.draw_starts(batch_size, rng))
        return o, a, r, c, self.sequence_length
