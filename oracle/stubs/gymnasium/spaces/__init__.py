"""Stub `gymnasium.spaces` (Discrete / Box), see ../__init__.py."""
import numpy as np


class Space:
    def __init__(self, seed=None):
        self._np_random = None
        self.seed(seed)

    def seed(self, seed=None):
        if isinstance(seed, np.random.Generator):
            self._np_random = seed
        else:
            self._np_random = np.random.default_rng(seed)
        return [seed]

    @property
    def np_random(self):
        return self._np_random


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        self.n = int(n)
        self.start = int(start)
        super().__init__(seed)

    def sample(self, mask=None):
        if mask is not None:
            valid = np.flatnonzero(np.asarray(mask) == 1)
            if len(valid) == 0:
                return self.start
            return self.start + int(self._np_random.choice(valid))
        return self.start + int(self._np_random.integers(self.n))

    def contains(self, x):
        return self.start <= int(x) < self.start + self.n


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        low = np.asarray(low, dtype=dtype)
        high = np.asarray(high, dtype=dtype)
        if shape is not None:
            low = np.broadcast_to(low, shape).copy()
            high = np.broadcast_to(high, shape).copy()
        self.low, self.high = low, high
        self.shape = low.shape
        self.dtype = np.dtype(dtype)
        super().__init__(seed)

    def sample(self):
        return self._np_random.uniform(self.low, self.high).astype(self.dtype)


class MultiDiscrete(Space):
    def __init__(self, nvec, seed=None):
        self.nvec = np.asarray(nvec)
        super().__init__(seed)


class Dict(Space, dict):
    pass


class Tuple(Space):
    pass
