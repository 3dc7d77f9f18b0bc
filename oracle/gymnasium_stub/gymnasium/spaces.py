import numpy as np


class Space:
    pass


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        self.n = int(n)
        self.start = int(start)
        self.shape = ()
        self.dtype = np.int64
        self._rng = np.random.default_rng(seed)

    def sample(self, mask=None):
        if mask is not None:
            idx = np.flatnonzero(mask)
            return int(self.start + self._rng.choice(idx))
        return int(self.start + self._rng.integers(self.n))

    def contains(self, x):
        return self.start <= int(x) < self.start + self.n


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        self.low = low
        self.high = high
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.dtype = dtype
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return self._rng.uniform(self.low, self.high, size=self.shape).astype(self.dtype)

    def contains(self, x):
        return np.shape(x) == self.shape
