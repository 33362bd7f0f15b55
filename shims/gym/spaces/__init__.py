import numpy as np

from . import utils


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        shape = tuple(shape) if shape is not None else np.shape(low)
        self.shape, self.dtype = shape, dtype
        self.low = np.broadcast_to(np.asarray(low, dtype), shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype), shape).copy()

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)


class Discrete:
    def __init__(self, n):
        self.n, self.shape = int(n), ()

    def sample(self):
        return int(np.random.randint(self.n))
