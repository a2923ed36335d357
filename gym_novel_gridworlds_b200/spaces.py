"""Plain value-holder spaces (the hot path never samples them; they exist so `env.action_space.n`
and `env.observation_space.shape` read as they do on the reference, stale values included — SURVEY Q11)."""
import numpy as np


class Space(object):
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = dtype


class Discrete(Space):
    def __init__(self, n):
        super().__init__((), np.int64)
        self.n = int(n)

    def __repr__(self):
        return "Discrete(%d)" % self.n


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.asarray(low).shape
        super().__init__(shape, dtype)
        self.low = np.broadcast_to(np.asarray(low), self.shape)
        self.high = np.broadcast_to(np.asarray(high), self.shape)


class Dict(Space):
    def __init__(self, spaces=None, **kw):
        super().__init__(None, None)
        self.spaces = dict(spaces or {}, **kw)
