"""Value-holder spaces (the reference only constructs them; nothing on the hot path samples them)."""
import numpy as np


class Space(object):
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = dtype


class Discrete(Space):
    def __init__(self, n):
        self.n = n
        Space.__init__(self, (), np.int64)

    def sample(self):
        return int(np.random.randint(self.n))


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.asarray(low).shape
        self.low = np.broadcast_to(np.asarray(low), shape)
        self.high = np.broadcast_to(np.asarray(high), shape)
        Space.__init__(self, shape, dtype)


class Dict(Space):
    def __init__(self, spaces=None, **kw):
        self.spaces = dict(spaces or {}, **kw)
        Space.__init__(self, None, None)
