import os
from collections import OrderedDict

import numpy as np


class Space(object):
    shape = None
    dtype = None


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=None):
        if dtype is None:
            if os.environ.get("GYM_SHIM_BOX_FLOAT32") == "1":
                dtype = np.float32
            else:
                dtype = np.uint8 if np.all(np.asarray(high) == 255) else np.float32
        self.dtype = np.dtype(dtype)
        if shape is None:
            low, high = np.asarray(low), np.asarray(high)
            shape = low.shape
        self.shape = tuple(int(s) for s in shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high)

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def __repr__(self):
        return "Box" + str(self.shape)


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def contains(self, x):
        if isinstance(x, (int, np.integer)):
            return 0 <= int(x) < self.n
        x = np.asarray(x)
        return x.shape == () and x.dtype.kind in "iu" and 0 <= int(x) < self.n

    def sample(self):
        return int(np.random.randint(self.n))

    def __repr__(self):
        return "Discrete(%d)" % self.n


class Tuple(Space):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return len(self.spaces)

    def contains(self, x):
        return len(x) == len(self.spaces) and all(s.contains(p) for s, p in zip(self.spaces, x))

    def sample(self):
        return tuple(s.sample() for s in self.spaces)


class Dict(Space):
    def __init__(self, spaces):
        self.spaces = OrderedDict(spaces)

    def __getitem__(self, k):
        return self.spaces[k]

    def __len__(self):
        return len(self.spaces)
