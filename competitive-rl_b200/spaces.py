"""Minimal gym.spaces stand-ins (gym is an unpinned dependency of the reference and is
absent from this image).  Same attribute surface as the gym classes the reference's
wrappers build (utils/atari_wrappers.py:12-37, 184-259): shape, dtype, low, high, n,
spaces, contains(), sample().  If a real `gym` is importable its classes are used."""
from collections import OrderedDict

import numpy as np

try:  # pragma: no cover - not available in the build image
    from gym.spaces import Box, Dict, Discrete, Tuple  # noqa: F401
except Exception:  # noqa: BLE001

    class Space(object):
        shape = None
        dtype = None

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.asarray(low).shape
            self.shape = tuple(int(s) for s in shape)
            self.low = np.full(self.shape, low, dtype=self.dtype) if np.isscalar(low) else \
                np.broadcast_to(np.asarray(low, self.dtype), self.shape).copy()
            self.high = np.full(self.shape, high, dtype=self.dtype) if np.isscalar(high) else \
                np.broadcast_to(np.asarray(high, self.dtype), self.shape).copy()

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

        def sample(self):
            if self.dtype.kind in "iu":
                return np.random.randint(self.low, self.high.astype(np.int64) + 1).astype(self.dtype)
            return np.random.uniform(self.low, self.high).astype(self.dtype)

        def __repr__(self):
            return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)

        def __eq__(self, other):
            return isinstance(other, Box) and self.shape == other.shape and self.dtype == other.dtype and \
                np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high)

    class Discrete(Space):
        def __init__(self, n):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.dtype(np.int64)

        def contains(self, x):
            try:
                return 0 <= int(x) < self.n and int(x) == x
            except Exception:  # noqa: BLE001
                return False

        def sample(self):
            return int(np.random.randint(self.n))

        def __repr__(self):
            return "Discrete(%d)" % self.n

        def __eq__(self, other):
            return isinstance(other, Discrete) and self.n == other.n

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

        def __repr__(self):
            return "Tuple(%s)" % ", ".join(repr(s) for s in self.spaces)

    class Dict(Space):
        def __init__(self, spaces):
            self.spaces = OrderedDict(spaces)

        def __getitem__(self, k):
            return self.spaces[k]

        def __len__(self):
            return len(self.spaces)

        def sample(self):
            return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

        def __repr__(self):
            return "Dict(%s)" % ", ".join("%r: %r" % kv for kv in self.spaces.items())
