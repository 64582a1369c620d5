"""gym.utils.seeding stand-in.  gym hashes the seed with sha512 and seeds a RandomState with the
32-bit words; the exact stream is irrelevant to the oracle (track draws are recorded and injected),
so big seeds are simply split into 32-bit words."""
import numpy as np


def np_random(seed=None):
    rng = np.random.RandomState()
    if seed is not None:
        s, words = int(seed), []
        while True:
            words.append(s & 0xFFFFFFFF)
            s >>= 32
            if s == 0:
                break
        rng.seed(words)
    return rng, seed
