"""pygame.surfarray stand-in: array3d returns a (W, H, 3) copy like pygame does."""
import numpy as np


def array3d(surface):
    return np.ascontiguousarray(np.transpose(surface.rgb, (1, 0, 2)))
