import numpy as np


def flatdim(space):
    return int(space.n) if hasattr(space, "n") else int(np.prod(space.shape))
