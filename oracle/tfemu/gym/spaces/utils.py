import numpy as np


def flatdim(space):
    from . import Box, Discrete
    if isinstance(space, Box):
        return int(np.prod(space.shape))
    if isinstance(space, Discrete):
        return int(space.n)
    raise NotImplementedError(type(space))
