"""``init_seeds`` (``/root/reference/sac_eo/common/seeding.py:7-16``).  The reference also calls ``tf.random.set_seed``;
nothing on this path draws from TensorFlow's generator (all draws are NumPy), so it has no analogue."""
import os
import random

import numpy as np


def init_seeds(seed, env=None):
    seed = int(seed)
    if env is not None:
        env.seed(seed)
    np.random.seed(seed)
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
