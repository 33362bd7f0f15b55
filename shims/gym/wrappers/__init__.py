import numpy as np


class RescaleAction:
    """Affine map of actions from [min_action, max_action] to the wrapped environment's bounds."""

    def __init__(self, env, min_action, max_action):
        self.env, self.min_action, self.max_action = env, float(min_action), float(max_action)
        self.observation_space = env.observation_space
        from ..spaces import Box
        self.action_space = Box(min_action, max_action, env.action_space.shape)

    def __getattr__(self, name):
        return getattr(self.env, name)

    def step(self, action):
        lo, hi = self.env.action_space.low, self.env.action_space.high
        a = lo + (hi - lo) * ((np.asarray(action) - self.min_action) / (self.max_action - self.min_action))
        return self.env.step(np.clip(a, lo, hi))
