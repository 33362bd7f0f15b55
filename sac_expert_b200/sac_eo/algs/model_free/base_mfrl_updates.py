"""``BaseOnPolicyUpdate`` (``/root/reference/sac_eo/algs/model_free/base_mfrl_updates.py:1-29``)."""


class BaseOnPolicyUpdate:
    def __init__(self, actor, update_kwargs):
        self._setup(update_kwargs)
        self.actor = actor

    def _setup(self, update_kwargs):
        raise NotImplementedError

    def update(self, rollout_data, expert_reg=None):
        raise NotImplementedError
