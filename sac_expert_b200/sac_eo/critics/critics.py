"""``QCritic`` with the reference's interface (``/root/reference/sac_eo/critics/critics.py:60-110``):
``_forward`` returns the NORMALISED-space output [B, 1], ``value`` multiplies by ``max(ret_std, 1e-8)``."""
import numpy as np
import torch

from ..common.device_net import DeviceNet
from ..common.nn_utils import broadcast_activations, check_two_hidden, create_nn_weights
from ..envs.synthetic import flatdim


class QCritic(DeviceNet):
    def __init__(self, env, layers, activations, gain):
        super().__init__()
        self.s_dim, self.a_dim = flatdim(env.observation_space), flatdim(env.action_space)
        self.layers = check_two_hidden(layers)
        self.activations = broadcast_activations(layers, activations)
        # critic_init_type / critic_layer_norm are ignored by the reference as well (critics.py:74)
        self._host_weights = create_nn_weights(self.s_dim + self.a_dim, 1, self.layers, gain)
        self.trainable = ["W0", "b0", "W1", "b1", "W2", "b2"]

    def set_rms(self, normalizer):
        self._rms = normalizer.get_rms()
        self.s_rms, self.a_rms, _, _, self.ret_rms = self._rms
        self._push_rms()
        self._rms_pushed = self._rms_versions()

    def _push_rms(self):
        if self._pop is not None and self._rms is not None:
            self._pop.set_norm(self._agent, s_mean=self.s_rms.mean, s_std=self.s_rms.std, a_mean=self.a_rms.mean,
                               a_std=self.a_rms.std, ret_std=self.ret_rms.std)

    def _q(self, s, a, scale):
        pop = self._need_device()
        self._sync_rms()
        target = self._table.startswith("t")
        net = int(self._table[1]) - 1
        s_, a_ = self._as_rows(s, self.s_dim), self._as_rows(np.asarray(a), self.a_dim)
        q = pop.critic_forward(torch.from_numpy(s_)[None], torch.from_numpy(a_)[None], target=target, scale_ret=scale)
        return self._host(q[0, net])

    def _forward(self, s, a):
        return self._q(s, a, False)[:, None]

    def value(self, s, a):
        return self._q(s, a, True)
