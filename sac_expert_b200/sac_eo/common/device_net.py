"""Base of the reference-compatible network objects: weights live on the host until the algorithm binds the
object to (population, agent); from then on every accessor goes through the device tables and every forward
through the C ABI (no CPU compute path)."""
import numpy as np
import torch

from ...lib import SaceoError
from .nn_utils import flat_to_list, list_to_flat


class DeviceNet:
    _table = None          # population net name: "actor", "q1", "t1", "m1", ...

    def __init__(self):
        self._pop, self._agent = None, 0
        self._host_weights = None
        self._rms = None

    # ---- binding -------------------------------------------------------------------------
    def _bind(self, pop, agent, table):
        self._pop, self._agent, self._table = pop, agent, table
        pop.set_net(agent, table, self._host_weights)
        self._push_rms()
        self._rms_pushed = self._rms_versions()

    def _need_device(self):
        if self._pop is None:
            raise SaceoError(f"{type(self).__name__} is not bound to a device population yet "
                             "(construct the algorithm first); there is no CPU forward path")
        return self._pop

    def _push_rms(self):
        pass

    def _rms_versions(self):
        return None if self._rms is None else tuple(getattr(r, "version", 0) for r in self._rms)

    def _sync_rms(self):
        """The reference's networks hold live references to the RunningNormalizer objects, so ``update_rms`` takes
        effect on the very next forward / update (normalizer.py:60-110).  Here the statistics live in the device
        normaliser record: re-push them whenever a bound normaliser changed since the last push."""
        v = self._rms_versions()
        if self._pop is not None and v != getattr(self, "_rms_pushed", None):
            self._push_rms()
            self._rms_pushed = v

    # ---- weights (Keras get_weights()/set_weights() order) -------------------------------
    def _shapes(self):
        return [w.shape for w in self._host_weights]

    def _get_list(self):
        if self._pop is not None:
            return self._pop.get_net(self._agent, self._table)
        return [w.copy() for w in self._host_weights]

    def _set_list(self, weights):
        weights = [np.asarray(w, np.float32) for w in weights]
        if [w.shape for w in weights] != [tuple(s) for s in self._shapes()]:
            raise ValueError(f"weight shapes {[w.shape for w in weights]} do not match {self._shapes()}")
        self._host_weights = weights
        if self._pop is not None:
            self._pop.set_net(self._agent, self._table, weights)

    def get_weights(self, flat=False):
        ws = self._get_list()
        return list_to_flat(ws) if flat else ws

    def set_weights(self, weights, from_flat=False, increment=False):
        if from_flat:
            weights = flat_to_list(self._shapes(), np.asarray(weights))
        if increment:
            weights = [a + b for a, b in zip(weights, self._get_list())]
        self._set_list(weights)

    @staticmethod
    def _as_rows(x, dim):
        """``transform_features`` (nn_utils.py:78-84): cast to f32, add the batch dimension."""
        x = np.asarray(x, np.float32)
        return x.reshape(1, dim) if x.ndim == 1 else x.reshape(-1, dim)

    @staticmethod
    def _host(t: torch.Tensor):
        """Device result -> CPU tensor (has ``.numpy()``, which is what reference call sites use)."""
        return t.detach().cpu()
