"""Running normalisation statistics - host side, same interface as
``/root/reference/sac_eo/common/normalizer.py``.  ``normalize``/``denormalize`` are what every network
forward applies (they run INSIDE the CUDA kernels from the per-agent normaliser record; these host methods
exist for callers and tests); ``update`` is per-episode host bookkeeping (out of the hot path)."""
import numpy as np


class RunningNormalizer:
    def __init__(self, dim):
        self.dim = dim
        self.version = 0          # bumped by every change of the statistics: bound networks re-push them to the device
        self.reset()

    def reset(self):
        self.version += 1
        self.t_last = 0
        if self.dim == 1:                     # scalars stay Python floats (normalizer.py:17-24)
            self.mean, self.var, self.std = 0.0, 0.0, 1.0
        else:
            self.mean = np.zeros(self.dim, np.float32)
            self.var = np.zeros(self.dim, np.float32)
            self.std = np.ones(self.dim, np.float32)

    def _den(self):
        return np.maximum(self.std, 1e-8)

    def normalize(self, data, center=True):
        return (data - self.mean) / self._den() if center else data / self._den()

    def denormalize(self, data_norm, center=True):
        return data_norm * self._den() + self.mean if center else data_norm * self._den()

    def update(self, data):
        """Merges a batch into the running mean / unbiased variance with the reference's arithmetic (normalizer.py:55-87):
        the batch is first scaled by the CURRENT std, batch mean and sum of squares are taken in that scaled space and
        scaled back, everything in the dtype NumPy promotion gives (float32 for float32 trajectories), max(1, n-1)
        denominators, std = 1 while only one sample was seen.  Same operations in the same order, so the statistics are
        bit-identical to the reference's (tests/test_reference_pin.py)."""
        scale = self._den()
        scale_sq = np.square(scale)
        z = data / scale
        nb = z.shape[0]
        zb = z.mean(axis=0)
        ssq = np.sum(np.square(z - zb), axis=0)
        n0, n = self.t_last, self.t_last + nb
        var = (scale_sq * ssq + self.var * np.maximum(1, n0 - 1)
               + (nb / n) * n0 * scale_sq * np.square(zb - self.mean / scale)) / np.maximum(1, n - 1)
        mean = (nb * zb * scale + n0 * self.mean) / n
        self.mean, self.var = mean.astype("float32"), var.astype("float32")
        self.std = np.ones_like(self.var) if n == 1 else np.sqrt(self.var)
        self.t_last = n
        self.version += 1

    def instantiate(self, t, mean, var, ignore=None):
        self.version += 1
        self.t_last, self.mean, self.var = t, mean, var
        if t == 0:
            self.reset()
        elif t == 1:
            self.std = np.abs(self.mean)
        else:
            self.std = np.sqrt(self.var)

    def get_stats(self):
        return {"t": self.t_last, "mean": self.mean, "var": self.var}


class RunningNormalizers:
    """s / a / r / delta / return normalisers shared by actor, critics and models."""

    def __init__(self, s_dim, a_dim, gamma, init_rms_stats=None):
        self.gamma = gamma
        self.s_rms, self.a_rms = RunningNormalizer(s_dim), RunningNormalizer(a_dim)
        self.r_rms, self.delta_rms, self.ret_rms = RunningNormalizer(1), RunningNormalizer(s_dim), RunningNormalizer(1)
        self.set_rms_stats(init_rms_stats)

    def get_rms(self):
        return self.s_rms, self.a_rms, self.r_rms, self.delta_rms, self.ret_rms

    def update_rms(self, s_traj, a_traj, r_traj, sp_traj):
        self.s_rms.update(s_traj)
        self.a_rms.update(a_traj)
        self.r_rms.update(r_traj)
        self.delta_rms.update(sp_traj - s_traj)
        # discounted return-to-go in float64, the recurrence scipy's lfilter([1], [1, -gamma]) runs on the reversed
        # rewards (buffer_utils.py:8): y[i] = r[i] + gamma * y[i+1]
        r64 = np.asarray(r_traj, np.float64)
        ret, acc = np.zeros(len(r64)), np.float64(0.0)
        g = np.float64(self.gamma)
        for i in range(len(r64) - 1, -1, -1):
            acc = r64[i] + g * acc
            ret[i] = acc
        self.ret_rms.update(ret)

    def set_rms_stats(self, stats):
        if stats is not None:
            for k in ("s_rms", "a_rms", "r_rms", "delta_rms", "ret_rms"):
                getattr(self, k).instantiate(**stats[k])

    def get_rms_stats(self):
        return {k: getattr(self, k).get_stats() for k in ("s_rms", "a_rms", "r_rms", "delta_rms", "ret_rms")}
