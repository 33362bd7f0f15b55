"""``TrajectoryBuffer`` with the reference's interface (``/root/reference/sac_eo/common/buffers.py``): host
NumPy arrays (``s_all`` ... ``current_size``) with append + tail truncation, mirrored row-by-row into the
device replay table of one agent once ``attach`` ed.  Minibatch sampling keeps the reference's RNG call
(``np.random.randint(current_size, size=batch_size)``, :135) on the host and performs the five row gathers
on the device (bit-exact copies).  GAE / ``get_update_info`` belong to the on-policy path (out of scope)."""
import numpy as np
import torch

from ...lib import SaceoError


class TrajectoryBuffer:
    def __init__(self, s_dim, a_dim, gamma, lam, buffer_size=None):
        self.s_dim, self.a_dim, self.gamma, self.lam, self.buffer_size = s_dim, a_dim, gamma, lam, buffer_size
        self._pop, self._agent = None, 0
        self.reset()

    def reset(self):
        self.s_all = np.empty((0, self.s_dim), np.float32)
        self.a_all = np.empty((0, self.a_dim), np.float32)
        self.r_all = np.empty((0,), np.float32)
        self.sp_all = np.empty((0, self.s_dim), np.float32)
        self.d_all = np.empty((0,))          # float64 done flags
        self.idx_all = np.empty((0,))
        self.traj_total = self.steps_total = self.current_size = 0

    def attach(self, pop, agent=0):
        """Mirror this buffer into the device replay table of ``agent`` (existing rows are uploaded)."""
        if self.buffer_size and self.buffer_size > pop.spec.replay_capacity:
            raise ValueError("device replay_capacity is smaller than buffer_size")
        self._pop, self._agent = pop, agent
        if self.current_size:
            pop.append_rows(agent, self.s_all, self.a_all, self.r_all, self.sp_all, self.d_all)

    def add(self, s_traj, a_traj, r_traj, sp_traj, d_traj):
        s_traj = np.asarray(s_traj).reshape(-1, self.s_dim)
        a_traj = np.asarray(a_traj).reshape(-1, self.a_dim)
        r_traj, d_traj = np.atleast_1d(np.asarray(r_traj)), np.atleast_1d(np.asarray(d_traj))
        sp_traj = np.asarray(sp_traj).reshape(-1, self.s_dim)
        cat = lambda old, new: np.concatenate((old, new), axis=0)
        self.s_all, self.a_all = cat(self.s_all, s_traj), cat(self.a_all, a_traj)
        self.r_all, self.sp_all = cat(self.r_all, r_traj), cat(self.sp_all, sp_traj)
        self.d_all = cat(self.d_all, d_traj)
        self.idx_all = cat(self.idx_all, np.ones_like(r_traj) * self.traj_total)
        if self.buffer_size and len(self.r_all) > self.buffer_size:        # keep the LAST buffer_size rows (:60-66)
            keep = slice(-self.buffer_size, None)
            self.s_all, self.a_all, self.r_all = self.s_all[keep], self.a_all[keep], self.r_all[keep]
            self.sp_all, self.d_all, self.idx_all = self.sp_all[keep], self.d_all[keep], self.idx_all[keep]
        self.current_size = len(self.r_all)
        self.traj_total += 1
        self.steps_total += len(r_traj)
        if self._pop is not None:
            if not self.buffer_size and self.current_size > self._pop.spec.replay_capacity:
                raise SaceoError("unbounded TrajectoryBuffer outgrew the device replay_capacity")
            self._pop.append_rows(self._agent, s_traj, a_traj, r_traj, sp_traj, d_traj)

    def _device_gather(self, idx):
        if self._pop is None:
            raise SaceoError("TrajectoryBuffer is not attached to a device population; minibatch gathers have no CPU path")
        if len(idx) != self._pop.spec.B:
            raise ValueError(f"batch_size {len(idx)} != device batch {self._pop.spec.B}")
        full = np.zeros((self._pop.spec.n_agents, len(idx)), np.int64)
        full[self._agent] = idx
        out = self._pop.gather(torch.from_numpy(full))
        return [t[self._agent].cpu().numpy() for t in out]

    def get_offmodel_info(self, batch_size=None):
        if batch_size:
            idx = np.random.randint(self.current_size, size=batch_size)
            self.last_idx = idx
            s, a, sp, r, d = self._device_gather(idx)
            return s, a, sp, r, d
        return self.s_all, self.a_all, self.sp_all, self.r_all, self.d_all

    def get_model_info(self, batch_size=None):
        if batch_size:
            idx = np.random.randint(self.current_size, size=batch_size)
            if self._pop is not None and batch_size == self._pop.spec.B:
                s, a, sp, r, _ = self._device_gather(idx)
                return s, a, sp, r
            # small per-episode resample of the expert buffer (SAC_expert.py:421-422): host bookkeeping, not the
            # update hot path - the rows are re-uploaded as expert_reg afterwards
            return self.s_all[idx], self.a_all[idx], self.sp_all[idx], self.r_all[idx]
        return self.s_all, self.a_all, self.sp_all, self.r_all

    def get_states(self, batch_size=None):
        if batch_size:
            return self.get_offmodel_info(batch_size)[0]
        return self.s_all
