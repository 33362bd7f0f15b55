"""``TRPO`` actor update with the reference's interface (``/root/reference/sac_eo/algs/model_free/trpo.py:8-317``):
surrogate gradient (+ entropy regulariser), conjugate-gradient solve against the average Fisher information matrix,
step length ``sqrt(2 delta / vFv)`` and the back-tracking line search - all on the device through
``saceo_trpo_grad`` / ``saceo_cg_solve`` / ``saceo_trpo_eval`` / ``saceo_actor_step`` (include/saceo.h); the host keeps
only the scalars the reference keeps in NumPy (advantage statistics, the accept / shrink decisions, the temperature).

The reference builds ``grad_final`` only inside its expert branches (:107-111, :154-158) as ``(1 - epsilon) neg_pg +
epsilon MSE_grads``: with ``expert_reg`` and two models this class runs that blend on the device
(``saceo_onpolicy_expert_grad`` + ``saceo_grad_blend``, same RNG consumption: ``rng.shuffle`` of the expert rows,
then one ``np.random.normal`` per half for ``actor.sample``); without ``expert_reg`` it implements the ``epsilon = 0``
slice (``grad_final = neg_pg``; the reference raises NameError there).  The single-model branch (:77-111) raises in the
reference (``epsilon * None`` for the temperature gradient, :111) and is refused here.
``trust_sub`` must be 1 (the Fisher rows are the rollout rows).  The temperature uses Keras-Adam on one scalar
(:29-32, :169-172), kept on the host like the reference's other per-update scalars."""
import numpy as np

from .base_mfrl_updates import BaseOnPolicyUpdate
from ...common.update_utils import make_F


class TRPO(BaseOnPolicyUpdate):
    def __init__(self, actor, update_kwargs, gemm_mode=0, device=0):
        super().__init__(actor, update_kwargs)
        self._gemm_mode, self._device = gemm_mode, device
        self._F = None

    def _setup(self, update_kwargs):
        self.adv_center = update_kwargs['adv_center']
        self.adv_scale = update_kwargs['adv_scale']
        self.delta = update_kwargs['delta_trpo']
        self.cg_it = update_kwargs['cg_it']
        self.trust_sub = update_kwargs['trust_sub']
        self.trust_damp = update_kwargs['trust_damp']
        self.kl_maxfactor = update_kwargs['kl_maxfactor']
        self.ent_reg = update_kwargs['ent_reg']
        self.ent_targ = update_kwargs['ent_targ']
        self.alpha = np.float32(0.0)
        self.alpha_lr = update_kwargs['alpha_lr']
        self._alpha_m, self._alpha_v, self._alpha_t = 0.0, 0.0, 0

    def _alpha_step(self, grad):
        """tf.keras Adam on the scalar temperature (beta 0.9 / 0.999, epsilon 1e-7 outside the bias correction),
        then ``alpha = max(alpha, 0)`` (:169-172)."""
        g = np.float32(grad)
        self._alpha_t += 1
        self._alpha_m = np.float32(0.9 * self._alpha_m + 0.1 * g)
        self._alpha_v = np.float32(0.999 * self._alpha_v + 0.001 * g * g)
        lr_t = np.float32(self.alpha_lr * np.sqrt(1 - 0.999 ** self._alpha_t) / (1 - 0.9 ** self._alpha_t))
        self.alpha = np.float32(max(self.alpha - lr_t * self._alpha_m / (np.sqrt(self._alpha_v) + 1e-7), 0.0))

    def update(self, rollout_data, expert_reg=None):
        if self.trust_sub != 1:
            raise ValueError("trust_sub must be 1: the Fisher rows are the rollout rows on the device path")
        s_all, a_all, adv_all = rollout_data[0], rollout_data[1], rollout_data[2]
        eps = None
        if expert_reg is not None:
            s_expert, _, sp_expert, epsilon, models, _, rng = expert_reg[:7]
            if len(models) != 2:
                raise NotImplementedError("TRPO.update with one model: the reference's own branch raises TypeError at "
                                          "trpo.py:111 ((1 - epsilon) * alpha_grad + epsilon * None)")
            F = make_F(self.actor, s_all, 1, self.trust_damp, gemm_mode=self._gemm_mode, device=self._device,
                       models=models, expert=(s_expert, sp_expert))
            idx = np.arange(len(s_expert))
            rng.shuffle(idx)                                                  # :115-116
            sections = np.array_split(idx, 2)
            u = [np.random.normal(size=(len(sec), self.actor.a_dim)) for sec in sections]   # actor.sample x2, :135-136
            F.set_expert_draws(np.concatenate(sections), np.concatenate(u))
            eps = float(epsilon)
        else:
            F = make_F(self.actor, s_all, 1, self.trust_damp, gemm_mode=self._gemm_mode, device=self._device)
        pop = F.pop
        logs = pop.trpo_update(np.asarray(a_all, np.float32)[None], np.asarray(adv_all, np.float32)[None],
                               delta=self.delta, cg_iters=self.cg_it, trust_damp=self.trust_damp,
                               kl_maxfactor=self.kl_maxfactor, alpha=np.asarray([self.alpha], np.float32),
                               adv_center=self.adv_center, adv_scale=self.adv_scale, expert_eps=eps)
        if self.ent_reg:
            # alpha_grad = -(mean entropy - ent_targ) at the pre-update policy; apply_gradients([alpha_grad * -1]) (:169-172).
            # The surrogate tape above saw the temperature before this step, as in the reference.
            self._alpha_step(logs[0]['ent'] - self.ent_targ)
        self.actor.set_weights(pop.get_net(0, "actor"))
        pop.close()
        log_actor = logs[0]
        log_actor['alpha'] = float(self.alpha)
        log_actor['epsilon'] = 0.0 if expert_reg is None else float(expert_reg[3])
        return log_actor
