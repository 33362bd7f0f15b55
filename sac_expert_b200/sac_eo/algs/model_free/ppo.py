"""``PPO`` actor update with the reference's interface (``/root/reference/sac_eo/algs/model_free/ppo.py:6-238``):
``actor_update_it`` epochs over ``actor_nminibatch`` shuffled minibatches; per minibatch the clipped-surrogate gradient
with the entropy regulariser and ``clip_by_global_norm`` (``saceo_ppo_grad``) and one Keras-Adam step of the actor
(``saceo_actor_adam``), then the tv / kl statistics of the whole rollout (``saceo_trpo_eval``) and the adaptive learning
rate.  The host keeps what the reference keeps in NumPy: the shuffle (``np.random.shuffle``, same global-RNG
consumption), the per-minibatch advantage statistics, the temperature scalar and the learning-rate adaptation.

With ``expert_reg`` the loss becomes ``(1 - epsilon) pg_loss + epsilon MSE`` (:176-213): the MSE gradient through
``tf_clip(actor.sample(s_expert))`` and ``models[0]`` comes from ``saceo_onpolicy_expert_grad``, the blend and the
global-norm clip from ``saceo_grad_blend``; one ``np.random.normal`` draw per minibatch step, like the reference.  The
reference's ``use_expert_actions`` branch reads ``sp_pred`` before assignment (:196-199 vs :211-212) and is refused."""
import numpy as np
import torch

from .base_mfrl_updates import BaseOnPolicyUpdate
from ...common.update_utils import make_F


class PPO(BaseOnPolicyUpdate):
    def __init__(self, actor, update_kwargs, gemm_mode=0, device=0):
        super().__init__(actor, update_kwargs)
        self._gemm_mode, self._device = gemm_mode, device
        self._m = self._v = None          # actor optimiser slots persist across update() calls like the Keras optimiser's
        self._t = 0

    def _setup(self, update_kwargs):
        self.actor_lr = update_kwargs['actor_lr']
        self.actor_update_it = update_kwargs['actor_update_it']
        self.actor_nminibatch = update_kwargs['actor_nminibatch']
        self.adv_center = update_kwargs['adv_center']
        self.adv_scale = update_kwargs['adv_scale']
        self.eps = update_kwargs['eps_ppo']
        self.max_grad_norm = update_kwargs['max_grad_norm']
        self.adaptlr = update_kwargs['adaptlr']
        self.adapt_factor = update_kwargs['adapt_factor']
        self.adapt_minthresh = update_kwargs['adapt_minthresh']
        self.adapt_maxthresh = update_kwargs['adapt_maxthresh']
        self.ent_reg = update_kwargs['ent_reg']
        self.ent_targ = update_kwargs['ent_targ']
        self.alpha = np.float32(0.0)
        self.alpha_lr = update_kwargs['alpha_lr']
        self._alpha_m, self._alpha_v, self._alpha_t = 0.0, 0.0, 0

    def _alpha_step(self, grad):
        """tf.keras Adam on the scalar temperature, then ``alpha = max(alpha, 0)`` (:221-224)."""
        g = np.float32(grad)
        self._alpha_t += 1
        self._alpha_m = np.float32(0.9 * self._alpha_m + 0.1 * g)
        self._alpha_v = np.float32(0.999 * self._alpha_v + 0.001 * g * g)
        lr_t = np.float32(self.alpha_lr * np.sqrt(1 - 0.999 ** self._alpha_t) / (1 - 0.9 ** self._alpha_t))
        self.alpha = np.float32(max(self.alpha - lr_t * self._alpha_m / (np.sqrt(self._alpha_v) + 1e-7), 0.0))

    def _population(self, s_rows, expert_reg=None):
        kw = {}
        if expert_reg is not None:
            kw = dict(models=list(expert_reg[4])[:1], expert=(expert_reg[0], expert_reg[2]))     # models[0] only, ppo.py:202
        F = make_F(self.actor, s_rows, 1, 0.0, gemm_mode=self._gemm_mode, device=self._device, **kw)
        self._F = F
        pop = F.pop
        pop.set_hyper(0, lr_pi=self.actor_lr)
        if self._m is not None:
            pop.t["actor_m"].copy_(self._m); pop.t["actor_v"].copy_(self._v)
            pop.t["adam_t"][0, 2] = self._t
        return pop

    def update(self, rollout_data, expert_reg=None):
        if expert_reg is not None and expert_reg[5]:
            raise NotImplementedError("use_expert_actions: the reference's branch reads sp_pred before assignment "
                                      "(ppo.py:196-199 vs :211-212) and raises UnboundLocalError")
        eps_exp = None if expert_reg is None else float(expert_reg[3])
        s_all = np.asarray(rollout_data[0], np.float32)
        a_all = np.asarray(rollout_data[1], np.float32)
        adv_all = np.asarray(rollout_data[2])
        n_samples = s_all.shape[0]
        n_batch = int(n_samples / self.actor_nminibatch)

        # whole-rollout reference quantities (:44-46)
        full = self._population(s_all)
        first = full.trpo_eval(act=a_all[None], want_nlp=True, want_kl_info=True)
        nlp_old_dev, kl_ref = first["nlp"], first["kl_info"]
        nlp_old_all = nlp_old_dev[0].cpu().numpy()
        ent = float(first["stats"][0, 3])

        mb = self._population(s_all[:n_batch], expert_reg)
        mbF = self._F
        mb.t["actor"].copy_(full.t["actor"])
        pre_all = post_all = 0.0
        nb = 0
        for _ in range(self.actor_update_it):
            idx = np.arange(n_samples)
            np.random.shuffle(idx)                                          # :56-57
            sections = np.arange(0, n_samples, n_batch)[1:]
            batches = np.array_split(idx, sections)
            if n_samples % n_batch != 0:
                batches = batches[:-1]
            for b in batches:
                adv = adv_all[b]
                adv_mean, adv_std = np.mean(adv), np.std(adv) + 1e-8        # :71-77
                if self.adv_center:
                    adv = adv - adv_mean
                if self.adv_scale:
                    adv = adv / adv_std
                mb.t["fvp_states"][0].copy_(torch.from_numpy(s_all[b]))
                if eps_exp is None:
                    grad, stats = mb.ppo_grad(a_all[b][None], np.asarray(adv, np.float32)[None], nlp_old_all[b][None],
                                              np.asarray([self.alpha], np.float32), self.eps, self.max_grad_norm)
                    stats = stats[0].cpu().numpy()
                    w_pg = 1.0
                else:
                    # pg_loss = (1 - epsilon) pg_loss + epsilon MSE(models[0].sample(sE, tf_clip(actor.sample(sE)))) (:190-213);
                    # the clip by global norm follows the blend (:226-231)
                    E = len(expert_reg[0])
                    mbF.set_expert_draws(np.arange(E), np.random.normal(size=(E, self.actor.a_dim)))      # actor.sample, :191
                    grad, stats = mb.ppo_grad(a_all[b][None], np.asarray(adv, np.float32)[None], nlp_old_all[b][None],
                                              np.asarray([self.alpha], np.float32), self.eps, None)
                    g_mse, _ = mb.onpolicy_expert_grad(n_models=1, clip_actions=True)
                    grad, bst = mb.grad_blend(grad, g_mse, eps_exp, self.max_grad_norm)
                    stats = stats[0].cpu().numpy()
                    stats[4:6] = bst[0, 4:6].cpu().numpy()
                    w_pg = 1.0 - eps_exp
                if self.ent_reg:                                            # :221-224, alpha_grad = -(1 - eps)(ent - ent_targ)
                    self._alpha_step(w_pg * (float(stats[3]) - self.ent_targ))
                mb.actor_adam(grad)                                         # :234
                pre_all += float(stats[4]); post_all += float(stats[5])
                nb += 1
        self._m, self._v = mb.t["actor_m"].clone(), mb.t["actor_v"].clone()
        self._t = int(mb.t["adam_t"][0, 2])

        full.t["actor"].copy_(mb.t["actor"])
        last = full.trpo_eval(act=a_all[None], nlp_old=nlp_old_dev, kl_ref=kl_ref, want_rows=True)
        tv, kl = float(last["stats"][0, 2]), float(last["stats"][0, 1])
        outside = float((last["rows"][0, :, 2] > self.eps).float().mean())
        self.actor.set_weights(full.get_net(0, "actor"))
        full.close(); mb.close()

        log_actor = {'ent': ent, 'tv': tv, 'kl': kl, 'alpha': float(self.alpha), 'actor_lr': float(self.actor_lr),
                     'outside_clip': outside, 'actor_grad_norm_pre': pre_all / max(nb, 1),
                     'actor_grad_norm': post_all / max(nb, 1)}
        if self.adaptlr:                                                    # :107-116
            if tv > (self.adapt_maxthresh * 0.5 * self.eps):
                self.actor_lr = self.actor_lr / (1 + self.adapt_factor)
            elif tv < (self.adapt_minthresh * 0.5 * self.eps):
                self.actor_lr = self.actor_lr * (1 + self.adapt_factor)
        return log_actor
