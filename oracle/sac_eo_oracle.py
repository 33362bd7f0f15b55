"""CPU oracle for the SAC-EO gradient-update hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``sac_expert_b200``) must never import it and
fails loudly when its CUDA library is missing.

PARITY STATUS: PINNED TO THE REFERENCE'S PYTHON, NOT TO TENSORFLOW'S KERNELS.  The reference (noc-lab/sac-expert) is
TensorFlow-2-eager Python and ships no tests, golden vectors or known-answer fixtures for this path; TensorFlow and gym
are not installable in the build container (no wheels, no network).  This oracle is a *restatement* of the reference
algorithm in PyTorch-CPU (eager, autograd standing in for ``tf.GradientTape``) that follows the reference's operation
order line by line, citing each source location.  Since round 2 it is pinned against outputs of the reference's OWN,
UNMODIFIED code: ``oracle/tfemu`` executes the ~55 TensorFlow / gym symbols the reference imports on torch-CPU, the
reference's classes are imported from /root/reference and run over it (``tests/golden/make_golden_reference.py``), and
``tests/test_reference_pin.py`` holds this oracle (and the CUDA path) to the resulting vectors ``tests/golden/ref_*.npz``:
SAC_exp._update / SAC._update (3 consecutive updates, 1 / 2 / 0 models, both std parameterisations), BC._update,
TRPO.update (gradient, expert blend, Fisher-vector product, fp32 CG, line search), PPO.update, _update_models +
the adaptive expert weight, and the pure-NumPy TrajectoryBuffer / RunningNormalizers (bit-exact, no emulation involved).
What stays a restatement is the semantics of the TensorFlow PRIMITIVES themselves (Keras Adam's formula, reduce_min /
clip_by_value / maximum tie gradients, NumPy -> tensor dtype conversion, Dense) - listed with their TF sources in
``oracle/tfemu/tensorflow/__init__.py``; a real-TensorFlow cross-check hook exists (``oracle/tf_reference.py``) but has
never met a TensorFlow install.  Also checked against reference outputs: the GaussianActor entropy / logstd
parameterisation and the TRPO line-search control flow against the TRPO log the reference recorded in
sac_eo/logs/TEMPLOG_0 (tests/golden/templog0_trpo_log.json).  Own cross-checks (tests/test_oracle.py): fp32 vs fp64
twin, autograd vs hand-derived analytic backward (``analytic_update``), Fisher-vector product in double-backprop form vs
J^T M J form, one unit test per reference quirk.

All citations are relative to /root/reference/ (read-only, absent on the GPU box).

Weight-list layout (``sac_eo/common/nn_utils.py:86-138``): Keras ``Sequential`` of
``Dense`` layers, ``y = x @ W + b`` with ``W: [in, out]``; ``get_weights()`` order is
``[W0, b0, W1, b1, W2, b2]`` (+ ``logstd[1, A]`` last for a state-independent-std actor,
``sac_eo/actors/continuous_actors.py:201-209``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

LOG2PI = math.log(2.0 * math.pi)
LOG2 = math.log(2.0)
MIN_LOG_STD = -5.0   # continuous_actors.py:250
MAX_LOG_STD = 2.0    # continuous_actors.py:251
ADAM_B1 = 0.9
ADAM_B2 = 0.999
ADAM_EPS = 1e-7      # tf.keras.optimizers.Adam default, SAC_expert.py:108-115

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def act_fn(name: str):
    """``create_activations`` (nn_utils.py:5-22): tanh / relu / elu only."""
    if name == "tanh":
        return torch.tanh
    if name == "relu":
        return torch.relu
    if name == "elu":
        return torch.nn.functional.elu
    raise ValueError("activations must be tanh, relu or elu")


def mlp(theta: Sequence[Tensor], x: Tensor, acts: Sequence[str]) -> Tensor:
    """Keras Sequential of Dense layers (nn_utils.py:101-136): hidden layers carry an
    activation, the final layer is linear."""
    n_layers = len(theta) // 2 if len(theta) % 2 == 0 else (len(theta) - 1) // 2
    h = x
    for l in range(n_layers):
        h = h @ theta[2 * l] + theta[2 * l + 1]
        if l < n_layers - 1:
            h = act_fn(acts[l])(h)
    return h


def normalize(x: Tensor, mean, std, center: bool = True) -> Tensor:
    """``RunningNormalizer.normalize`` (normalizer.py:26-41)."""
    den = torch.clamp(torch.as_tensor(std, dtype=x.dtype), min=1e-8)
    if center:
        return (x - torch.as_tensor(mean, dtype=x.dtype)) / den
    return x / den


def denormalize(x: Tensor, mean, std, center: bool = True) -> Tensor:
    """``RunningNormalizer.denormalize`` (normalizer.py:43-58)."""
    den = torch.clamp(torch.as_tensor(std, dtype=x.dtype), min=1e-8)
    if center:
        return x * den + torch.as_tensor(mean, dtype=x.dtype)
    return x * den


def gather(replay: Dict[str, np.ndarray], idx: np.ndarray):
    """``TrajectoryBuffer.get_offmodel_info`` (buffers.py:135-142): five fancy-index row
    gathers with the SAME idx (uniform with replacement)."""
    return (replay["s"][idx], replay["a"][idx], replay["sp"][idx], replay["r"][idx],
            replay["d"][idx])


@dataclass
class NetCfg:
    """Static description of one agent (what the reference passes through its kwargs)."""
    S: int
    A: int
    actor_hidden: Tuple[int, int] = (256, 256)
    critic_hidden: Tuple[int, int] = (256, 256)
    model_hidden: Tuple[int, int] = (512, 512)
    actor_acts: Tuple[str, str] = ("relu", "relu")
    critic_acts: Tuple[str, str] = ("relu", "relu")
    model_acts: Tuple[str, str] = ("relu", "relu")
    per_state_std: bool = True
    separate_reward_nn: bool = False      # model output S (True) or S+1 (False)
    delta_clip_pred: float = 0.0          # 0/None => no clip (base_world_model.py:80-82)
    num_models: int = 2                   # 0 => plain SAC (SAC.py), 1 or 2 => SAC-EO branches
    std_mult: float = 1.0                 # only used by the GaussianActor._forward (CG) path
    reward_hidden: Optional[Tuple[int, int]] = None   # separate_reward_nn: the reward net (None: like the model net)
    reward_acts: Optional[Tuple[str, str]] = None

    @property
    def Ao(self) -> int:
        return 2 * self.A if self.per_state_std else self.A

    @property
    def model_out(self) -> int:
        return self.S if self.separate_reward_nn else self.S + 1


def head(cfg: NetCfg, theta_pi: Sequence[Tensor], x: Tensor, u: Optional[Tensor], st: Dict,
         deterministic: bool = False) -> Tuple[Tensor, Tensor]:
    """``SquashedGaussianActor.evaluate`` / ``.sample`` (continuous_actors.py:270-306,
    327-379).  No ``logstd_init``/``std_mult``; logstd clipped to [-5, 2]; tanh squash;
    neglogp = Gaussian part + 2(log2 - a - softplus(-2a)) correction."""
    dt = x.dtype
    out = mlp(theta_pi, normalize(x, st["s_mean"], st["s_std"]).to(dt), cfg.actor_acts)
    if cfg.per_state_std:
        mean, logstd = torch.split(out, cfg.A, dim=-1)
    else:
        mean = out
        logstd = theta_pi[6] * torch.ones_like(mean)
    logstd = torch.clamp(logstd, MIN_LOG_STD, MAX_LOG_STD)
    std = torch.exp(logstd)
    if deterministic:
        z = mean
    else:
        z = mean + std * u.to(dt)
    nlp_vec = ((z - mean) / torch.exp(logstd)) ** 2 + 2 * logstd + LOG2PI
    nlp = 0.5 * nlp_vec.sum(-1)
    corr = 2.0 * (LOG2 - z - torch.nn.functional.softplus(-2.0 * z))
    nlp = nlp + corr.sum(-1)
    pi = torch.as_tensor(st["act_limit"], dtype=dt) * torch.tanh(z)
    return pi, nlp


def q_forward(cfg: NetCfg, theta: Sequence[Tensor], s_: Tensor, a_: Tensor, st: Dict) -> Tensor:
    """``QCritic._forward`` (critics.py:84-94) -> [B, 1] in *normalised* return space."""
    sa = torch.cat([normalize(s_, st["s_mean"], st["s_std"]),
                    normalize(a_, st["a_mean"], st["a_std"])], -1)
    return mlp(theta, sa.to(s_.dtype), cfg.critic_acts)


def q_value(cfg: NetCfg, theta, s_, a_, st) -> Tensor:
    """``QCritic.value`` (critics.py:96-103): squeeze * max(ret_std, 1e-8)."""
    v = q_forward(cfg, theta, s_, a_, st).squeeze(-1)
    return denormalize(v, 0.0, st["ret_std"], center=False)


def model_sample(cfg: NetCfg, theta_m, s_: Tensor, a_: Tensor, st: Dict) -> Tensor:
    """``MSEModel.sample`` / ``GaussianModel.sample(deterministic=True)``
    (continuous_models.py:244-254, 56-70) via ``BaseWorldModel._forward``
    (base_world_model.py:65-87).  The model may carry its own normaliser set
    (``only_model_normalizer``, SAC_expert.py:139-144): keys prefixed ``m_``."""
    sa = torch.cat([normalize(s_, st["m_s_mean"], st["m_s_std"]),
                    normalize(a_, st["m_a_mean"], st["m_a_std"])], -1)
    pred = mlp(theta_m, sa.to(s_.dtype), cfg.model_acts)
    delta = pred if cfg.separate_reward_nn else pred[:, :-1]
    if cfg.delta_clip_pred:
        delta = torch.clamp(delta, -cfg.delta_clip_pred, cfg.delta_clip_pred)
    return s_ + denormalize(delta, st["m_d_mean"], st["m_d_std"])


def keras_adam(theta: List[Tensor], grads: List[Tensor], m: List[Tensor], v: List[Tensor],
               t: int, lr: float):
    """tf.keras Adam, one ``apply_gradients`` call (SAC_expert.py:243,250,338,347):
    epsilon 1e-7 OUTSIDE the bias correction (differs from torch.optim.Adam)."""
    t = t + 1
    dt = theta[0].dtype
    lr_t = lr * math.sqrt(1.0 - ADAM_B2 ** t) / (1.0 - ADAM_B1 ** t)
    new_theta, new_m, new_v = [], [], []
    for p, g, mi, vi in zip(theta, grads, m, v):
        mi = ADAM_B1 * mi + (1.0 - ADAM_B1) * g
        vi = ADAM_B2 * vi + (1.0 - ADAM_B2) * g * g
        p = p - torch.as_tensor(lr_t, dtype=dt) * mi / (torch.sqrt(vi) + ADAM_EPS)
        new_theta.append(p.detach())
        new_m.append(mi.detach())
        new_v.append(vi.detach())
    return new_theta, new_m, new_v, t


def polyak(target: List[Tensor], live: List[Tensor], tau: float) -> List[Tensor]:
    """``_update_q_target`` (SAC_expert.py:362-373): NumPy fp32
    ``target*(1-tau) + live*tau`` - two rounded products and a rounded sum."""
    dt = target[0].dtype
    one_m = torch.as_tensor(1.0 - tau, dtype=dt)
    tau_t = torch.as_tensor(tau, dtype=dt)
    return [tg * one_m + lv * tau_t for tg, lv in zip(target, live)]


# --------------------------------------------------------------------------------------
# one full update
# --------------------------------------------------------------------------------------
def _req(ts):
    return [t.detach().clone().requires_grad_(True) for t in ts]


def sac_eo_update(cfg: NetCfg, state: Dict, batch: Dict, hyper: Dict) -> Dict:
    """One ``SAC_exp._update`` (SAC_expert.py:463-477) - or ``SAC._update``
    (SAC.py:236-250) when ``cfg.num_models == 0`` - with every random draw injected.

    state: 'actor','q1','q2','t1','t2','m1','m2' weight lists; 'alpha' scalar tensor;
           'adam_<net>' = dict(m=list, v=list, t=int) for net in q1,q2,actor,alpha;
           normaliser stats: s_mean,s_std,a_mean,a_std,ret_std and the model set
           m_s_mean,m_s_std,m_a_mean,m_a_std,m_d_mean,m_d_std; act_limit.
    batch: s,a,sp,r,d (already gathered; d may be float64), u1,u2,u5 [B,A]; for SAC-EO also
           sE,spE [E,S], u3 (,u4) and I1 (,I2) index arrays (the array_split of the
           rng-shuffled arange, SAC_expert.py:301-309).
    hyper: gamma,tau,lr_q,lr_pi,lr_alpha,eps (epsilon coefficient), target_entropy,
           do_polyak (num_timesteps % target_update_int == 0, SAC_expert.py:475).
    Returns every intermediate the parity tests compare.
    """
    dt = state["actor"][0].dtype
    st = state
    s = torch.as_tensor(batch["s"]).to(dt)
    a = torch.as_tensor(batch["a"]).to(dt)
    sp = torch.as_tensor(batch["sp"]).to(dt)
    r = torch.as_tensor(batch["r"]).to(dt)
    # (1-done) is formed in float64 NumPy then cast by TF (SAC_expert.py:227) - exact for 0/1
    one_m_d = torch.as_tensor(1.0 - np.asarray(batch["d"], dtype=np.float64)).to(dt)
    B = s.shape[0]
    alpha = state["alpha"].detach().clone().to(dt)
    out: Dict = {}

    # ---- 1. TD target, no gradient (SAC_expert.py:211-229) ---------------------------
    with torch.no_grad():
        a1, nlp1 = head(cfg, st["actor"], sp, torch.as_tensor(batch["u1"]), st)
        qt1 = q_value(cfg, st["t1"], sp, a1, st)
        qt2 = q_value(cfg, st["t2"], sp, a1, st)
        min_next = torch.minimum(qt1, qt2)
        next_value = min_next + alpha * nlp1            # alpha RAW (may be negative)
        y = r + hyper["gamma"] * (one_m_d * next_value)
    out["y"] = y
    out["a1"], out["nlp1"] = a1, nlp1

    # ---- 2. critics (SAC_expert.py:232-250) ------------------------------------------
    new = {}
    for k in ("q1", "q2"):
        th = _req(st[k])
        pred = q_forward(cfg, th, s, a, st)                          # normalised space
        loss = (0.5 * ((pred - y[:, None]) ** 2).sum(-1)).mean()    # vs DEnormalised target
        grads = list(torch.autograd.grad(loss, th))
        out["L_" + k] = loss.detach()
        out["g_" + k] = grads
        ad = st["adam_" + k]
        th_new, m_new, v_new, t_new = keras_adam([p.detach() for p in th], grads, ad["m"],
                                                 ad["v"], ad["t"], hyper["lr_q"])
        new[k] = th_new
        new["adam_" + k] = dict(m=m_new, v=v_new, t=t_new)

    # ---- 3. actor (SAC_expert.py:262-338; plain: SAC.py:178-198) ---------------------
    th_pi = _req(st["actor"])
    a2, nlp2 = head(cfg, th_pi, s, torch.as_tensor(batch["u2"]), st)
    qa = q_forward(cfg, new["q1"], s, a2, st)          # UPDATED critics
    qb = q_forward(cfg, new["q2"], s, a2, st)
    min_q = torch.minimum(qa, qb)                      # reduce_min: ties split equally
    l_pi = (-alpha * nlp2[:, None] - min_q).mean()
    out["L_pi"] = l_pi.detach()
    out["a2"], out["nlp2"] = a2.detach(), nlp2.detach()
    if cfg.num_models == 0:
        p_loss = l_pi
        out["mse"] = torch.zeros((), dtype=dt)
    else:
        sE = torch.as_tensor(batch["sE"]).to(dt)
        spE = torch.as_tensor(batch["spE"]).to(dt)
        eps = hyper["eps"]
        if cfg.num_models == 1:                        # SAC_expert.py:271-297
            c, _ = head(cfg, th_pi, sE, torch.as_tensor(batch["u3"]), st)
            sp_pred = model_sample(cfg, st["m1"], sE, c, st)
            mse = (0.5 * ((spE - sp_pred) ** 2).sum(-1)).mean()
        else:                                          # SAC_expert.py:299-335
            I1 = torch.as_tensor(np.asarray(batch["I1"]), dtype=torch.long)
            I2 = torch.as_tensor(np.asarray(batch["I2"]), dtype=torch.long)
            c1, _ = head(cfg, th_pi, sE[I1], torch.as_tensor(batch["u3"]), st)
            c2, _ = head(cfg, th_pi, sE[I2], torch.as_tensor(batch["u4"]), st)
            p1 = model_sample(cfg, st["m1"], sE[I1], c1, st)
            p2 = model_sample(cfg, st["m2"], sE[I2], c2, st)
            dl = ((spE[I1] - p1) ** 2).sum(-1) + ((spE[I2] - p2) ** 2).sum(-1)
            mse = (0.5 * dl).mean()
        out["mse"] = mse.detach()
        p_loss = (1 - eps) * l_pi + eps * mse
    g_pi = list(torch.autograd.grad(p_loss, th_pi))
    out["p_loss"] = p_loss.detach()
    out["g_actor"] = g_pi
    ad = st["adam_actor"]
    th_new, m_new, v_new, t_new = keras_adam([p.detach() for p in th_pi], g_pi, ad["m"], ad["v"],
                                             ad["t"], hyper["lr_pi"])
    new["actor"] = th_new
    new["adam_actor"] = dict(m=m_new, v=v_new, t=t_new)

    # ---- 4. temperature (SAC_expert.py:341-348) --------------------------------------
    al = alpha.detach().clone().requires_grad_(True)
    with torch.no_grad():
        _, nlp3 = head(cfg, new["actor"], s, torch.as_tensor(batch["u5"]), st)   # UPDATED actor
    alpha_loss = -al * (-nlp3[:, None] + hyper["target_entropy"]).mean()
    (g_al,) = torch.autograd.grad(alpha_loss, al)
    out["alpha_loss"] = alpha_loss.detach()
    out["g_alpha"] = g_al
    out["nlp3"] = nlp3
    ad = st["adam_alpha"]
    (al_new,), (am,), (av,), at = keras_adam([al.detach()], [g_al], [ad["m"]], [ad["v"]], ad["t"],
                                             hyper["lr_alpha"])
    al_new = torch.clamp(al_new, min=1e-5)             # SAC_expert.py:348
    new["alpha"] = al_new
    new["adam_alpha"] = dict(m=am, v=av, t=at)

    # ---- 5. Polyak (SAC_expert.py:362-373, gated at :475) ----------------------------
    if hyper.get("do_polyak", True):
        new["t1"] = polyak(st["t1"], new["q1"], hyper["tau"])
        new["t2"] = polyak(st["t2"], new["q2"], hyper["tau"])
    else:
        new["t1"] = [t.clone() for t in st["t1"]]
        new["t2"] = [t.clone() for t in st["t2"]]
    out["new"] = new
    return out


# --------------------------------------------------------------------------------------
# behaviour cloning from expert observations  BC.py:309-363  (the epsilon = 1 slice of the SAC-EO actor step)
# --------------------------------------------------------------------------------------
def bc_update(cfg: NetCfg, state: Dict, batch: Dict, hyper: Dict) -> Dict:
    """``BC._update_actor`` (BC.py:309-363): ONLY the expert-observation term - counterfactual actions
    ``actor.sample(sE, deterministic=False)`` go through the frozen model(s), loss = mean(0.5*sum (s'E - pred)^2)
    (two models: both halves of a per-update shuffle summed row-wise, :329-354), one Keras-Adam step on the actor
    (``self.actor_optimizer``, BC.py:113, learning rate ``lr_pi``).  Critics, temperature and targets are untouched.
    ``batch``: sE, spE, u3 (and I1, I2, u4 for two models) exactly as in ``sac_eo_update``."""
    st = state
    dt = st["actor"][0].dtype
    th_pi = _req(st["actor"])
    sE = torch.as_tensor(batch["sE"]).to(dt)
    spE = torch.as_tensor(batch["spE"]).to(dt)
    if cfg.num_models == 1:
        c, _ = head(cfg, th_pi, sE, torch.as_tensor(batch["u3"]), st)
        mse = (0.5 * ((spE - model_sample(cfg, st["m1"], sE, c, st)) ** 2).sum(-1)).mean()
    elif cfg.num_models == 2:
        I1 = torch.as_tensor(np.asarray(batch["I1"]), dtype=torch.long)
        I2 = torch.as_tensor(np.asarray(batch["I2"]), dtype=torch.long)
        c1, _ = head(cfg, th_pi, sE[I1], torch.as_tensor(batch["u3"]), st)
        c2, _ = head(cfg, th_pi, sE[I2], torch.as_tensor(batch["u4"]), st)
        p1 = model_sample(cfg, st["m1"], sE[I1], c1, st)
        p2 = model_sample(cfg, st["m2"], sE[I2], c2, st)
        mse = (0.5 * (((spE[I1] - p1) ** 2).sum(-1) + ((spE[I2] - p2) ** 2).sum(-1))).mean()
    else:
        raise ValueError("BC needs one or two dynamics models")
    g_pi = list(torch.autograd.grad(mse, th_pi))
    ad = st["adam_actor"]
    th_new, m_new, v_new, t_new = keras_adam([p.detach() for p in th_pi], g_pi, ad["m"], ad["v"], ad["t"], hyper["lr_pi"])
    return dict(mse=mse.detach(), g_actor=g_pi, new=dict(actor=th_new, adam_actor=dict(m=m_new, v=v_new, t=t_new)))


# --------------------------------------------------------------------------------------
# dynamics-model fitting  mbrl_onpolicy_alg.py:301-319, SAC_expert.py:480-556
# --------------------------------------------------------------------------------------
def model_loss(cfg: NetCfg, theta_m, s_: Tensor, a_: Tensor, sp_: Tensor, r_: Tensor, st: Dict,
               reward_loss_coef: float = 1.0, delta_clip_loss: float = 0.0,
               reward_clip_loss: float = 0.0, logstd: Optional[Tensor] = None,
               scale_model_loss: bool = False) -> Tensor:
    """``MSEModel.get_loss`` (continuous_models.py:280-302) / ``GaussianModel.get_loss`` (:101-131) on
    ``BaseWorldModel._forward(s, a, clip=False)`` (base_world_model.py:65-87): the net predicts the
    NORMALISED state delta in its first S outputs and the normalised reward in the last one; targets
    are normalised with the model's ``delta_rms`` / ``r_rms`` and optionally clipped.
    MSE:      loss = mean_b(0.5*sum_j (dn-dp)^2 + coef*0.5*(rn-rp)^2)
    Gaussian: ``logstd`` [1,S] (trainable, NOT clipped): neglogp = 0.5*sum_j(((dn-dp)/exp(ls))^2 + 2 ls +
              log 2pi); loss = mean_b(scale*neglogp + coef*0.5*(rn-rp)^2), scale = stop_gradient(mean(
              exp(ls)^2)) with ``scale_model_loss`` else 1.
    ``separate_reward_nn`` (base_world_model.py:72-74): ``theta_m`` = the model net's six tensors followed by the reward
    net's six; the model net predicts the S delta columns, the reward net (same normalised input) the reward."""
    sa = torch.cat([normalize(s_, st["m_s_mean"], st["m_s_std"]),
                    normalize(a_, st["m_a_mean"], st["m_a_std"])], -1)
    if cfg.separate_reward_nn:
        delta_pred = mlp(theta_m[:6], sa.to(s_.dtype), cfg.model_acts)
        r_pred = mlp(theta_m[6:12], sa.to(s_.dtype), cfg.reward_acts or cfg.model_acts).squeeze(-1)
    else:
        pred = mlp(theta_m, sa.to(s_.dtype), cfg.model_acts)
        delta_pred, r_pred = pred[:, :-1], pred[:, -1]
    delta_norm = normalize(sp_ - s_, st["m_d_mean"], st["m_d_std"])
    if delta_clip_loss:
        delta_norm = torch.clamp(delta_norm, -delta_clip_loss, delta_clip_loss)
    if logstd is None:
        delta_loss = 0.5 * ((delta_norm - delta_pred) ** 2).sum(-1)
    else:
        vec = ((delta_norm - delta_pred) / torch.exp(logstd)) ** 2 + 2 * logstd + math.log(2 * math.pi)
        delta_loss = 0.5 * vec.sum(-1)
        if scale_model_loss:
            delta_loss = (torch.exp(logstd) ** 2).mean().detach() * delta_loss
    r_norm = normalize(r_, st.get("m_r_mean", 0.0), st.get("m_r_std", 1.0))
    if reward_clip_loss:
        r_norm = torch.clamp(r_norm, -reward_clip_loss, reward_clip_loss)
    r_loss = 0.5 * (r_norm - r_pred) ** 2
    return (delta_loss + reward_loss_coef * r_loss).mean()


def apply_model_grads(cfg: NetCfg, models: List[List[Tensor]], adam: Dict, batches: List[Dict], st: Dict,
                      fit: Dict) -> Dict:
    """``MBRLOnPolicyAlg._apply_model_grads`` (mbrl_onpolicy_alg.py:301-319): the losses of ALL models
    (each on its own minibatch) are summed under one tape, the gradient list is optionally clipped
    with ``tf.clip_by_global_norm(grads, model_max_grad_norm * num_models)`` and ONE Keras Adam
    (lr ``model_lr``, one shared step counter) applies it to every model's tensors.
    ``adam`` = dict(m=[per model lists], v=[...], t=int); ``batches[k]`` = dict(s,a,sp,r) tensors.
    ``GaussianModel``: pass each model's tensor list with the ``logstd`` [1,S] variable appended
    (``model_trainable = nn weights + [logstd]``, continuous_models.py:27) and ``fit["gaussian"] = True``.
    ``separate_reward_nn`` (MSE loss): each model's list = model net tensors + reward net tensors
    (``trainable = model_trainable + reward_trainable``, continuous_models.py:216-219)."""
    dt = models[0][0].dtype
    gauss = bool(fit.get("gaussian", False))
    live = [[w.detach().clone().requires_grad_(True) for w in th] for th in models]
    losses = []
    for k, th in enumerate(live):
        b = batches[k]
        losses.append(model_loss(cfg, th[:-1] if gauss else th, b["s"].to(dt), b["a"].to(dt), b["sp"].to(dt),
                                 b["r"].to(dt), st, fit.get("reward_loss_coef", 1.0), fit.get("delta_clip_loss", 0.0),
                                 fit.get("reward_clip_loss", 0.0), th[-1] if gauss else None,
                                 bool(fit.get("scale_model_loss", False))))
    flat_params = [w for th in live for w in th]
    grads = list(torch.autograd.grad(sum(losses), flat_params))
    gnorm = torch.sqrt(sum((g ** 2).sum() for g in grads))
    mgn = fit.get("model_max_grad_norm", 0.0)
    if mgn:                                   # tf.clip_by_global_norm: g * clip * min(1/norm, 1/clip)
        clip = torch.as_tensor(mgn * len(models), dtype=dt)
        scale = clip * torch.minimum(1.0 / gnorm, 1.0 / clip)
        grads = [g * scale for g in grads]
    flat_m = [x for ml in adam["m"] for x in ml]
    flat_v = [x for vl in adam["v"] for x in vl]
    new_theta, new_m, new_v, t = keras_adam([w.detach() for w in flat_params], grads, flat_m, flat_v,
                                            adam["t"], fit.get("model_lr", 1e-3))
    n = len(models[0])
    split = lambda xs: [xs[i * n:(i + 1) * n] for i in range(len(models))]
    return dict(losses=[l.detach() for l in losses], grads=split(grads), gnorm=gnorm.detach(),
                models=split(new_theta), m=split(new_m), v=split(new_v), t=t)


def model_fit_batches(n_train: int, num_models: int, model_batch_size: int, batch_shuffle: bool,
                      rng=np.random) -> List[np.ndarray]:
    """One epoch of minibatch indices the way ``_update_models`` draws them (SAC_expert.py:524-540):
    per-model independent ``np.random.shuffle`` when ``model_batch_shuffle`` else one shared shuffle
    tiled over the models; split every ``model_batch_size`` columns, dropping a ragged last batch.
    Returns a list of ``[num_models, model_batch_size]`` int arrays."""
    idx = np.arange(n_train)
    if batch_shuffle:
        idx = np.tile(idx, (num_models, 1))
        for row in idx:
            rng.shuffle(row)
    else:
        rng.shuffle(idx)
        idx = np.tile(idx, (num_models, 1))
    sections = np.arange(0, n_train, model_batch_size)[1:]
    batches = np.array_split(idx, sections, axis=1)
    if n_train % model_batch_size != 0:
        batches = batches[:-1]
    return batches


# --------------------------------------------------------------------------------------
# adaptive expert weight (host side, once per episode)  SAC_expert.py:375-460, 579-608
# --------------------------------------------------------------------------------------
def model_mse_on_expert(cfg: NetCfg, state: Dict, sE, aE, spE, u=None, use_expert_actions=False):
    """MSE bookkeeping at the end of ``_update_models`` (SAC_expert.py:579-608): mean over
    models of mean_i 0.5*sum_j (model_k.sample(sE, a) - s'E)^2 with a = expert actions or one
    shared stochastic actor draw."""
    dt = state["actor"][0].dtype
    sE = torch.as_tensor(sE).to(dt)
    spE = torch.as_tensor(spE).to(dt)
    with torch.no_grad():
        if use_expert_actions:
            act = torch.as_tensor(aE).to(dt)
        else:
            act, _ = head(cfg, state["actor"], sE, torch.as_tensor(u), state)
        vals = []
        for k in range(cfg.num_models):
            pred = model_sample(cfg, state["m%d" % (k + 1)], sE, act, state)
            vals.append((0.5 * ((pred - spE) ** 2).sum(-1)).mean())
    return torch.stack(vals).mean()


def adaptive_epsilon(epsilon: float, *, scale_by_true_mse=False, mse_cf=None, j_cur=0.0, j_exp=1.0,
                     min_mult=False, exp_mult=False, mult_coeff=1.0,
                     disc_mode: Optional[str] = None, disc: Optional[np.ndarray] = None) -> float:
    """``SAC_exp._expert_preprocess`` (SAC_expert.py:381-418)."""
    eps = epsilon
    if scale_by_true_mse:
        eps = 1.0 / (epsilon * float(mse_cf) + 1.0)
        if j_cur > 0:
            if min_mult:
                eps = eps * (-min(mult_coeff * (j_cur / j_exp) - 1.0, 0.0))
            if exp_mult:
                eps = eps * math.exp(-mult_coeff * j_cur / j_exp)
    elif disc_mode in ("max", "median", "total"):
        val = {"max": np.max, "median": np.median, "total": np.sum}[disc_mode](disc)
        eps = 1.0 / (epsilon * float(val) + 1.0)
    return eps


# --------------------------------------------------------------------------------------
# CG / Fisher-vector product   update_utils.py:4-24, trpo.py:179-187,200-227
# --------------------------------------------------------------------------------------
def gaussian_forward(cfg: NetCfg, theta: Sequence[Tensor], s_: Tensor, st: Dict):
    """``GaussianActor._forward`` (continuous_actors.py:74-100) - the parameterisation the
    TRPO KL uses (NOT ``evaluate``): softplus std + logstd_init, floor log(1e-3)."""
    dt = s_.dtype
    out = mlp(theta, normalize(s_, st["s_mean"], st["s_std"]).to(dt), cfg.actor_acts)
    if cfg.per_state_std:
        mean, o2 = torch.split(out, cfg.A, dim=-1)
        logstd = torch.log(torch.nn.functional.softplus(o2))
        logstd_init = math.log(cfg.std_mult) - math.log(math.log(2.0))   # :39-41
    else:
        mean = out
        logstd = theta[6] * torch.ones_like(mean)
        logstd_init = math.log(cfg.std_mult)                             # :43-44
    logstd = logstd + logstd_init
    logstd = torch.maximum(logstd, torch.as_tensor(math.log(1e-3), dtype=dt))
    return mean, logstd


def kl_forward(mean, logstd, mean_ref, logstd_ref):
    """``GaussianActor.kl(direction='forward')`` (continuous_actors.py:176-184)."""
    num = (mean - mean_ref) ** 2 + torch.exp(2 * logstd_ref)
    vec = num / torch.exp(2 * logstd) + 2 * logstd - 2 * logstd_ref - 1
    return 0.5 * vec.sum(-1)


def flat(ts: Sequence[Tensor]) -> Tensor:
    """``list_to_flat`` (nn_utils.py:177-182)."""
    return torch.cat([t.reshape(-1) for t in ts], -1)


def unflat(like: Sequence[Tensor], vec: Tensor) -> List[Tensor]:
    """``flat_to_list`` (nn_utils.py:162-175)."""
    res, o = [], 0
    for t in like:
        n = t.numel()
        res.append(vec[o:o + n].reshape(t.shape))
        o += n
    return res


def make_F(cfg: NetCfg, theta: Sequence[Tensor], s_all, st: Dict, damp: float, trust_sub: int = 1):
    """``TRPO._make_F`` (trpo.py:200-227): double back-prop Hessian-vector product of the
    mean forward KL to the current policy, plus ``trust_damp * x``."""
    dt = theta[0].dtype
    s_sub = torch.as_tensor(s_all).to(dt)[::trust_sub]
    with torch.no_grad():
        mean_ref, logstd_ref = gaussian_forward(cfg, theta, s_sub, st)

    def F(x):
        x = torch.as_tensor(x).to(dt)
        th = _req(theta)
        mean, logstd = gaussian_forward(cfg, th, s_sub, st)
        kl = kl_forward(mean, logstd, mean_ref, logstd_ref).mean()
        grads = torch.autograd.grad(kl, th, create_graph=True)
        gx = (flat(grads) * x).sum()
        res = torch.autograd.grad(gx, th, allow_unused=True)
        res = [r if r is not None else torch.zeros_like(p) for r, p in zip(res, th)]
        return (flat(res) + damp * x).detach()

    return F


def make_F_gn(cfg: NetCfg, theta: Sequence[Tensor], s_all, st: Dict, damp: float, trust_sub: int = 1):
    """Same operator in Gauss-Newton form  F x = (1/N) sum_s J^T diag(exp(-2 logstd), 2) J x
    + damp x   - the form the CUDA kernels implement (JVP then VJP)."""
    dt = theta[0].dtype
    s_sub = torch.as_tensor(s_all).to(dt)[::trust_sub]
    N = s_sub.shape[0]

    def F(x):
        x = torch.as_tensor(x).to(dt)
        th = _req(theta)
        tang = unflat(th, x)

        def fwd(*ps):
            m, ls = gaussian_forward(cfg, list(ps), s_sub, st)
            return torch.cat([m, ls], -1)

        outv, jx = torch.autograd.functional.jvp(fwd, tuple(th), tuple(tang), create_graph=False)
        A = cfg.A
        logstd = outv[:, A:]
        w = torch.cat([torch.exp(-2 * logstd) * jx[:, :A], 2.0 * jx[:, A:]], -1) / N
        out2 = fwd(*th)
        res = torch.autograd.grad(out2, th, grad_outputs=w.detach(), allow_unused=True)
        res = [r if r is not None else torch.zeros_like(p) for r, p in zip(res, th)]
        return (flat(res) + damp * x).detach()

    return F


def cg(f_Ax, b: Tensor, cg_iters: int = 20, residual_tol: float = 1e-10) -> Tensor:
    """``cg`` (update_utils.py:4-24), OpenAI-Baselines conjugate gradient, same order of
    operations, dtype of ``b`` throughout."""
    p = b.clone()
    r = b.clone()
    x = torch.zeros_like(b)
    rdotr = r.dot(r)
    for _ in range(cg_iters):
        z = f_Ax(p)
        v = rdotr / p.dot(z)
        x = x + v * p
        r = r - v * z
        newrdotr = r.dot(r)
        mu = newrdotr / rdotr
        p = r + mu * p
        rdotr = newrdotr
        if rdotr < residual_tol:
            break
    return x


def trpo_step(f_Ax, b: Tensor, delta: float, cg_iters: int = 20):
    """trpo.py:179-187: v = cg(F, b); vFv = v.F(v); eta = sqrt(2 delta / vFv)."""
    v = cg(f_Ax, b, cg_iters)
    vFv = v.dot(f_Ax(v))
    eta = torch.sqrt(2 * delta / vFv)
    return v, vFv, eta * v


def gaussian_neglogp(mean, logstd, a_):
    """``GaussianActor.neglogp`` (continuous_actors.py:137-143)."""
    vec = ((a_ - mean) / torch.exp(logstd)) ** 2 + 2 * logstd + math.log(2 * math.pi)
    return 0.5 * vec.sum(-1)


def gaussian_entropy(logstd):
    """``GaussianActor.entropy`` (continuous_actors.py:145-148)."""
    return 0.5 * (2 * logstd + math.log(2 * math.pi) + 1).sum(-1)


def trpo_normalise_adv(adv: np.ndarray, adv_center: bool = True, adv_scale: bool = True) -> np.ndarray:
    """trpo.py:41-48 (and again :243-249): NumPy mean / std (+1e-8) of the advantages."""
    adv = np.asarray(adv)
    mean, std = np.mean(adv), np.std(adv) + 1e-8
    if adv_center:
        adv = adv - mean
    if adv_scale:
        adv = adv / std
    return adv


def trpo_surrogate_grad(cfg: NetCfg, theta: Sequence[Tensor], s_all, a_all, adv_all, nlp_old, alpha: float,
                        ent_targ: float, st: Dict):
    """Surrogate tape of ``TRPO.update`` (trpo.py:52-63; identical in the expert branches :79-89, :121-131):
    ``pg_loss = mean(-ratio adv) - alpha (mean entropy - ent_targ)``; returns (neg_pg list, alpha_grad, pg_loss)."""
    dt = theta[0].dtype
    s_, a_ = torch.as_tensor(s_all).to(dt), torch.as_tensor(a_all).to(dt)
    adv, old = torch.as_tensor(adv_all).to(dt), torch.as_tensor(nlp_old).to(dt)
    th = _req(theta)
    al = torch.tensor(float(alpha), dtype=dt, requires_grad=True)
    mean, logstd = gaussian_forward(cfg, th, s_, st)
    ratio = torch.exp(old - gaussian_neglogp(mean, logstd, a_))
    pg_loss = (ratio * adv * -1).mean()
    ent_loss = gaussian_entropy(logstd).mean()
    pg_loss = pg_loss - al * (ent_loss - ent_targ)
    grads = torch.autograd.grad(pg_loss, th + [al], allow_unused=True)
    neg_pg = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads[:-1], th)]
    return [g.detach() for g in neg_pg], grads[-1].detach(), pg_loss.detach()


def trpo_eval(cfg: NetCfg, theta: Sequence[Tensor], s_all, a_all, adv_all, nlp_old, kl_info_ref, st: Dict) -> Dict:
    """Quantities ``TRPO._backtrack`` recomputes after every trial step (trpo.py:251-263, :283-288)."""
    dt = theta[0].dtype
    s_, a_ = torch.as_tensor(s_all).to(dt), torch.as_tensor(a_all).to(dt)
    adv, old = torch.as_tensor(adv_all).to(dt), torch.as_tensor(nlp_old).to(dt)
    with torch.no_grad():
        mean, logstd = gaussian_forward(cfg, theta, s_, st)
        nlp = gaussian_neglogp(mean, logstd, a_)
        ratio = torch.exp(old - nlp)
        res = {"nlp": nlp, "surr": (ratio * adv).mean(), "tv": 0.5 * (ratio - 1.).abs().mean(),
               "ent": gaussian_entropy(logstd).mean(), "kl_info": torch.stack((mean, logstd), -1)}
        if kl_info_ref is not None:
            ref = torch.as_tensor(kl_info_ref).to(dt)
            res["kl"] = kl_forward(mean, logstd, ref[..., 0], ref[..., 1]).mean()
    return res


def actor_increment(cfg: NetCfg, theta: Sequence[Tensor], step_flat: Tensor) -> List[Tensor]:
    """``set_weights(step, from_flat=True, increment=True)`` (continuous_actors.py:211-233): add, then floor the
    state-independent logstd variable at log(1e-3)."""
    new = [t + d for t, d in zip(theta, unflat(theta, step_flat.to(theta[0].dtype)))]
    if not cfg.per_state_std:
        new[-1] = torch.maximum(new[-1], torch.as_tensor(math.log(1e-3), dtype=new[-1].dtype))
    return new


def gaussian_sample(cfg: NetCfg, theta: Sequence[Tensor], s_: Tensor, u: Optional[Tensor], st: Dict) -> Tensor:
    """``GaussianActor.sample`` (continuous_actors.py:103-123): ``_forward`` parameterisation, ``a = mean +
    exp(logstd) u`` with u ~ np.random.normal (None: deterministic), NO squash and NO clipping."""
    mean, logstd = gaussian_forward(cfg, theta, s_, st)
    return mean if u is None else mean + torch.exp(logstd) * torch.as_tensor(u).to(mean.dtype)


def trpo_expert_blend(cfg: NetCfg, theta: Sequence[Tensor], neg_pg: Sequence[Tensor], batch: Dict, eps: float, st: Dict):
    """Two-model expert branch of ``TRPO.update`` (trpo.py:113-165): the expert rows are shuffled with the algorithm's
    own generator and split in two (``batch["I1"], batch["I2"]``, :115-117), each half goes through
    ``actor.sample`` (fresh noise ``u3``, ``u4``) and its own model, ``MSE = mean(0.5 (sum (s'E1 - p1)^2 + sum (s'E2 -
    p2)^2))`` (equal halves required), ``grad_final = (1 - eps) neg_pg + eps MSE_grads``; the temperature gradient is
    NOT blended (:165 is commented out).  Returns (grad_final, mse, norm_pg, norm_MSE) with the norms as the reference
    logs them: sums of per-tensor L2 norms (:160-163).
    The single-model branch (:77-111) is not restated: ``tape.gradient(MSE_loss, [..., self.alpha])`` yields None for
    the temperature and ``(1 - epsilon) * alpha_grad + epsilon * None`` (:111) raises TypeError in the reference."""
    dt = theta[0].dtype
    sE, spE = torch.as_tensor(batch["sE"]).to(dt), torch.as_tensor(batch["spE"]).to(dt)
    I1 = torch.as_tensor(np.asarray(batch["I1"]), dtype=torch.long)
    I2 = torch.as_tensor(np.asarray(batch["I2"]), dtype=torch.long)
    th = _req(theta)
    c1 = gaussian_sample(cfg, th, sE[I1], batch["u3"], st)
    c2 = gaussian_sample(cfg, th, sE[I2], batch["u4"], st)
    p1 = model_sample(cfg, st["m1"], sE[I1], c1, st)
    p2 = model_sample(cfg, st["m2"], sE[I2], c2, st)
    mse = (0.5 * (((spE[I1] - p1) ** 2).sum(-1) + ((spE[I2] - p2) ** 2).sum(-1))).mean()
    g = torch.autograd.grad(mse, th, allow_unused=True)
    g = [x if x is not None else torch.zeros_like(p) for x, p in zip(g, th)]
    final = [(1 - eps) * a + eps * b for a, b in zip(neg_pg, g)]
    norm_pg = sum(float(torch.linalg.norm(a)) for a in neg_pg)
    norm_mse = sum(float(torch.linalg.norm(b)) for b in g)
    return [f.detach() for f in final], mse.detach(), norm_pg, norm_mse


def ppo_expert_blend(cfg: NetCfg, theta: Sequence[Tensor], neg_pg: Sequence[Tensor], batch: Dict, eps: float, st: Dict):
    """Expert branch of ``PPO._apply_actor_grad`` (ppo.py:176-213, ``use_expert_actions = False``): ALL expert rows (in
    order, no shuffle) through ``tf_clip(actor.sample(s_expert))`` (continuous_actors.py:103-129, clip to the action
    limits) and ``models[0].sample``; ``MSE = mean(0.5 sum (s'E - pred)^2)``; the tape differentiates ``(1 - eps)
    pg_loss + eps MSE``, i.e. ``neg_pg_final = (1 - eps) neg_pg + eps MSE_grads``.  ``batch["u3"]`` holds the [E, A]
    noise of the one ``actor.sample`` call.  Returns (grad_final, mse)."""
    dt = theta[0].dtype
    sE, spE = torch.as_tensor(batch["sE"]).to(dt), torch.as_tensor(batch["spE"]).to(dt)
    th = _req(theta)
    lim = torch.as_tensor(np.asarray(st["act_limit"])).to(dt)
    a = gaussian_sample(cfg, th, sE, batch["u3"], st)
    a = torch.maximum(torch.minimum(a, lim), -lim)                         # tf.clip_by_value: gradient on the closed interval
    p = model_sample(cfg, st["m1"], sE, a, st)
    mse = (0.5 * ((spE - p) ** 2).sum(-1)).mean()
    g = torch.autograd.grad(mse, th, allow_unused=True)
    g = [x if x is not None else torch.zeros_like(q) for x, q in zip(g, th)]
    return [((1 - eps) * a_ + eps * b_).detach() for a_, b_ in zip(neg_pg, g)], mse.detach()


def backtrack(trial, eta_v, kl_maxfactor: float, delta: float):
    """Control flow of ``TRPO._backtrack`` (trpo.py:251-301).  ``trial(step) -> (theta, stats, improve)`` applies
    ``theta_k + step`` and evaluates it (``stats`` needs "kl", "tv").  Returns (theta, stats, improve, adj, step, tv_pre,
    kl_pre); ``adj == 0`` means ten shrinks did not help and the caller restores the old parameters (:292-301)."""
    th, e, improve = trial(eta_v)
    tv_pre, kl_pre = float(e["tv"]), float(e["kl"])
    adj = 1
    for _ in range(10):                                                            # :267-291
        if float(e["kl"]) > kl_maxfactor * delta:
            pass
        elif float(improve) < 0:
            pass
        else:
            break
        factor = np.sqrt(2)
        adj = adj / factor
        eta_v = eta_v / factor
        th, e, improve = trial(eta_v)
    else:
        adj = 0
    return th, e, improve, adj, eta_v, tv_pre, kl_pre


def trpo_update(cfg: NetCfg, theta: Sequence[Tensor], s_all, a_all, adv_all, st: Dict, *, delta: float = 0.01,
                cg_iters: int = 20, trust_sub: int = 1, trust_damp: float = 0.01, kl_maxfactor: float = 1.5,
                alpha: float = 0.0, ent_targ: float = 0.0, adv_center: bool = True, adv_scale: bool = True,
                expert: Optional[Dict] = None, eps: float = 0.0):
    """``TRPO.update`` (trpo.py:36-198): with ``expert`` (a dict like ``draw_batch``'s: sE, spE, I1, I2, u3, u4) the
    two-model gradient blend ``grad_final = (1 - eps) neg_pg + eps MSE_grads``; without it the epsilon = 0 slice
    (``grad_final = neg_pg``; the reference only defines ``grad_final`` inside its expert branches, :107-111,
    :154-158); followed by
    ``TRPO._backtrack`` (:229-317).  Returns (new theta, log dict, pg_vec, eta_v_flat)."""
    with torch.no_grad():
        m0, l0 = gaussian_forward(cfg, theta, torch.as_tensor(s_all).to(theta[0].dtype), st)
        nlp_old = gaussian_neglogp(m0, l0, torch.as_tensor(a_all).to(theta[0].dtype))
    adv = trpo_normalise_adv(adv_all, adv_center, adv_scale)
    neg_pg, _, _ = trpo_surrogate_grad(cfg, theta, s_all, a_all, adv, nlp_old, alpha, ent_targ, st)
    if expert is not None:                                                         # :113-165, two-model branch
        neg_pg, _, _, _ = trpo_expert_blend(cfg, theta, neg_pg, expert, eps, st)
    pg_vec = flat(neg_pg) * -1                                                     # :176
    if np.allclose(pg_vec.numpy(), 0) or delta == 0.0:                             # :179-180
        eta_v = torch.zeros_like(pg_vec)
    else:
        F = make_F(cfg, theta, s_all, st, trust_damp, trust_sub)                   # :182
        _, _, eta_v = trpo_step(F, pg_vec, delta, cg_iters)                        # :183-187
    # _backtrack (adv is normalised a second time there, :243-249)
    adv2 = trpo_normalise_adv(adv, adv_center, adv_scale)
    before = trpo_eval(cfg, theta, s_all, a_all, adv2, nlp_old, None, st)
    kl_ref, surr_before, ent = before["kl_info"], before["surr"], before["ent"]

    def trial(step):
        th = actor_increment(cfg, theta, step)
        e = trpo_eval(cfg, th, s_all, a_all, adv2, nlp_old, kl_ref, st)
        return th, e, e["surr"] - surr_before

    th, e, improve, adj, eta_v, tv_pre, kl_pre = backtrack(trial, eta_v, kl_maxfactor, delta)
    if adj == 0:                                                                   # :292-301 no policy update
        th = [t.clone() for t in theta]
        e = trpo_eval(cfg, th, s_all, a_all, adv2, nlp_old, kl_ref, st)
        improve = e["surr"] - surr_before
    log = {"ent": float(ent), "tv_pre": tv_pre, "kl_pre": kl_pre, "tv": float(e["tv"]), "kl": float(e["kl"]),
           "adj": adj, "improve": float(improve)}
    return th, log, pg_vec, eta_v


def ppo_actor_grad(cfg: NetCfg, theta: Sequence[Tensor], s_act, a_act, adv_act, nlp_old_act, alpha: float,
                   ent_targ: float, eps_ppo: float, max_grad_norm: Optional[float], st: Dict):
    """Gradient part of ``PPO._apply_actor_grad`` (ppo.py:132-147 tape, :226-231 global-norm clip), expert_reg = None.
    Returns (neg_pg list after clipping, alpha_grad, grad_norm_pre, grad_norm_post)."""
    dt = theta[0].dtype
    s_, a_ = torch.as_tensor(s_act).to(dt), torch.as_tensor(a_act).to(dt)
    adv, old = torch.as_tensor(adv_act).to(dt), torch.as_tensor(nlp_old_act).to(dt)
    th = _req(theta)
    al = torch.tensor(float(alpha), dtype=dt, requires_grad=True)
    mean, logstd = gaussian_forward(cfg, th, s_, st)
    ratio = torch.exp(old - gaussian_neglogp(mean, logstd, a_))
    ratio_clip = torch.clamp(ratio, 1. - eps_ppo, 1. + eps_ppo)
    surr, clip = ratio * adv * -1, ratio_clip * adv * -1
    # tf.maximum routes the gradient to its first argument where it is >= the second (torch.maximum would split ties)
    pg_loss = torch.where(surr >= clip, surr, clip).mean()
    pg_loss = pg_loss - al * (gaussian_entropy(logstd).mean() - ent_targ)
    grads = torch.autograd.grad(pg_loss, th + [al], allow_unused=True)
    neg_pg = [(g if g is not None else torch.zeros_like(p)).detach() for g, p in zip(grads[:-1], th)]
    norm_pre = torch.sqrt(sum((g ** 2).sum() for g in neg_pg))
    if max_grad_norm is not None:                                       # tf.clip_by_global_norm
        neg_pg = [g * (max_grad_norm / torch.maximum(norm_pre, torch.as_tensor(float(max_grad_norm), dtype=dt))) for g in neg_pg]
    norm_post = torch.sqrt(sum((g ** 2).sum() for g in neg_pg))
    return neg_pg, grads[-1].detach(), float(norm_pre), float(norm_post)


def ppo_update(cfg: NetCfg, theta: Sequence[Tensor], adam: Dict, s_all, a_all, adv_all, st: Dict, *, actor_lr: float = 3e-4,
               actor_update_it: int = 2, actor_nminibatch: int = 4, eps_ppo: float = 0.2,
               max_grad_norm: Optional[float] = 0.5, alpha: float = 0.0, ent_targ: float = 0.0,
               adv_center: bool = True, adv_scale: bool = True, np_rng=np.random):
    """``PPO.update`` (ppo.py:41-119) with expert_reg = None and ent_reg off: epochs x shuffled minibatches
    (``np.random.shuffle`` of the global RNG, ragged tail dropped), per-minibatch advantage normalisation, the clipped
    surrogate step and Keras-Adam on the actor.  ``adam`` = {"m": [...], "v": [...], "t": int} is updated in place.
    Returns (theta, log)."""
    dt = theta[0].dtype
    s_all, a_all, adv_all = np.asarray(s_all), np.asarray(a_all), np.asarray(adv_all)
    with torch.no_grad():
        m0, l0 = gaussian_forward(cfg, theta, torch.as_tensor(s_all).to(dt), st)
        nlp_old = gaussian_neglogp(m0, l0, torch.as_tensor(a_all).to(dt)).numpy()
        kl_ref = torch.stack((m0, l0), -1)
        ent = float(gaussian_entropy(l0).mean())
    n_samples = s_all.shape[0]
    n_batch = int(n_samples / actor_nminibatch)
    pre_all = post_all = 0.0
    theta = [t.clone() for t in theta]
    for _ in range(actor_update_it):
        idx = np.arange(n_samples)
        np_rng.shuffle(idx)
        sections = np.arange(0, n_samples, n_batch)[1:]
        batches = np.array_split(idx, sections)
        if n_samples % n_batch != 0:
            batches = batches[:-1]
        for b in batches:
            adv_b = trpo_normalise_adv(adv_all[b], adv_center, adv_scale)
            g, _, pre, post = ppo_actor_grad(cfg, theta, s_all[b], a_all[b], adv_b, nlp_old[b], alpha, ent_targ, eps_ppo,
                                             max_grad_norm, st)
            pre_all += pre; post_all += post
            theta, adam["m"], adam["v"], adam["t"] = keras_adam(theta, g, adam["m"], adam["v"], adam["t"], actor_lr)
    e = trpo_eval(cfg, theta, s_all, a_all, np.zeros(n_samples), nlp_old, kl_ref, st)
    with torch.no_grad():
        ratio_diff = (torch.exp(torch.as_tensor(nlp_old).to(dt) - e["nlp"]) - 1.).abs()
    nb = actor_update_it * len(batches)
    log = {"ent": ent, "tv": float(e["tv"]), "kl": float(e["kl"]), "outside_clip": float((ratio_diff > eps_ppo).double().mean()),
           "actor_grad_norm_pre": pre_all / nb, "actor_grad_norm": post_all / nb}
    return theta, log


# --------------------------------------------------------------------------------------
# synthetic problem builders (shared by tests, smoke and the CPU baseline)
# --------------------------------------------------------------------------------------
def orthogonal(rng: np.random.Generator, shape, gain: float) -> np.ndarray:
    """Keras ``Orthogonal`` initialiser semantics (QR of a Gaussian, sign-fixed), used by
    ``create_initializer`` (nn_utils.py:24-46)."""
    rows, cols = shape
    a = rng.standard_normal((max(rows, cols), min(rows, cols)))
    q, r_ = np.linalg.qr(a)
    q = q * np.sign(np.diag(r_))
    if rows < cols:
        q = q.T
    return (gain * q[:rows, :cols]).astype(np.float32)


def init_net(rng, n_in, hidden, n_out, gain) -> List[np.ndarray]:
    """``create_nn`` (nn_utils.py:86-138): orthogonal(sqrt 2) hidden kernels, orthogonal(gain)
    final kernel, zero biases."""
    dims = [n_in, hidden[0], hidden[1], n_out]
    ws = []
    for l in range(3):
        g = math.sqrt(2.0) if l < 2 else gain
        ws.append(orthogonal(rng, (dims[l], dims[l + 1]), g))
        ws.append(np.zeros(dims[l + 1], np.float32))
    return ws


def make_problem(cfg: NetCfg, B: int, E: int, N: int, seed: int, identity_norm: bool = False,
                 perturb: float = 0.0) -> Tuple[Dict, Dict, Dict, Dict]:
    """Synthetic MuJoCo-shaped problem: returns (state_np, replay, expert, hyper)
    as float32 NumPy.  ``perturb`` adds noise to biases / Adam slots so parity tests do not only
    exercise the all-zero start."""
    rng = np.random.default_rng(seed)
    S, A = cfg.S, cfg.A
    st: Dict = {}
    st["actor"] = init_net(rng, S, cfg.actor_hidden, cfg.Ao, 0.01)
    if not cfg.per_state_std:
        st["actor"].append(np.zeros((1, A), np.float32))
    st["q1"] = init_net(rng, S + A, cfg.critic_hidden, 1, 1.0)
    st["q2"] = init_net(rng, S + A, cfg.critic_hidden, 1, 1.0)
    st["t1"] = [w.copy() for w in st["q1"]]
    st["t2"] = [w.copy() for w in st["q2"]]
    st["m1"] = init_net(rng, S + A, cfg.model_hidden, cfg.model_out, 0.01)
    st["m2"] = init_net(rng, S + A, cfg.model_hidden, cfg.model_out, 0.01)
    st["alpha"] = np.float32(math.log(0.1))
    for k in ("q1", "q2", "actor"):
        st["adam_" + k] = dict(m=[np.zeros_like(w) for w in st[k]],
                               v=[np.zeros_like(w) for w in st[k]], t=0)
    st["adam_alpha"] = dict(m=np.float32(0), v=np.float32(0), t=0)
    if perturb:
        for k in ("actor", "q1", "q2", "t1", "t2", "m1", "m2"):
            st[k] = [w + (perturb * rng.standard_normal(w.shape)).astype(np.float32) for w in st[k]]
        for k in ("q1", "q2", "actor"):
            ad = st["adam_" + k]
            ad["m"] = [(1e-3 * rng.standard_normal(w.shape)).astype(np.float32) for w in ad["m"]]
            ad["v"] = [(1e-5 * rng.random(w.shape)).astype(np.float32) for w in ad["v"]]
            ad["t"] = 7
        st["adam_alpha"] = dict(m=np.float32(0.01), v=np.float32(1e-3), t=7)
        st["alpha"] = np.float32(0.2)
    if identity_norm:
        z = lambda n: np.zeros(n, np.float32)
        o = lambda n: np.ones(n, np.float32)
        st.update(s_mean=z(S), s_std=o(S), a_mean=z(A), a_std=o(A), ret_std=np.float32(1.0),
                  m_s_mean=z(S), m_s_std=o(S), m_a_mean=z(A), m_a_std=o(A), m_d_mean=z(S), m_d_std=o(S))
    else:
        nm = lambda n: rng.standard_normal(n).astype(np.float32) * 0.3
        sd = lambda n: rng.uniform(0.5, 2.0, n).astype(np.float32)
        st.update(s_mean=nm(S), s_std=sd(S), a_mean=nm(A) * 0.3, a_std=sd(A),
                  ret_std=np.float32(rng.uniform(0.5, 2.0)),
                  m_s_mean=nm(S), m_s_std=sd(S), m_a_mean=nm(A) * 0.3, m_a_std=sd(A),
                  m_d_mean=nm(S) * 0.1, m_d_std=sd(S) * 0.2)
    st["act_limit"] = np.ones(A, np.float32)
    replay = dict(s=rng.standard_normal((N, S)).astype(np.float32),
                  a=rng.uniform(-1, 1, (N, A)).astype(np.float32),
                  sp=rng.standard_normal((N, S)).astype(np.float32),
                  r=rng.standard_normal(N).astype(np.float32),
                  d=(rng.random(N) < 0.01).astype(np.float64))
    expert = dict(sE=rng.standard_normal((E, S)).astype(np.float32),
                  aE=rng.uniform(-1, 1, (E, A)).astype(np.float32),
                  spE=rng.standard_normal((E, S)).astype(np.float32))
    hyper = dict(gamma=0.995, tau=5e-3, lr_q=3e-4, lr_pi=1e-4, lr_alpha=1e-4, eps=1e-3,
                 target_entropy=float(-A), do_polyak=True)
    return st, replay, expert, hyper


def draw_batch(cfg: NetCfg, replay: Dict, expert: Dict, B: int, seed: int) -> Dict:
    """Draws idx / noise / expert split the way the reference consumes its RNGs
    (SURVEY.md App. A): idx, u1, [perm], u2, u3, u4, u5."""
    rng = np.random.default_rng(seed)
    N = replay["s"].shape[0]
    A = cfg.A
    idx = rng.integers(0, N, size=B).astype(np.int64)
    s, a, sp, r, d = gather(replay, idx)
    batch = dict(idx=idx, s=s, a=a, sp=sp, r=r, d=d)
    batch["u1"] = rng.standard_normal((B, A)).astype(np.float32)
    batch["u2"] = rng.standard_normal((B, A)).astype(np.float32)
    if cfg.num_models > 0:
        E = expert["sE"].shape[0]
        perm = np.arange(E)
        rng.shuffle(perm)
        if cfg.num_models == 1:
            batch["I1"] = np.arange(E)
            batch["u3"] = rng.standard_normal((E, A)).astype(np.float32)
        else:
            sec = np.array_split(perm, 2)
            batch["I1"], batch["I2"] = sec[0], sec[1]
            batch["u3"] = rng.standard_normal((len(sec[0]), A)).astype(np.float32)
            batch["u4"] = rng.standard_normal((len(sec[1]), A)).astype(np.float32)
        batch["sE"], batch["spE"] = expert["sE"], expert["spE"]
    batch["u5"] = rng.standard_normal((B, A)).astype(np.float32)
    return batch


def to_torch_state(st_np: Dict, dtype=torch.float32) -> Dict:
    """NumPy problem -> torch tensors of ``dtype`` (float32 = the reference arithmetic,
    float64 = the twin used to separate kernel error from fp32 noise)."""
    T = lambda x: torch.as_tensor(np.asarray(x)).to(dtype)
    out: Dict = {}
    for k, v in st_np.items():
        if k.startswith("adam_"):
            if isinstance(v["m"], list):
                out[k] = dict(m=[T(x) for x in v["m"]], v=[T(x) for x in v["v"]], t=int(v["t"]))
            else:
                out[k] = dict(m=T(v["m"]), v=T(v["v"]), t=int(v["t"]))
        elif isinstance(v, list):
            out[k] = [T(x) for x in v]
        else:
            out[k] = T(v)
    return out
