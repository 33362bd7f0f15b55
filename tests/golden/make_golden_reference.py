"""Golden vectors produced by the REFERENCE'S OWN CODE.  Run in the build container only:

    python tests/golden/make_golden_reference.py

The reference's unmodified Python files are imported from /root/reference and executed over ``oracle/tfemu`` (an eager
torch-CPU executor for the TensorFlow / gym symbols they import - TensorFlow itself is not installable here; what that
does and does not pin is in oracle/tfemu/README.md).  For every case below the script

  1. builds the reference's ``SquashedGaussianActor`` / ``QCritic`` x4 / ``MSEModel`` x{0,1,2} / ``SAC_exp`` (or ``SAC``)
     through the reference's own parser defaults and factories (train.py:60-105 without ``init_env``),
  2. loads an oracle problem (weights, Adam slots, normaliser statistics, replay and expert rows) into them,
  3. runs ``alg._update(num_timesteps, expert_reg)`` K times with the global NumPy RNG seeded, RECORDING every
     ``np.random.randint`` / ``np.random.normal`` call and every ``alg.rng.shuffle`` result in call order,
  4. stores inputs (small cases) or the generating seed + an input checksum (benchmark-shaped case), the recorded draws,
     and the reference's outputs per step: TD target, the gradients handed to each optimiser, logged losses, every
     network / target / alpha value, and the final Adam slots.

tests/test_reference_pin.py replays the recorded draws through oracle/sac_eo_oracle.py (CPU tier) and through
libsaceo (GPU tier) and compares.  /root/reference does not exist on the GPU box: only the .npz files travel.
"""
import hashlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "tfemu"))
sys.path.insert(0, REF)

from oracle.sac_eo_oracle import NetCfg, make_problem  # noqa: E402


class _Env:
    """observation_space / action_space only; never stepped."""

    def __init__(self, S, A):
        import gym
        self.observation_space = gym.spaces.Box(-np.inf, np.inf, (S,), np.float32)
        self.action_space = gym.spaces.Box(-1.0, 1.0, (A,), np.float32)

    def seed(self, s):
        pass


class _Recorder:
    """Wraps the global NumPy RNG entry points the update path uses; keeps (kind, value) in call order."""

    def __init__(self):
        self.calls = []
        self._randint, self._normal, self._shuffle = np.random.randint, np.random.normal, np.random.shuffle

    def __enter__(self):
        def randint(*a, **k):
            v = self._randint(*a, **k)
            self.calls.append(("randint", np.array(v)))
            return v

        def normal(*a, **k):
            v = self._normal(*a, **k)
            self.calls.append(("normal", np.array(v)))
            return v
        def shuffle(x):
            self._shuffle(x)
            self.calls.append(("shuffle", np.array(x)))
        np.random.randint, np.random.normal, np.random.shuffle = randint, normal, shuffle
        return self

    def __exit__(self, *exc):
        np.random.randint, np.random.normal, np.random.shuffle = self._randint, self._normal, self._shuffle
        return False


def build(cfg: NetCfg, st, replay, expert, hyper, B, target_update_int, alg_type=None, gaussian_logstd=None, setup_kw=None):
    from sac_eo.actors import init_actor
    from sac_eo.algs import init_alg
    from sac_eo.common.train_parser import create_train_parser
    from sac_eo.common.train_utils import gather_inputs
    from sac_eo.critics import init_critics
    from sac_eo.models import init_world_models
    alg_type = alg_type or ("sac_imit" if cfg.num_models > 0 else "sac")
    inputs = gather_inputs(create_train_parser().parse_args(["--alg_type", alg_type]))
    ak, ck, mk, msk, al = (inputs[k] for k in ("actor_kwargs", "critic_kwargs", "model_kwargs", "model_setup_kwargs",
                                               "alg_kwargs"))
    ak.update(actor_layers=list(cfg.actor_hidden), actor_activations=list(cfg.actor_acts), actor_weights=None,
              actor_per_state_std=cfg.per_state_std, actor_squash=True, actor_std_mult=cfg.std_mult)
    ck.update(critic_layers=list(cfg.critic_hidden), critic_activations=list(cfg.critic_acts), critic_weights=None)
    mk.update(model_layers=list(cfg.model_hidden), model_activations=list(cfg.model_acts), model_weights=None,
              reward_weights=None, num_models=max(cfg.num_models, 1), gaussian_model=gaussian_logstd is not None)
    msk.update(separate_reward_nn=cfg.separate_reward_nn, delta_clip_pred=cfg.delta_clip_pred or None)
    msk.update(setup_kw or {})
    al.update(init_rms_stats=None, alg_seed=0, gamma=hyper["gamma"], soft_tau=hyper["tau"], q_crit_lr=hyper["lr_q"],
              mbpo_actor_lr=hyper["lr_pi"], mbpo_alpha_lr=hyper["lr_alpha"], sac_batch_size=B,
              target_update_int=target_update_int, only_model_normalizer=True, epsilon=hyper["eps"],
              expert_buffer_size=len(expert["sE"]), env_buffer_size=len(replay["r"]))
    env = _Env(cfg.S, cfg.A)
    actor = init_actor(env, **ak)
    expert_actor = init_actor(env, **ak)
    critics, q_targets, q_critics = init_critics(env, **ck)
    models = init_world_models(env, **mk, model_setup_kwargs=msk)
    alg = init_alg(0, env, env, env, actor, critics, q_targets, q_critics, models, al, inputs["mf_update_kwargs"],
                   expert_actor, None)
    # ---- the oracle problem: weights, Adam slots, alpha ---------------------------------------------------------------
    actor.set_weights([np.asarray(w) for w in st["actor"]])
    for net, key in zip(q_critics, ("q1", "q2")):
        net.set_weights([np.asarray(w) for w in st[key]])
    for net, key in zip(q_targets, ("t1", "t2")):
        net.set_weights([np.asarray(w) for w in st[key]])
    for k_, (net, key) in enumerate(zip(models, ("m1", "m2"))):
        extra = [] if gaussian_logstd is None else [np.asarray(gaussian_logstd[k_], np.float32).reshape(1, -1)]
        net.set_weights([np.asarray(w) for w in st[key]] + extra)
    alg.alpha.assign(float(st["alpha"]))
    import torch
    for opt, variables, key in ((alg.q_critic1_optimizer, q_critics[0].trainable, "adam_q1"),
                                (alg.q_critic2_optimizer, q_critics[1].trainable, "adam_q2"),
                                (alg.actor_optimizer, actor.trainable, "adam_actor")):
        opt.iterations = int(st[key]["t"])
        for v, m_, v_ in zip(variables, st[key]["m"], st[key]["v"]):
            opt.slots[id(v)] = (torch.tensor(np.asarray(m_)).reshape(tuple(v.shape)).clone(),
                                torch.tensor(np.asarray(v_)).reshape(tuple(v.shape)).clone())
    alg.alpha_optimizer.iterations = int(st["adam_alpha"]["t"])
    alg.alpha_optimizer.slots[id(alg.alpha)] = (torch.tensor(np.float32(st["adam_alpha"]["m"])),
                                               torch.tensor(np.float32(st["adam_alpha"]["v"])))
    # ---- normaliser statistics: the running normalisers' attributes, then shared the reference's way -----------------
    nz = alg.normalizer
    nz.s_rms.mean, nz.s_rms.std = st["s_mean"].copy(), st["s_std"].copy()
    nz.a_rms.mean, nz.a_rms.std = st["a_mean"].copy(), st["a_std"].copy()
    nz.ret_rms.std = np.float32(st["ret_std"])
    mz = alg.model_normalizer
    mz.s_rms.mean, mz.s_rms.std = st["m_s_mean"].copy(), st["m_s_std"].copy()
    mz.a_rms.mean, mz.a_rms.std = st["m_a_mean"].copy(), st["m_a_std"].copy()
    mz.delta_rms.mean, mz.delta_rms.std = st["m_d_mean"].copy(), st["m_d_std"].copy()
    alg._set_rms()
    alg.env_data.add(replay["s"], replay["a"], replay["r"], replay["sp"], replay["d"])
    assert alg.env_data.current_size == len(replay["r"])
    expert_reg = (expert["sE"], expert["aE"], expert["spE"], hyper["eps"], False)
    return alg, expert_reg


def flat(ws):
    return np.concatenate([np.asarray(w, np.float32).ravel() for w in ws])


def run_case(name, cfg: NetCfg, B, E, N, seed, K, eps, target_update_int, store_inputs, proj_dim=0):
    st, replay, expert, hyper = make_problem(cfg, B, E, N, seed=seed, perturb=0.05)
    hyper["eps"] = eps
    alg, expert_reg = build(cfg, st, replay, expert, hyper, B, target_update_int)
    out = dict(meta=np.array([cfg.S, cfg.A, B, E, N, seed, K, target_update_int, cfg.num_models], np.int64),
               eps=np.float64(eps))
    h = hashlib.sha256()
    for k in ("actor", "q1", "q2", "t1", "t2", "m1", "m2"):
        h.update(flat(st[k]).tobytes())
    for k in ("s", "a", "sp", "r", "d"):
        h.update(np.ascontiguousarray(replay[k]).tobytes())
    out["input_sha256"] = np.frombuffer(h.digest(), np.uint8)
    if store_inputs:
        for k in ("actor", "q1", "q2", "t1", "t2", "m1", "m2"):
            for i, w in enumerate(st[k]):
                out[f"in_{k}_{i}"] = np.asarray(w, np.float32)
        for k in ("q1", "q2", "actor"):
            for i, (m_, v_) in enumerate(zip(st["adam_" + k]["m"], st["adam_" + k]["v"])):
                out[f"in_adam_{k}_m_{i}"], out[f"in_adam_{k}_v_{i}"] = m_, v_
        for k in ("s_mean", "s_std", "a_mean", "a_std", "ret_std", "m_s_mean", "m_s_std", "m_a_mean", "m_a_std",
                  "m_d_mean", "m_d_std", "alpha"):
            out["in_" + k] = np.asarray(st[k])
        for k, v in replay.items():
            out["in_replay_" + k] = v
        for k, v in expert.items():
            out["in_expert_" + k] = v

    def snap(step):
        nets = dict(actor=alg.actor.get_weights(), q1=alg.q_critics[0].get_weights(), q2=alg.q_critics[1].get_weights(),
                    t1=alg.q_targets[0].get_weights(), t2=alg.q_targets[1].get_weights())
        for k, ws in nets.items():
            f = flat(ws)
            if proj_dim:        # benchmark-sized nets: fixed random projections + norm instead of 1.4 MB of weights
                R = np.random.default_rng(12345).standard_normal((proj_dim, f.size))
                out[f"step{step}_proj_{k}"] = R @ (f.astype(np.float64) - flat(st[k]).astype(np.float64))
                out[f"step{step}_dnorm_{k}"] = np.float64(np.linalg.norm(f.astype(np.float64) - flat(st[k])))
            else:
                out[f"step{step}_theta_{k}"] = f
        out[f"step{step}_alpha"] = np.float32(alg.alpha.numpy())

    ys = []
    orig_target = alg._get_Q_target

    def spy_target(sp, r, done):
        y = orig_target(sp, r, done)
        ys.append(np.array(y.numpy()))
        return y
    alg._get_Q_target = spy_target
    shuffles = []
    orig_shuffle = alg.rng.shuffle

    class _Rng:             # alg.rng.shuffle(idx) in place (SAC_expert.py:301-303): keep the permutation it produced
        def shuffle(self_, x):
            orig_shuffle(x)
            shuffles.append(np.array(x))

        def __getattr__(self_, k):
            return getattr(alg.rng.__class__, k).__get__(alg.rng)
    real_rng, alg.rng = alg.rng, _Rng()
    np.random.seed(1000 + seed)
    for step in range(K):
        with _Recorder() as rec:
            if cfg.num_models > 0:
                alg._update(step, expert_reg)
            else:
                alg._update(step)
        kinds = [k for k, _ in rec.calls]
        want = ["randint", "normal", "normal"] + ["normal"] * cfg.num_models + ["normal"]
        assert kinds == want, (kinds, want)                  # consumption order of SURVEY.md App. A
        vals = [v for _, v in rec.calls]
        out[f"step{step}_idx"] = vals[0].astype(np.int64)
        for j, v in enumerate(vals[1:]):
            out[f"step{step}_normal{j}"] = v                 # float64, as drawn
        if cfg.num_models == 2:
            out[f"step{step}_perm"] = shuffles[-1].astype(np.int64)
        out[f"step{step}_y"] = ys[-1]
        g1, g2, ga = (alg.q_critic1_optimizer.last_grads, alg.q_critic2_optimizer.last_grads,
                      alg.actor_optimizer.last_grads)
        if proj_dim:
            for k, g in (("q1", g1), ("q2", g2), ("actor", ga)):
                f = flat(g).astype(np.float64)
                R = np.random.default_rng(54321).standard_normal((proj_dim, f.size))
                out[f"step{step}_gproj_{k}"], out[f"step{step}_gnorm_{k}"] = R @ f, np.float64(np.linalg.norm(f))
        else:
            out[f"step{step}_g_q1"], out[f"step{step}_g_q2"], out[f"step{step}_g_actor"] = flat(g1), flat(g2), flat(ga)
        out[f"step{step}_g_alpha"] = np.float32(alg.alpha_optimizer.last_grads[0])
        if cfg.num_models > 0:       # SAC.py keeps its losses local (the log_train call is commented out, SAC.py:217)
            out[f"step{step}_p_loss"] = np.float32(alg.logger.train_dict["p_loss"][-1])
            out[f"step{step}_alpha_loss"] = np.float32(alg.logger.train_dict["alpha_loss"][-1])
        snap(step)
    alg.rng = real_rng
    if not proj_dim:
        for key, opt, variables in (("q1", alg.q_critic1_optimizer, alg.q_critics[0].trainable),
                                    ("actor", alg.actor_optimizer, alg.actor.trainable)):
            ms, vs = zip(*(opt.get_slot_arrays(v) for v in variables))
            out[f"final_adam_{key}_m"], out[f"final_adam_{key}_v"] = flat(ms), flat(vs)
    path = os.path.join(OUT, f"ref_{name}.npz")
    buf = io.BytesIO()
    np.savez_compressed(buf, **out)
    open(path, "wb").write(buf.getvalue())
    print(f"{name}: {len(buf.getvalue()) / 1024:.0f} KiB, alpha per step {[float(out[f'step{i}_alpha']) for i in range(K)]}")


def run_trpo_case(name, cfg: NetCfg, N, E, seed, eps, delta, cg_it, kl_maxfactor, trust_damp):
    """``TRPO.update(rollout_data, expert_reg)`` (trpo.py:36-198: surrogate + entropy gradient, two-model expert blend,
    ``_make_F``, ``cg``, step length, ``_backtrack``) on a ``GaussianActor``, plus one stand-alone Fisher-vector product."""
    import torch
    from sac_eo.actors import init_actor
    from sac_eo.algs.model_free import trpo as trpo_mod
    from sac_eo.common.normalizer import RunningNormalizers
    from sac_eo.common.train_parser import create_train_parser
    from sac_eo.common.train_utils import gather_inputs
    from sac_eo.models import init_world_models
    from oracle.sac_eo_oracle import gaussian_forward, to_torch_state
    st, replay, expert, hyper = make_problem(cfg, 8, E, max(N, 300), seed=seed, perturb=0.2)
    inputs = gather_inputs(create_train_parser().parse_args(["--alg_type", "mbrl"]))
    ak, mk, msk, uk = (inputs[k] for k in ("actor_kwargs", "model_kwargs", "model_setup_kwargs", "mf_update_kwargs"))
    ak.update(actor_layers=list(cfg.actor_hidden), actor_activations=list(cfg.actor_acts), actor_weights=None,
              actor_per_state_std=cfg.per_state_std, actor_squash=False, actor_std_mult=cfg.std_mult)
    mk.update(model_layers=list(cfg.model_hidden), model_activations=list(cfg.model_acts), model_weights=None,
              reward_weights=None, num_models=2, gaussian_model=False)
    msk.update(separate_reward_nn=False, delta_clip_pred=None)
    uk.update(delta_trpo=delta, cg_it=cg_it, trust_sub=1, trust_damp=trust_damp, kl_maxfactor=kl_maxfactor,
              ent_reg=False, ent_targ=-cfg.A, adv_center=True, adv_scale=True)
    env = _Env(cfg.S, cfg.A)
    actor = init_actor(env, **ak)
    models = init_world_models(env, **mk, model_setup_kwargs=msk)
    actor.set_weights([np.asarray(w) for w in st["actor"]])
    for net, key in zip(models, ("m1", "m2")):
        net.set_weights([np.asarray(w) for w in st[key]])
    nz, mz = RunningNormalizers(cfg.S, cfg.A, 0.99), RunningNormalizers(cfg.S, cfg.A, 0.99)
    nz.s_rms.mean, nz.s_rms.std = st["s_mean"].copy(), st["s_std"].copy()
    mz.s_rms.mean, mz.s_rms.std = st["m_s_mean"].copy(), st["m_s_std"].copy()
    mz.a_rms.mean, mz.a_rms.std = st["m_a_mean"].copy(), st["m_a_std"].copy()
    mz.delta_rms.mean, mz.delta_rms.std = st["m_d_mean"].copy(), st["m_d_std"].copy()
    actor.set_rms(nz)
    for m_ in models:
        m_.set_rms(mz)
    algo = trpo_mod.TRPO(actor, uk)
    # rollout: the policy's own actions (ratios near 1), random advantages
    rng = np.random.default_rng(seed + 500)
    s_all = replay["s"][:N]
    th64 = to_torch_state(st, torch.float64)
    with torch.no_grad():
        mean, ls = gaussian_forward(cfg, th64["actor"], torch.as_tensor(s_all, dtype=torch.float64), th64)
    a_all = (mean + torch.exp(ls) * torch.from_numpy(rng.standard_normal((N, cfg.A)))).numpy().astype(np.float32)
    adv_all = (rng.standard_normal(N) * 1.5 + 0.3).astype(np.float32)
    out = dict(meta=np.array([cfg.S, cfg.A, N, E, seed, cg_it, int(cfg.per_state_std)], np.int64),
               hyper=np.array([eps, delta, kl_maxfactor, trust_damp, cfg.std_mult], np.float64),
               s_all=s_all, a_all=a_all, adv_all=adv_all)
    for k in ("actor", "m1", "m2"):
        for i, w in enumerate(st[k]):
            out[f"in_{k}_{i}"] = np.asarray(w, np.float32)
    for k in ("s_mean", "s_std", "m_s_mean", "m_s_std", "m_a_mean", "m_a_std", "m_d_mean", "m_d_std"):
        out["in_" + k] = np.asarray(st[k])
    for k, v in expert.items():
        out["in_expert_" + k] = v
    # ---- one Fisher-vector product of the untouched policy (trpo.py:200-227) ------------------------------------------
    x = rng.standard_normal(int(actor.d)).astype(np.float32)
    out["fvp_x"], out["fvp_Fx"] = x, algo._make_F(s_all)(x).numpy()
    out["neglogp_old"] = actor.neglogp(s_all, a_all).numpy()
    out["entropy"] = actor.entropy(s_all).numpy()
    # ---- the update, with its internals observed -----------------------------------------------------------------------
    seen = {}
    orig_cg, orig_bt = trpo_mod.cg, algo._backtrack

    def spy_cg(f_Ax, b, cg_iters=20, residual_tol=1e-10):
        v = orig_cg(f_Ax, b, cg_iters=cg_iters, residual_tol=residual_tol)
        seen["pg_vec"], seen["v_flat"] = np.array(b), np.array(v)
        return v

    def spy_bt(eta_v_flat, *a):
        seen["eta_v_flat"] = np.array(eta_v_flat)
        return orig_bt(eta_v_flat, *a)
    trpo_mod.cg, algo._backtrack = spy_cg, spy_bt
    alg_rng = np.random.default_rng(0)
    perm_seen = []

    class _Rng:
        def shuffle(self_, v):
            alg_rng.shuffle(v)
            perm_seen.append(np.array(v))
    np.random.seed(2000 + seed)
    with _Recorder() as rec:
        log = algo.update((s_all, a_all, adv_all, None, None, None),
                          (expert["sE"], expert["aE"], expert["spE"], eps, models, False, _Rng()))
    trpo_mod.cg = orig_cg
    assert [k for k, _ in rec.calls] == ["normal", "normal"], [k for k, _ in rec.calls]
    out["u3"], out["u4"] = rec.calls[0][1], rec.calls[1][1]
    out["perm"] = perm_seen[0].astype(np.int64)
    for k, v in seen.items():
        out[k] = np.asarray(v)
    for k in ("ent", "tv_pre", "kl_pre", "tv", "kl", "adj", "improve", "norm_pg", "norm_MSE"):
        out["log_" + k] = np.float64(np.asarray(log[k]))
    out["theta_new"] = flat(actor.get_weights())
    np.savez_compressed(os.path.join(OUT, f"ref_{name}.npz"), **out)
    print(f"{name}: adj {float(out['log_adj']):.4f} kl_pre {float(out['log_kl_pre']):.4g} kl {float(out['log_kl']):.4g} "
          f"improve {float(out['log_improve']):.4g} |eta_v| {np.linalg.norm(out['eta_v_flat']):.4g}")


def run_fit_case(name, cfg: NetCfg, n_rows, E, seed, epochs, mbs, max_grad_norm, lr, gaussian=False, setup_kw=None):
    """``SAC_exp._update_models`` (SAC_expert.py:478-621: per-model shuffled minibatches, ``_apply_model_grads`` =
    summed losses, global-norm clip, ONE shared Keras Adam; then the MSE-on-expert bookkeeping) followed by
    ``_expert_preprocess`` with ``scale_epsilon_by_true_MSE`` (:375-404)."""
    st, replay, expert, hyper = make_problem(cfg, 8, E, n_rows, seed=seed, perturb=0.05)
    rng = np.random.default_rng(seed + 900)
    logstd = (0.4 * rng.standard_normal((2, cfg.S)) - 0.3).astype(np.float32) if gaussian else None
    alg, _ = build(cfg, st, replay, expert, hyper, 8, 1, gaussian_logstd=logstd, setup_kw=setup_kw)
    expert["rE"] = rng.standard_normal(E).astype(np.float32)
    alg.model_data.add(replay["s"], replay["a"], replay["r"], replay["sp"], replay["d"])
    alg.expert_data.add(expert["sE"], expert["aE"], expert["rE"], expert["spE"], np.zeros(E))
    alg.model_holdout_ratio, alg.model_num_epochs, alg.model_batch_size = 0.0, epochs, mbs
    alg.model_batch_shuffle, alg.model_max_updates, alg.model_max_grad_norm = True, 10 ** 6, max_grad_norm
    alg.reset_model_optimizer, alg.use_expert_actions = False, False
    import tensorflow as tf
    alg.model_optimizer = tf.keras.optimizers.Adam(learning_rate=lr)
    out = dict(meta=np.array([cfg.S, cfg.A, n_rows, E, seed, epochs, mbs], np.int64),
               hyper=np.array([max_grad_norm, lr], np.float64), in_expert_rE=expert["rE"])
    if gaussian:
        out["in_logstd"] = logstd
        sk = dict(setup_kw or {})
        out["setup"] = np.array([float(sk.get("reward_loss_coef", 1.0)), float(sk.get("delta_clip_loss") or 0.0),
                                 float(bool(sk.get("scale_model_loss", False)))], np.float64)
    for k in ("actor", "m1", "m2"):
        for i, w in enumerate(st[k]):
            out[f"in_{k}_{i}"] = np.asarray(w, np.float32)
    for k in ("s_mean", "s_std", "m_s_mean", "m_s_std", "m_a_mean", "m_a_std", "m_d_mean", "m_d_std"):
        out["in_" + k] = np.asarray(st[k])
    for k, v in replay.items():
        out["in_replay_" + k] = v
    for k in ("sE", "aE", "spE"):
        out["in_expert_" + k] = expert[k]
    np.random.seed(3000 + seed)
    with _Recorder() as rec:
        alg._update_models()
    kinds = [k for k, _ in rec.calls]
    assert kinds == ["shuffle"] * (2 * epochs) + ["normal"], kinds         # per-model shuffles per epoch, then actor.sample
    out["shuffles"] = np.stack([v for k, v in rec.calls if k == "shuffle"]).astype(np.int64)
    out["u_cf"] = rec.calls[-1][1]
    for k, mdl in zip(("m1", "m2"), alg.models):
        out["theta_" + k] = flat(mdl.get_weights()[:6])
        if gaussian:
            out["logstd_" + k] = np.asarray(mdl.get_weights()[6], np.float32).ravel()
    out["mse_expert"] = np.float64(alg.model_MSE_on_expert_data[-1])
    out["mse_counterfactual"] = np.float64(alg.model_MSE_on_expert_counterfactual_action[-1])
    # adaptive expert weight from the bookkeeping above (SAC_expert.py:381-404)
    alg.scale_epsilon_by_true_MSE, alg.epsilon, alg.min_mult, alg.exp_mult, alg.mult_coeff = True, 2.0, True, True, 0.75
    alg.current_reward, alg.expert_reward, alg.expert_batch_size = 120.0, 400.0, None
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        reg = alg._expert_preprocess()
    out["epsilon_coef"] = np.float64(reg[3])
    out["adaptive"] = np.array([2.0, 0.75, 120.0, 400.0], np.float64)
    np.savez_compressed(os.path.join(OUT, f"ref_{name}.npz"), **out)
    print(f"{name}: {len(kinds) - 1} shuffles, mse_expert {float(out['mse_expert']):.5f} "
          f"mse_cf {float(out['mse_counterfactual']):.5f} epsilon_coef {float(out['epsilon_coef']):.5f}")


FIT_CASES = dict(
    # name: (cfg, rows, E, seed, epochs, minibatch, max_grad_norm, lr)
    fit_mse_relu=(NetCfg(S=5, A=2, actor_hidden=(16, 16), critic_hidden=(8, 8), model_hidden=(24, 24), num_models=2),
                  72, 8, 31, 2, 16, 0.05, 1e-3),
    # GaussianModel.get_loss (continuous_models.py:101-131): trainable logstd in the joint optimiser, scale_model_loss
    fit_gauss_tanh=(NetCfg(S=5, A=2, actor_hidden=(16, 16), critic_hidden=(8, 8), model_hidden=(24, 24), num_models=2,
                           model_acts=("tanh", "relu")), 72, 8, 32, 2, 16, 0.5, 1e-3, True,
                    dict(reward_loss_coef=0.5, delta_clip_loss=2.5, scale_model_loss=True)),
)

def run_bc_case(name, cfg: NetCfg, E, seed, K):
    """``BC._update`` (BC.py:298-363): K consecutive actor steps on the expert-observation MSE alone."""
    st, replay, expert, hyper = make_problem(cfg, 8, E, 50, seed=seed, perturb=0.05)
    alg, expert_reg = build(cfg, st, replay, expert, hyper, 8, 1, alg_type="bc")
    out = dict(meta=np.array([cfg.S, cfg.A, E, seed, K], np.int64), lr_pi=np.float64(hyper["lr_pi"]))
    for k in ("actor", "m1", "m2"):
        for i, w in enumerate(st[k]):
            out[f"in_{k}_{i}"] = np.asarray(w, np.float32)
    for i, (m_, v_) in enumerate(zip(st["adam_actor"]["m"], st["adam_actor"]["v"])):
        out[f"in_adam_actor_m_{i}"], out[f"in_adam_actor_v_{i}"] = m_, v_
    out["in_adam_actor_t"] = np.int64(st["adam_actor"]["t"])
    for k in ("s_mean", "s_std", "m_s_mean", "m_s_std", "m_a_mean", "m_a_std", "m_d_mean", "m_d_std"):
        out["in_" + k] = np.asarray(st[k])
    for k, v in expert.items():
        out["in_expert_" + k] = v
    perms = []
    orig = alg.rng.shuffle

    class _Rng:
        def shuffle(self_, x):
            orig(x)
            perms.append(np.array(x))
    alg.rng = _Rng()
    np.random.seed(4000 + seed)
    for step in range(K):
        with _Recorder() as rec:
            alg._update(step, expert_reg)
        assert [k for k, _ in rec.calls] == ["normal", "normal"]
        out[f"step{step}_u3"], out[f"step{step}_u4"] = rec.calls[0][1], rec.calls[1][1]
        out[f"step{step}_perm"] = perms[-1].astype(np.int64)
        out[f"step{step}_g_actor"] = flat(alg.actor_optimizer.last_grads)
        out[f"step{step}_mse"] = np.float32(alg.logger.train_dict["BC_MSE_loss"][-1])
        out[f"step{step}_theta_actor"] = flat(alg.actor.get_weights())
    np.savez_compressed(os.path.join(OUT, f"ref_{name}.npz"), **out)
    print(f"{name}: BC_MSE_loss per step {[float(out[f'step{i}_mse']) for i in range(K)]}")


BC_CASES = dict(
    bc2_relu=(NetCfg(S=5, A=2, actor_hidden=(32, 24), critic_hidden=(8, 8), model_hidden=(24, 24), num_models=2), 8, 41, 3),
)

def run_host_case(name="host_buffers_normalizers"):
    """The reference's pure-NumPy host classes, executed as they are (no emulation involved): TrajectoryBuffer append +
    tail truncation + seeded minibatch draws (buffers.py:41-71, 126-144), RunningNormalizers.update_rms
    (normalizer.py:55-87, 152-163)."""
    from sac_eo.common.buffers import TrajectoryBuffer
    from sac_eo.common.normalizer import RunningNormalizers
    S, A, cap = 4, 2, 50
    rng = np.random.default_rng(77)
    buf = TrajectoryBuffer(S, A, 0.99, 0.95, cap)
    nz = RunningNormalizers(S, A, 0.99)
    out = dict(meta=np.array([S, A, cap], np.int64))
    for t, n in enumerate((20, 25, 30)):
        tr = dict(s=rng.standard_normal((n, S)).astype(np.float32) * 2 + 1, a=rng.uniform(-1, 1, (n, A)).astype(np.float32),
                  r=rng.standard_normal(n).astype(np.float32), sp=rng.standard_normal((n, S)).astype(np.float32),
                  d=rng.random(n) < 0.1)
        for k, v in tr.items():
            out[f"traj{t}_{k}"] = v
        buf.add(tr["s"], tr["a"], tr["r"], tr["sp"], tr["d"])
        nz.update_rms(tr["s"], tr["a"], tr["r"], tr["sp"])
        out[f"after{t}_size"] = np.array([buf.current_size, buf.traj_total, buf.steps_total], np.int64)
    for k in ("s_all", "a_all", "r_all", "sp_all", "d_all", "idx_all"):
        out["buf_" + k] = getattr(buf, k)
    np.random.seed(7)
    for k, v in zip(("s", "a", "sp", "r", "d"), buf.get_offmodel_info(batch_size=16)):
        out["off_" + k] = v
    for k, v in zip(("s", "a", "sp", "r"), buf.get_model_info(batch_size=8)):
        out["mod_" + k] = v
    out["states"] = buf.get_states(batch_size=4)
    for nm_, r_ in zip(("s", "a", "r", "delta", "ret"), nz.get_rms()):
        out[f"rms_{nm_}_mean"], out[f"rms_{nm_}_var"], out[f"rms_{nm_}_std"] = (np.asarray(r_.mean), np.asarray(r_.var),
                                                                           np.asarray(r_.std))
        out[f"rms_{nm_}_t"] = np.int64(r_.t_last)
    x = rng.standard_normal((5, S)).astype(np.float32)
    out["norm_x"], out["norm_y"] = x, nz.s_rms.normalize(x)
    out["denorm_y"] = nz.delta_rms.denormalize(x)
    # ---- trajectory_sampler (samplers.py:3-77) on a deterministic stub environment / actor --------------------------------
    from sac_eo.common.samplers import trajectory_sampler
    for tag, (horizon, term_at, ev) in dict(trunc=(6, None, True), term=(9, 4, True), plain=(5, None, False)).items():
        res = trajectory_sampler(StubEnv(S, term_at), StubActor(A), horizon, eval=ev)
        for k, v in zip(("s", "a", "r", "sp", "d", "J"), res):
            out[f"samp_{tag}_{k}"] = np.asarray(v)
    # ---- the command line: every flag of the reference's parser with its default (train_parser.py) ---------------------
    import json
    from sac_eo.common.train_parser import create_train_parser, all_kwargs
    args = vars(create_train_parser().parse_args([]))
    json.dump(dict(defaults={k: (v if isinstance(v, (int, float, str, bool, list, type(None))) else repr(v)) for k, v in args.items()},
                   groups={k: list(v) for k, v in all_kwargs.items()}),
              open(os.path.join(OUT, "ref_train_parser.json"), "w"), indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(OUT, f"ref_{name}.npz"), **out)
    print(f"{name}: buffer size {buf.current_size}, s_rms.std {nz.s_rms.std}, {len(args)} parser flags")


class StubEnv:
    """Deterministic linear environment for the sampler pin (shared with tests/test_reference_pin.py by construction)."""

    def __init__(self, S, term_at=None):
        self.S, self.term_at, self.t = S, term_at, 0

    def reset(self, s_init=None):
        self.t = 0
        self.s = np.arange(self.S, dtype=np.float64) * 0.1 if s_init is None else np.asarray(s_init, np.float64)
        return self.s

    def step(self, a):
        self.t += 1
        self.s = 0.9 * self.s + 0.05 * float(np.sum(a)) + 0.01 * self.t
        return self.s, float(np.sum(self.s)) * 0.5, (self.term_at is not None and self.t >= self.term_at), {}


class StubActor:
    class _T:
        def __init__(self, v):
            self.v = v

        def numpy(self):
            return self.v

    def __init__(self, A):
        self.A = A

    def sample(self, s, deterministic=False):
        return self._T(np.tanh(np.asarray(s, np.float64)[:self.A] * 3.0) * 1.5)

    def clip(self, a):
        return np.clip(a, -1.0, 1.0)


def run_ppo_case(name, cfg: NetCfg, N, seed, update_it, nminibatch, eps_ppo, max_grad_norm, lr):
    """``PPO.update(rollout_data)`` (ppo.py:41-119, 121-238 with expert_reg = None): epochs x shuffled minibatches,
    per-minibatch advantage normalisation, clipped surrogate, global-norm clip, Keras Adam on the actor."""
    import torch
    from sac_eo.actors import init_actor
    from sac_eo.algs.model_free import ppo as ppo_mod
    from sac_eo.common.normalizer import RunningNormalizers
    from sac_eo.common.train_parser import create_train_parser
    from sac_eo.common.train_utils import gather_inputs
    from oracle.sac_eo_oracle import gaussian_forward, to_torch_state
    st, replay, expert, hyper = make_problem(cfg, 8, 4, max(N, 300), seed=seed, perturb=0.2)
    inputs = gather_inputs(create_train_parser().parse_args(["--alg_type", "mbrl", "--mf_algo", "ppo"]))
    ak, uk = inputs["actor_kwargs"], inputs["mf_update_kwargs"]
    ak.update(actor_layers=list(cfg.actor_hidden), actor_activations=list(cfg.actor_acts), actor_weights=None,
              actor_per_state_std=cfg.per_state_std, actor_squash=False, actor_std_mult=cfg.std_mult)
    uk.update(actor_lr=lr, actor_update_it=update_it, actor_nminibatch=nminibatch, eps_ppo=eps_ppo,
              max_grad_norm=max_grad_norm, adaptlr=False, ent_reg=False, ent_targ=-cfg.A, adv_center=True, adv_scale=True)
    actor = init_actor(_Env(cfg.S, cfg.A), **ak)
    actor.set_weights([np.asarray(w) for w in st["actor"]])
    nz = RunningNormalizers(cfg.S, cfg.A, 0.99)
    nz.s_rms.mean, nz.s_rms.std = st["s_mean"].copy(), st["s_std"].copy()
    actor.set_rms(nz)
    algo = ppo_mod.PPO(actor, uk)
    rng = np.random.default_rng(seed + 500)
    s_all = replay["s"][:N]
    th64 = to_torch_state(st, torch.float64)
    with torch.no_grad():
        mean, ls = gaussian_forward(cfg, th64["actor"], torch.as_tensor(s_all, dtype=torch.float64), th64)
    a_all = (mean + torch.exp(ls) * torch.from_numpy(rng.standard_normal((N, cfg.A)))).numpy().astype(np.float32)
    adv_all = (rng.standard_normal(N) * 1.5 + 0.3).astype(np.float32)
    out = dict(meta=np.array([cfg.S, cfg.A, N, seed, update_it, nminibatch, int(cfg.per_state_std)], np.int64),
               hyper=np.array([eps_ppo, max_grad_norm, lr, cfg.std_mult], np.float64), s_all=s_all, a_all=a_all, adv_all=adv_all)
    for i, w in enumerate(st["actor"]):
        out[f"in_actor_{i}"] = np.asarray(w, np.float32)
    out["in_s_mean"], out["in_s_std"] = st["s_mean"], st["s_std"]
    np.random.seed(5000 + seed)
    with _Recorder() as rec:
        log = algo.update((s_all, a_all, adv_all, None, None, None))
    assert [k for k, _ in rec.calls] == ["shuffle"] * update_it
    out["shuffles"] = np.stack([v for _, v in rec.calls]).astype(np.int64)
    for k in ("ent", "tv", "kl", "outside_clip", "actor_grad_norm_pre", "actor_grad_norm"):
        out["log_" + k] = np.float64(np.asarray(log[k]))
    out["theta_new"] = flat(actor.get_weights())
    ms, vs = zip(*(algo.actor_optimizer.get_slot_arrays(v) for v in actor.trainable))
    out["adam_m"], out["adam_v"] = flat(ms), flat(vs)
    np.savez_compressed(os.path.join(OUT, f"ref_{name}.npz"), **out)
    print(f"{name}: " + " ".join(f"{k} {float(out['log_' + k]):.4g}" for k in ("tv", "kl", "outside_clip", "actor_grad_norm_pre",
                                                                          "actor_grad_norm")))


PPO_CASES = dict(
    # name: (cfg, N, seed, update_it, nminibatch, eps_ppo, max_grad_norm, lr)
    ppo_psd_tanh=(NetCfg(S=9, A=3, actor_hidden=(32, 24), critic_hidden=(8, 8), model_hidden=(8, 8), per_state_std=True,
                         actor_acts=("tanh", "tanh"), std_mult=0.7, num_models=0), 100, 51, 2, 4, 0.2, 0.5, 3e-3),
)

def run_disc_case(name="disc_saceo2"):
    """``SAC_exp._calc_disc`` and the disagreement-scaled expert weight of ``_expert_preprocess`` (SAC_expert.py:405-460):
    counterfactual actions ``tf_clip(actor.sample(sE))`` through BOTH models, row-wise L2 disagreement, max / median /
    total, ``eps = 1 / (epsilon * max_disc + 1)``.  Same problem as the ``saceo2_relu`` case."""
    import contextlib
    cfg, B, E, N, seed, K, eps, tui, _, _ = CASES["saceo2_relu"]
    st, replay, expert, hyper = make_problem(cfg, B, E, N, seed=seed, perturb=0.05)
    hyper["eps"] = eps
    alg, _ = build(cfg, st, replay, expert, hyper, B, tui)
    alg.expert_data.add(expert["sE"], expert["aE"], np.zeros(E, np.float32), expert["spE"], np.zeros(E))
    alg.scale_max_disc, alg.epsilon, alg.expert_batch_size = True, 2.0, None
    out = dict(epsilon=np.float64(2.0))
    np.random.seed(6000 + seed)
    with _Recorder() as rec, contextlib.redirect_stdout(io.StringIO()):
        reg = alg._expert_preprocess()
        ratio, mx, med, tot = alg._calc_disc(expert["sE"], expert["aE"], expert["spE"])
    assert [k for k, _ in rec.calls] == ["normal", "normal"]
    out["u_pre"], out["u_disc"] = rec.calls[0][1], rec.calls[1][1]
    out["epsilon_coef"] = np.float64(reg[3])
    out["disc_ratio"], out["max_disc"], out["median_disc"], out["total_disc"] = (np.asarray(ratio), np.float64(mx),
                                                                                 np.float64(med), np.float64(tot))
    np.savez_compressed(os.path.join(OUT, f"ref_{name}.npz"), **out)
    print(f"{name}: epsilon_coef {float(reg[3]):.6f} max {float(mx):.5f} median {float(med):.5f} total {float(tot):.5f}")


def run_quirks(name="quirks"):
    """Behaviours of the reference that DESIGN.md / the oracle docstrings cite, observed by running its code: which
    call patterns raise.  Stored as exception type names (the emulation raises Python / torch errors where TensorFlow
    would raise its own; only "raises or not" and the Python-level error names are meaningful)."""
    import json
    rec = {}

    def attempt(key, fn):
        try:
            fn()
            rec[key] = None
        except Exception as e:                                  # noqa: BLE001 - the type is the datum
            rec[key] = type(e).__name__

    # TRPO.update without expert_reg: grad_final is only defined inside the expert branches (trpo.py:107-111, 154-158)
    cfg, N, E, seed, eps, delta, cg_it, klf, damp = TRPO_CASES["trpo_psd_tanh"]
    import torch
    from sac_eo.actors import init_actor
    from sac_eo.algs.model_free import trpo as trpo_mod
    from sac_eo.common.normalizer import RunningNormalizers
    from sac_eo.common.train_parser import create_train_parser
    from sac_eo.common.train_utils import gather_inputs
    from sac_eo.models import init_world_models
    st, replay, expert, hyper = make_problem(cfg, 8, E, 300, seed=seed, perturb=0.2)
    inputs = gather_inputs(create_train_parser().parse_args(["--alg_type", "mbrl"]))
    ak, mk, msk, uk = (inputs[k] for k in ("actor_kwargs", "model_kwargs", "model_setup_kwargs", "mf_update_kwargs"))
    ak.update(actor_layers=list(cfg.actor_hidden), actor_activations=list(cfg.actor_acts), actor_weights=None,
              actor_per_state_std=True, actor_squash=False)
    mk.update(model_layers=list(cfg.model_hidden), model_activations=list(cfg.model_acts), model_weights=None,
              reward_weights=None, num_models=2, gaussian_model=False)
    uk.update(ent_reg=False, ent_targ=-cfg.A)
    env = _Env(cfg.S, cfg.A)
    actor = init_actor(env, **ak)
    models = init_world_models(env, **mk, model_setup_kwargs=msk)
    nz = RunningNormalizers(cfg.S, cfg.A, 0.99)
    actor.set_rms(nz)
    for m_ in models:
        m_.set_rms(nz)
    algo = trpo_mod.TRPO(actor, uk)
    s_all = replay["s"][:32]
    a_all = np.random.default_rng(0).uniform(-1, 1, (32, cfg.A)).astype(np.float32)
    adv = np.random.default_rng(1).standard_normal(32).astype(np.float32)
    roll = (s_all, a_all, adv, None, None, None)

    class _Rng:
        def shuffle(self_, v):
            np.random.default_rng(0).shuffle(v)
    np.random.seed(0)
    attempt("trpo_update_without_expert_reg", lambda: algo.update(roll))
    attempt("trpo_update_one_model_branch",
            lambda: algo.update(roll, (expert["sE"], expert["aE"], expert["spE"], 0.3, models[:1], False, _Rng())))
    attempt("trpo_update_two_model_branch",
            lambda: algo.update(roll, (expert["sE"], expert["aE"], expert["spE"], 0.3, models, False, _Rng())))
    # SAC-EO two-model branch with an odd number of expert rows: the two halves differ in length (SAC_expert.py:329-332)
    cfg2, B, E2, N2, seed2, K, eps2, tui, _, _ = CASES["saceo2_relu"]
    st2, replay2, expert2, hyper2 = make_problem(cfg2, B, 7, N2, seed=seed2, perturb=0.05)
    alg, reg = build(cfg2, st2, replay2, expert2, hyper2, B, tui)
    attempt("saceo_update_odd_expert_rows_two_models", lambda: alg._update(0, reg))
    st3, replay3, expert3, hyper3 = make_problem(cfg2, B, 8, N2, seed=seed2, perturb=0.05)
    alg, reg = build(cfg2, st3, replay3, expert3, hyper3, B, tui)
    alg.alpha.assign(-3.0)                                      # raw temperature may be negative; clamped AFTER its step (:348)
    attempt("saceo_update_even_expert_rows_two_models", lambda: alg._update(0, reg))
    rec["alpha_after_update_from_minus_3"] = float(alg.alpha.numpy())
    json.dump(rec, open(os.path.join(OUT, "ref_quirks.json"), "w"), indent=1, sort_keys=True)
    print(name, rec)


TRPO_CASES = dict(
    # name: (cfg, N, E, seed, eps, delta, cg_it, kl_maxfactor, trust_damp)
    trpo_psd_tanh=(NetCfg(S=9, A=3, actor_hidden=(32, 24), critic_hidden=(8, 8), model_hidden=(24, 24), per_state_std=True,
                          actor_acts=("tanh", "tanh"), std_mult=0.7), 96, 8, 21, 0.5, 0.02, 10, 1.5, 0.01),
    trpo_sis_relu=(NetCfg(S=9, A=3, actor_hidden=(32, 24), critic_hidden=(8, 8), model_hidden=(24, 24), per_state_std=False,
                          actor_acts=("relu", "relu"), std_mult=0.7), 96, 8, 22, 0.2, 0.02, 10, 0.6, 0.01),
)

CASES = dict(
    # name: (cfg, B, E, N, seed, K, eps, target_update_int, store_inputs, proj_dim)
    saceo2_relu=(NetCfg(S=5, A=2, actor_hidden=(32, 24), critic_hidden=(32, 24), model_hidden=(32, 24), num_models=2),
                 32, 8, 300, 11, 3, 0.3, 2, True, 0),
    saceo1_tanh_elu_sis=(NetCfg(S=6, A=3, actor_hidden=(24, 32), critic_hidden=(32, 32), model_hidden=(24, 24),
                                actor_acts=("tanh", "tanh"), critic_acts=("elu", "elu"), model_acts=("tanh", "tanh"),
                                per_state_std=False, num_models=1, delta_clip_pred=0.05),
                         24, 6, 200, 12, 3, 0.25, 1, True, 0),
    sac_plain_relu=(NetCfg(S=5, A=2, actor_hidden=(32, 24), critic_hidden=(32, 24), model_hidden=(8, 8),
                           num_models=0), 32, 8, 300, 13, 3, 0.0, 2, True, 0),
    saceo2_hopper_256=(NetCfg(S=11, A=3), 256, 20, 2000, 14, 2, 1e-3, 1, False, 48),
    # separate_reward_nn (base_world_model.py:34-41, 72-74): the dynamics net predicts the S delta columns only
    saceo2_sepreward=(NetCfg(S=5, A=2, actor_hidden=(32, 24), critic_hidden=(32, 24), model_hidden=(32, 24), num_models=2,
                             separate_reward_nn=True, model_acts=("elu", "tanh")), 32, 8, 300, 15, 2, 0.4, 1, True, 0),
)


if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CASES) + list(TRPO_CASES) + list(FIT_CASES) + list(BC_CASES) + list(PPO_CASES)
                 + ["host_buffers_normalizers", "disc_saceo2", "quirks"]):
        if name in CASES:
            cfg, *rest = CASES[name]
            run_case(name, cfg, *rest)
        elif name in PPO_CASES:
            cfg, *rest = PPO_CASES[name]
            run_ppo_case(name, cfg, *rest)
        elif name == "quirks":
            run_quirks()
        elif name == "disc_saceo2":
            run_disc_case()
        elif name == "host_buffers_normalizers":
            run_host_case()
        elif name in BC_CASES:
            cfg, *rest = BC_CASES[name]
            run_bc_case(name, cfg, *rest)
        elif name in FIT_CASES:
            cfg, *rest = FIT_CASES[name]
            run_fit_case(name, cfg, *rest)
        else:
            cfg, *rest = TRPO_CASES[name]
            run_trpo_case(name, cfg, *rest)
