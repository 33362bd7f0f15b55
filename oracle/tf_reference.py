"""Cross-check hook against the REAL reference (TensorFlow eager) - TEST INFRASTRUCTURE, like everything under oracle/.

Nothing here can run in the build container or on a stock GPU box: TensorFlow and gym are not installable offline
(BASELINE.md 3).  ``bench.py`` / ``tests/test_oracle.py`` probe for them (and for the reference tree
under ``baseline/_ref`` or ``/root/reference``); only when ALL are present this module
  (1) builds the reference's own ``SquashedGaussianActor`` / ``QCritic`` x4 / ``MSEModel`` x2 / ``SAC_exp`` around a stub
      environment with the benchmark's spaces (``train.py:60-105`` without ``init_env``),
  (2) fills ``env_data`` / ``expert_data`` with the oracle problem's rows and weights,
  (3) runs ``SAC_exp._update`` once with a seeded global NumPy RNG and compares the updated parameters with
      ``oracle.sac_eo_update`` fed with the draws that the same RNG state produces (consumption order of SURVEY.md
      App. A) - this would extend the pin of tests/test_reference_pin.py (reference code over oracle/tfemu) to
      TensorFlow's own kernels,
  (4) times the reference's ``_update`` for the CPU arm (``cpu_baseline.kind = "reference (TF eager)"``).
It has never been executed (no TensorFlow anywhere in this project's environments); every failure is reported to the
caller as a reason string, never raised."""
import os
import sys
import time

import numpy as np


def available(tree):
    try:
        import tensorflow  # noqa: F401
        import gym  # noqa: F401
    except Exception as e:
        return False, "import failed: %s" % type(e).__name__
    if not tree or not os.path.isdir(os.path.join(tree, "sac_eo")):
        return False, "reference tree not found"
    return True, "ok"


class _StubEnv:
    """observation_space / action_space of the benchmark shape; never stepped."""

    def __init__(self, S, A):
        import gym
        self.observation_space = gym.spaces.Box(-np.inf, np.inf, (S,), np.float32)
        self.action_space = gym.spaces.Box(-1.0, 1.0, (A,), np.float32)

    def seed(self, s):
        pass


def build_reference(tree, S, A, problem, extra_args=()):
    """-> (alg, expert_reg) with the oracle problem's weights, rows and hyper-parameters loaded."""
    if tree not in sys.path:
        sys.path.insert(0, tree)
    from sac_eo.actors import init_actor
    from sac_eo.algs import init_alg
    from sac_eo.common.train_parser import create_train_parser
    from sac_eo.common.train_utils import gather_inputs
    from sac_eo.critics import init_critics
    from sac_eo.models import init_world_models
    st, replay, expert, hyper = problem
    argv = ["--alg_type", "sac_imit", "--actor_squash", "--actor_per_state_std", "--actor_layers", "256", "256",
            "--critic_layers", "256", "256", "--actor_activations", "relu", "--critic_activations", "relu",
            "--gamma", str(hyper["gamma"]), "--soft_tau", str(hyper["tau"]), "--q_crit_lr", str(hyper["lr_q"]),
            "--mbpo_actor_lr", str(hyper["lr_pi"]), "--mbpo_alpha_lr", str(hyper["lr_alpha"]), "--epsilon", str(hyper["eps"]),
            "--expert_buffer_size", str(len(expert["sE"]))] + list(extra_args)
    inputs = gather_inputs(create_train_parser().parse_args(argv))
    for grp, key in (("actor_kwargs", "actor_weights"), ("critic_kwargs", "critic_weights"),
                     ("model_kwargs", "model_weights"), ("model_kwargs", "reward_weights")):
        inputs[grp][key] = None
    inputs["alg_kwargs"]["init_rms_stats"] = None
    inputs["alg_kwargs"]["alg_seed"] = 0
    env = _StubEnv(S, A)
    actor = init_actor(env, **inputs["actor_kwargs"])
    expert_actor = init_actor(env, **inputs["actor_kwargs"])
    critics, q_targets, q_critics = init_critics(env, **inputs["critic_kwargs"])
    models = init_world_models(env, **inputs["model_kwargs"], model_setup_kwargs=inputs["model_setup_kwargs"])
    alg = init_alg(0, env, env, env, actor, critics, q_targets, q_critics, models, inputs["alg_kwargs"],
                   inputs["mf_update_kwargs"], expert_actor, None)
    actor.set_weights([np.asarray(w) for w in st["actor"]])
    for net, key in zip(q_critics, ("q1", "q2")):
        net.set_weights([np.asarray(w) for w in st[key]])
    for net, key in zip(q_targets, ("t1", "t2")):
        net.set_weights([np.asarray(w) for w in st[key]])
    for net, key in zip(models, ("m1", "m2")):
        net.set_weights([np.asarray(w) for w in st[key]])
    alg.alpha.assign(float(st["alpha"]))
    alg.env_data.add(replay["s"], replay["a"], replay["r"], replay["sp"], replay["d"])
    expert_reg = (expert["sE"], np.zeros((len(expert["sE"]), A), np.float32), expert["spE"], hyper["eps"], False)
    return alg, expert_reg


def crosscheck_and_time(tree, shape=(11, 3), B=256, E=20, seconds=10.0):
    """-> dict(ok, reason, max_rel (oracle vs reference after one update), updates_per_s)."""
    ok, why = available(tree)
    if not ok:
        return {"ok": False, "reason": why}
    try:
        from oracle.sac_eo_oracle import NetCfg, make_problem, sac_eo_update, to_torch_state
        S, A = shape
        cfg = NetCfg(S=S, A=A)
        problem = make_problem(cfg, B, E, 5000, seed=0, identity_norm=True)
        st, replay, expert, hyper = problem
        alg, expert_reg = build_reference(tree, S, A, problem)
        # the draws SAC_exp._update will make from the global NumPy RNG and alg.rng, replayed for the oracle
        state = np.random.get_state()
        rng_state = alg.rng.bit_generator.state
        idx = np.random.randint(alg.env_data.current_size, size=B)
        u1 = np.random.normal(size=(B, A))
        order = np.arange(E); alg.rng.shuffle(order); I1, I2 = np.array_split(order, 2)
        u2 = np.random.normal(size=(B, A)); u3 = np.random.normal(size=(len(I1), A)); u4 = np.random.normal(size=(len(I2), A))
        u5 = np.random.normal(size=(B, A))
        np.random.set_state(state); alg.rng.bit_generator.state = rng_state
        batch = dict(idx=idx, u1=u1.astype(np.float32), u2=u2.astype(np.float32), u3=u3.astype(np.float32),
                     u4=u4.astype(np.float32), u5=u5.astype(np.float32), I1=I1, I2=I2)
        ref_before = [np.array(w) for w in alg.actor.get_weights()]
        alg._update(0, expert_reg)
        o = sac_eo_update(cfg, to_torch_state(st), batch, hyper)
        worst = 0.0
        for got, new, old in zip(alg.actor.get_weights(), o["new"]["actor"], ref_before):
            d = new.numpy() - old
            if np.linalg.norm(d) > 0:
                worst = max(worst, float(np.linalg.norm(np.asarray(got) - old - d) / np.linalg.norm(d)))
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            alg._update(n + 1, expert_reg)
            n += 1
        return {"ok": True, "reason": "ok", "max_rel_dtheta_actor": worst, "updates_per_s": n / (time.perf_counter() - t0)}
    except Exception as e:          # reported, never raised: this path has never met a real TensorFlow
        return {"ok": False, "reason": "%s: %s" % (type(e).__name__, e)}
