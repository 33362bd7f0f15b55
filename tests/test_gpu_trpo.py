"""GPU: TRPO surrogate gradient, line-search statistics, parameter step and the whole TRPO.update against the oracle
(trpo.py:36-198, :229-317; GaussianActor interface continuous_actors.py:74-100,137-192).  All through the C ABI."""
import math

import numpy as np
import pytest
import torch

from oracle import sac_eo_oracle as O
from sac_expert_b200 import lib
from sac_expert_b200.population import Population
from tests.helpers import rel, spec_from_cfg

pytestmark = pytest.mark.gpu


def _setup(per_state_std, acts, n=2, N=96, gemm_mode=lib.GEMM_FP32_SIMT, hidden=(64, 48), S=9, A=3, seed=20, perturb=0.2):
    cfg = O.NetCfg(S=S, A=A, actor_hidden=hidden, critic_hidden=(16, 16), num_models=0, per_state_std=per_state_std,
                   actor_acts=acts, std_mult=0.7)
    pop = Population(spec_from_cfg(cfg, n, 8, 0, 16, fvp_rows=N, gemm_mode=gemm_mode))
    probs = []
    rng = np.random.default_rng(seed)
    for i in range(n):
        st, replay, _, hyper = O.make_problem(cfg, 8, 2, 300, seed=seed + i, perturb=perturb)
        pop.load_agent(i, st, hyper)
        s = replay["s"][:N]
        th = O.to_torch_state(st, torch.float64)
        with torch.no_grad():       # rollout actions come from the policy itself (ratios near 1, as in TRPO's use)
            mean, ls = O.gaussian_forward(cfg, th["actor"], torch.as_tensor(s, dtype=torch.float64), th)
        a = (mean + torch.exp(ls) * torch.from_numpy(rng.standard_normal((N, A)))).numpy().astype(np.float32)
        pop.t["fvp_states"][i].copy_(torch.from_numpy(s))
        probs.append((st, s, a, rng.standard_normal(N).astype(np.float32) * (1 + i)))
    return cfg, pop, probs


@pytest.mark.parametrize("per_state_std,acts,mode", [
    (True, ("relu", "relu"), lib.GEMM_FP32_SIMT), (False, ("tanh", "tanh"), lib.GEMM_FP32_SIMT),
    (True, ("elu", "tanh"), lib.GEMM_FP32_SIMT), (True, ("tanh", "tanh"), lib.GEMM_TCGEN05_BF16X3)])
def test_surrogate_gradient_and_eval(per_state_std, acts, mode):
    # reference sizes on the tensor-core engine.  perturb stays small there: with 256-wide layers a 0.2 perturbation drives
    # many softplus stds to the 1e-3 floor, and the gradient's 1/std^2 weights then amplify the last-bit error of the
    # forward means by ~1e3 on ANY fp32 engine (measured: 1.2e-3 on the fp32 SIMT engine, 3.9e-3 on tcgen05, while the
    # Fisher-vector product on the same activations is at 1.5e-6 / 1.0e-5; tools/trpo_tc_check.py)
    big = dict(hidden=(256, 256), S=27, A=8, N=256, perturb=0.02) if mode == lib.GEMM_TCGEN05_BF16X3 else {}
    cfg, pop, probs = _setup(per_state_std, acts, gemm_mode=mode, **big)
    n, N, A, L = len(probs), len(probs[0][1]), cfg.A, pop.L
    rng = np.random.default_rng(5)
    act = np.stack([p[2] for p in probs]); adv = np.stack([O.trpo_normalise_adv(p[3]) for p in probs]).astype(np.float32)
    alpha = np.array([0.3, 0.0], np.float32)
    refs = []
    nlp_old = np.zeros((n, N), np.float32); kl_ref = np.zeros((n, N, A, 2), np.float32)
    for i, (st, s, a, _) in enumerate(probs):
        th = O.to_torch_state(st, torch.float64)
        e0 = O.trpo_eval(cfg, th["actor"], s, a, adv[i], np.zeros(N), None, th)
        nlp_old[i] = (e0["nlp"].numpy() + rng.normal(size=N) * 0.1).astype(np.float32)
        info = e0["kl_info"].numpy().copy()
        info[..., 0] += rng.normal(size=(N, A)) * 0.05; info[..., 1] += rng.normal(size=(N, A)) * 0.05
        kl_ref[i] = info.astype(np.float32)
        g, ag, _ = O.trpo_surrogate_grad(cfg, th["actor"], s, a, adv[i], nlp_old[i].astype(np.float64), float(alpha[i]), 0.0, th)
        e = O.trpo_eval(cfg, th["actor"], s, a, adv[i], nlp_old[i].astype(np.float64), kl_ref[i].astype(np.float64), th)
        refs.append((O.flat(g).numpy(), e, e0))
    grad, gstats = pop.trpo_grad(act, adv, nlp_old, alpha)
    ev = pop.trpo_eval(act, adv, nlp_old, kl_ref, want_nlp=True, want_kl_info=True, want_rows=True)
    grad, gstats = grad.cpu().numpy(), gstats.cpu().numpy()
    stats = ev["stats"].cpu().numpy()
    tol = 1e-3
    for i in range(n):
        g_ref, e, e0 = refs[i]
        assert rel(grad[i, :L.na], g_ref) < tol
        assert np.all(grad[i, L.na:] == 0)
        assert rel(ev["nlp"][i].cpu().numpy(), e["nlp"].numpy()) < tol
        assert rel(ev["kl_info"][i].cpu().numpy(), e0["kl_info"].numpy()) < tol
        for k, key in enumerate(("surr", "kl", "tv", "ent")):
            assert abs(stats[i, k] - float(e[key])) <= tol * max(abs(float(e[key])), 1e-2), key
        assert abs(gstats[i, 0] - float(e["surr"])) <= tol * max(abs(float(e["surr"])), 1e-2)
        assert abs(gstats[i, 3] - float(e["ent"])) <= tol * abs(float(e["ent"]))
        rows = ev["rows"][i].cpu().numpy()
        assert abs(rows[:, 3].mean() - float(e["ent"])) <= tol * abs(float(e["ent"]))
    # nlp_old = NULL means ratio 1: the gradient at the reference policy itself
    g1, _ = pop.trpo_grad(act, adv, None, None)
    for i, (st, s, a, _) in enumerate(probs):
        th = O.to_torch_state(st, torch.float64)
        e0 = O.trpo_eval(cfg, th["actor"], s, a, adv[i], np.zeros(N), None, th)
        g, _, _ = O.trpo_surrogate_grad(cfg, th["actor"], s, a, adv[i], e0["nlp"], 0.0, 0.0, th)
        assert rel(g1[i, :L.na].cpu().numpy(), O.flat(g).numpy()) < tol
    pop.close()


@pytest.mark.parametrize("per_state_std", [True, False])
def test_actor_step_is_exact_and_floors_the_logstd_variable(per_state_std):
    cfg, pop, probs = _setup(per_state_std, ("tanh", "tanh"))
    L, n = pop.L, len(probs)
    rng = np.random.default_rng(2)
    ref = pop.t["actor"].clone()
    d = torch.from_numpy(rng.standard_normal((n, L.na_stride)).astype(np.float32)).cuda()
    if not per_state_std:
        d[:, L.na - cfg.A] = -1e4
    scale = np.array([0.25, -1.5], np.float32)
    pop.actor_step(ref, d, scale)
    want = ref.cpu().numpy() + scale[:, None] * d.cpu().numpy()          # fp32: one product, one sum
    if not per_state_std:
        want[:, L.na - cfg.A:L.na] = np.maximum(want[:, L.na - cfg.A:L.na], np.float32(math.log(1e-3)))
    got = pop.t["actor"].cpu().numpy()
    assert got[:, :L.na].tobytes() == want[:, :L.na].astype(np.float32).tobytes()
    assert np.array_equal(got[:, L.na:], ref.cpu().numpy()[:, L.na:])    # padding untouched
    pop.close()


@pytest.mark.parametrize("per_state_std,acts,kl_maxfactor", [(True, ("tanh", "tanh"), 1.5), (False, ("relu", "tanh"), 1.5),
                                                              (True, ("tanh", "tanh"), 0.6), (True, ("tanh", "tanh"), 1e-9)])
def test_trpo_update_matches_the_oracle(per_state_std, acts, kl_maxfactor):
    cfg, pop, probs = _setup(per_state_std, acts, n=3, N=128)
    L = pop.L
    act = np.stack([p[2] for p in probs]); adv = np.stack([p[3] for p in probs])
    before = pop.t["actor"].cpu().numpy().copy()
    logs = pop.trpo_update(act, adv, delta=0.02, cg_iters=5, trust_damp=0.01, kl_maxfactor=kl_maxfactor)
    after = pop.t["actor"].cpu().numpy()
    for i, (st, s, a, ad) in enumerate(probs):
        th = O.to_torch_state(st, torch.float64)
        new, log, _, _ = O.trpo_update(cfg, th["actor"], s, a, ad, th, delta=0.02, cg_iters=5, trust_damp=0.01,
                                       kl_maxfactor=kl_maxfactor)
        assert abs(logs[i]["adj"] - log["adj"]) < 1e-6, (logs[i], log)
        step_ref = (O.flat(new) - O.flat(th["actor"])).numpy()
        if log["adj"] == 0:
            assert np.array_equal(after[i], before[i])
        else:
            assert rel(after[i, :L.na] - before[i, :L.na], step_ref) < 1e-2        # 5 fp32 CG iterations vs fp64
        for key in ("ent", "tv_pre", "kl_pre", "tv", "kl", "improve"):
            assert abs(logs[i][key] - log[key]) <= 2e-2 * max(abs(log[key]), 1e-3), (key, logs[i], log)
    pop.close()


def test_trpo_class_interface():
    """``TRPO(actor, update_kwargs).update(rollout_data)`` and the GaussianActor accessors through the mirror classes."""
    from sac_expert_b200.sac_eo.actors.init_actor import init_actor
    from sac_expert_b200.sac_eo.algs.model_free.trpo import TRPO
    from sac_expert_b200.sac_eo.envs.synthetic import SyntheticEnv
    np.random.seed(0)
    env = SyntheticEnv(7, 2)
    actor = init_actor(env, [32, 32], ["tanh"], 0.01, 1.0, "orthogonal", False, None, actor_per_state_std=False,
                       actor_squash=True)
    rng = np.random.default_rng(1)
    N = 80
    s = rng.standard_normal((N, 7)).astype(np.float32); a = rng.standard_normal((N, 2)).astype(np.float32)
    adv = rng.standard_normal(N).astype(np.float32)
    cfg = O.NetCfg(S=7, A=2, actor_hidden=(32, 32), critic_hidden=(8, 8), num_models=0, per_state_std=False,
                   actor_acts=("tanh", "tanh"), std_mult=1.0)
    w0 = actor.get_weights()
    st = {"actor": [torch.from_numpy(w.astype(np.float64)) for w in w0], "s_mean": torch.zeros(7, dtype=torch.float64),
          "s_std": torch.ones(7, dtype=torch.float64)}
    e = O.trpo_eval(cfg, st["actor"], s, a, adv, np.zeros(N), None, st)
    assert rel(actor.neglogp(s, a).numpy(), e["nlp"].numpy()) < 1e-4
    assert rel(actor.get_kl_info(s), e["kl_info"].numpy()) < 1e-4
    assert abs(float(actor.entropy(s).mean()) - float(e["ent"])) < 1e-4 * abs(float(e["ent"]))
    assert float(actor.kl(s, actor.get_kl_info(s)).abs().max()) < 1e-6
    kw = dict(adv_center=True, adv_scale=True, delta_trpo=0.02, cg_it=5, trust_sub=1, trust_damp=0.01, kl_maxfactor=1.5,
              ent_reg=True, ent_targ=0.0, alpha_lr=0.01)
    alg = TRPO(actor, kw)
    log = alg.update((s, a, adv, None, None, None))
    new, ref, _, _ = O.trpo_update(cfg, st["actor"], s, a, adv, st, delta=0.02, cg_iters=5)
    assert abs(log["adj"] - ref["adj"]) < 1e-6
    w1 = actor.get_weights()
    assert rel(np.concatenate([x.ravel() for x in w1]) - np.concatenate([x.ravel() for x in w0]),
               (O.flat(new) - O.flat(st["actor"])).numpy()) < 1e-2
    # entropy above the target of 0 pushes the temperature down; it is clamped at 0 (:169-172)
    assert log["alpha"] == 0.0 and log["epsilon"] == 0.0
    with pytest.raises(NotImplementedError):
        alg.update((s, a, adv, None, None, None), expert_reg=(s, a, s, 0.5, [], False, None))


# ----------------------------------------------------------------------------------------------------------
# PPO (ppo.py:41-119, :122-237)
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("per_state_std,max_norm", [(True, None), (False, 0.05)])
def test_ppo_gradient_clip_and_actor_adam(per_state_std, max_norm):
    cfg, pop, probs = _setup(per_state_std, ("tanh", "tanh"), n=2, N=128)
    n, N, L = len(probs), 128, pop.L
    rng = np.random.default_rng(8)
    act = np.stack([p[2] for p in probs]); adv = np.stack([O.trpo_normalise_adv(p[3]) for p in probs]).astype(np.float32)
    alpha = np.array([0.1, 0.0], np.float32)
    nlp_old = np.zeros((n, N), np.float32)
    refs = []
    for i, (st, s, a, _) in enumerate(probs):
        th = O.to_torch_state(st, torch.float64)
        e0 = O.trpo_eval(cfg, th["actor"], s, a, adv[i], np.zeros(N), None, th)
        nlp_old[i] = (e0["nlp"].numpy() + rng.normal(size=N) * 0.3).astype(np.float32)       # many rows outside the clip
        g, _, pre, post = O.ppo_actor_grad(cfg, th["actor"], s, a, adv[i], nlp_old[i].astype(np.float64), float(alpha[i]),
                                           0.0, 0.2, max_norm, th)
        ad = st["adam_actor"]                               # the problems start from a non-trivial optimiser state (t = 7)
        new, m, v, _ = O.keras_adam(th["actor"], g, [torch.from_numpy(x.astype(np.float64)) for x in ad["m"]],
                                    [torch.from_numpy(x.astype(np.float64)) for x in ad["v"]], int(ad["t"]), 1e-3)
        refs.append((O.flat(g).numpy(), pre, post, O.flat(new).numpy(), O.flat(m).numpy(), O.flat(v).numpy(),
                     O.flat(th["actor"]).numpy()))
        pop.set_hyper(i, lr_pi=1e-3)
    before_t = pop.t["adam_t"].cpu().numpy().copy()
    grad, stats = pop.ppo_grad(act, adv, nlp_old, alpha, 0.2, max_norm)
    pop.actor_adam(grad)
    grad, stats = grad.cpu().numpy(), stats.cpu().numpy()
    after_t = pop.t["adam_t"].cpu().numpy()
    assert np.array_equal(after_t[:, 2], before_t[:, 2] + 1) and np.array_equal(after_t[:, [0, 1, 3]], before_t[:, [0, 1, 3]])
    for i in range(n):
        g_ref, pre, post, new, m, v, old = refs[i]
        assert rel(grad[i, :L.na], g_ref) < 1e-3
        assert abs(stats[i, 4] - pre) < 1e-3 * pre and abs(stats[i, 5] - post) < 1e-3 * post
        if max_norm is not None:
            assert pre > max_norm and abs(stats[i, 5] - max_norm) < 1e-5 * max_norm
        assert rel(pop.t["actor_m"][i, :L.na].cpu().numpy(), m) < 1e-3
        assert rel(pop.t["actor_v"][i, :L.na].cpu().numpy(), v) < 2e-3
        assert rel(pop.t["actor"][i, :L.na].cpu().numpy() - old.astype(np.float32), new - old) < 2e-3
    pop.close()


def test_ppo_class_interface():
    from sac_expert_b200.sac_eo.actors.init_actor import init_actor
    from sac_expert_b200.sac_eo.algs.model_free.ppo import PPO
    from sac_expert_b200.sac_eo.envs.synthetic import SyntheticEnv
    np.random.seed(0)
    env = SyntheticEnv(7, 2)
    actor = init_actor(env, [32, 32], ["tanh"], 0.01, 1.0, "orthogonal", False, None, actor_per_state_std=True,
                       actor_squash=True)
    cfg = O.NetCfg(S=7, A=2, actor_hidden=(32, 32), critic_hidden=(8, 8), num_models=0, per_state_std=True,
                   actor_acts=("tanh", "tanh"), std_mult=1.0)
    rng = np.random.default_rng(1)
    N = 200                                             # 3 minibatches of 66, ragged tail of 2 dropped (:61-62)
    s = rng.standard_normal((N, 7)).astype(np.float32)
    w0 = actor.get_weights()
    st = {"actor": [torch.from_numpy(w.astype(np.float64)) for w in w0], "s_mean": torch.zeros(7, dtype=torch.float64),
          "s_std": torch.ones(7, dtype=torch.float64)}
    with torch.no_grad():
        mean, ls = O.gaussian_forward(cfg, st["actor"], torch.as_tensor(s, dtype=torch.float64), st)
    a = (mean + torch.exp(ls) * torch.from_numpy(rng.standard_normal((N, 2)))).numpy().astype(np.float32)
    adv = rng.standard_normal(N).astype(np.float32)
    kw = dict(actor_lr=3e-4, actor_update_it=2, actor_nminibatch=3, adv_center=True, adv_scale=True, eps_ppo=0.2,
              max_grad_norm=0.5, adaptlr=True, adapt_factor=0.03, adapt_minthresh=0.0, adapt_maxthresh=1e-6,
              ent_reg=False, ent_targ=0.0, alpha_lr=0.01)
    alg = PPO(actor, kw)
    np.random.seed(11)
    log = alg.update((s, a, adv, None, None, None))
    after_rng = np.random.randint(1 << 30)
    adam = {"m": [torch.zeros_like(t) for t in st["actor"]], "v": [torch.zeros_like(t) for t in st["actor"]], "t": 0}
    np.random.seed(11)
    new, ref = O.ppo_update(cfg, st["actor"], adam, s, a, adv, st, actor_lr=3e-4, actor_update_it=2, actor_nminibatch=3,
                            eps_ppo=0.2, max_grad_norm=0.5)
    assert after_rng == np.random.randint(1 << 30)       # same consumption of the global NumPy RNG
    w1 = actor.get_weights()
    step = np.concatenate([x.ravel() for x in w1]) - np.concatenate([x.ravel() for x in w0])
    assert rel(step, (O.flat(new) - O.flat(st["actor"])).numpy()) < 5e-3
    for k in ("ent", "tv", "kl", "outside_clip", "actor_grad_norm_pre", "actor_grad_norm"):
        assert abs(log[k] - ref[k]) <= 1e-2 * max(abs(ref[k]), 1e-4), (k, log, ref)
    assert alg._t == 6 and log["actor_lr"] == pytest.approx(3e-4)
    assert alg.actor_lr == pytest.approx(3e-4 / 1.03)    # tv above the (tiny) upper threshold: learning rate shrinks (:107-111)
    # optimiser slots persist into the next update, as the Keras optimiser's do
    alg.update((s, a, adv, None, None, None))
    assert alg._t == 12


# ----------------------------------------------------------------------------------------------------------
# expert-observation blend of the on-policy updates (trpo.py:92-158, ppo.py:176-213)
# ----------------------------------------------------------------------------------------------------------
def _setup_expert(per_state_std, acts, num_models, n=2, N=96, E=10, gemm_mode=lib.GEMM_FP32_SIMT, hidden=(64, 48),
                  model_hidden=(64, 64), S=9, A=3, seed=30, act_limit=None, model_acts=("relu", "tanh")):
    cfg = O.NetCfg(S=S, A=A, actor_hidden=hidden, critic_hidden=(16, 16), model_hidden=model_hidden, num_models=num_models,
                   per_state_std=per_state_std, actor_acts=acts, model_acts=model_acts, std_mult=0.7, delta_clip_pred=3.0)
    B = 8
    pop = Population(spec_from_cfg(cfg, n, B, E, 16, fvp_rows=N, gemm_mode=gemm_mode))
    probs = []
    rng = np.random.default_rng(seed)
    noise = np.zeros((n, 3 * B + E, A), np.float32)
    perm = np.zeros((n, E), np.int32)
    for i in range(n):
        st, replay, expert, hyper = O.make_problem(cfg, B, E, 300, seed=seed + i, perturb=0.2)
        if act_limit is not None:
            st["act_limit"] = np.full(A, act_limit, np.float32)
        batch = O.draw_batch(cfg, replay, expert, B, seed=seed + 100 + i)
        pop.load_agent(i, st, hyper)
        pop.set_expert(i, expert["sE"], expert["spE"])
        s = replay["s"][:N]
        th = O.to_torch_state(st, torch.float64)
        with torch.no_grad():
            mean, ls = O.gaussian_forward(cfg, th["actor"], torch.as_tensor(s, dtype=torch.float64), th)
        a = (mean + torch.exp(ls) * torch.from_numpy(rng.standard_normal((N, A)))).numpy().astype(np.float32)
        pop.t["fvp_states"][i].copy_(torch.from_numpy(s))
        if num_models == 2:
            noise[i, 2 * B:2 * B + E] = np.concatenate([batch["u3"], batch["u4"]])
            perm[i] = np.concatenate([batch["I1"], batch["I2"]])
        else:
            noise[i, 2 * B:2 * B + E] = batch["u3"]
            perm[i] = np.arange(E)
        probs.append((st, s, a, rng.standard_normal(N).astype(np.float32) * (1 + i), batch))
    pop.set_draws(noise=noise, perm=perm)
    return cfg, pop, probs


@pytest.mark.parametrize("per_state_std,acts,num_models,mode", [
    (True, ("tanh", "tanh"), 2, lib.GEMM_FP32_SIMT), (False, ("relu", "tanh"), 2, lib.GEMM_FP32_SIMT),
    (True, ("tanh", "relu"), 1, lib.GEMM_FP32_SIMT), (False, ("tanh", "tanh"), 1, lib.GEMM_FP32_SIMT),
    (True, ("tanh", "tanh"), 2, lib.GEMM_TCGEN05_BF16X3), (False, ("tanh", "tanh"), 1, lib.GEMM_TCGEN05_BF16X3)])
def test_onpolicy_expert_gradient_matches_the_oracle(per_state_std, acts, num_models, mode):
    """saceo_onpolicy_expert_grad: d MSE / d(actor trainable) through GaussianActor.sample and the frozen model(s) against
    autograd (oracle.trpo_expert_blend for the two-model TRPO form, oracle.ppo_expert_blend for the clipped one-model PPO
    form; action limit 0.6 so that some counterfactual actions really are clipped)."""
    big = dict(hidden=(256, 256), model_hidden=(512, 512), S=27, A=8, E=20, N=128) if mode == lib.GEMM_TCGEN05_BF16X3 else {}
    cfg, pop, probs = _setup_expert(per_state_std, acts, num_models, gemm_mode=mode,
                                    act_limit=0.6 if num_models == 1 else None, **big)
    L = pop.L
    grad, mse = pop.onpolicy_expert_grad(n_models=num_models, clip_actions=num_models == 1)
    grad, mse = grad.cpu().numpy(), mse.cpu().numpy()
    clipped = 0
    for i, (st, s, a, adv, batch) in enumerate(probs):
        th = O.to_torch_state(st, torch.float64)
        zero = [torch.zeros_like(t) for t in th["actor"]]
        if num_models == 2:
            g_ref, mse_ref, _, _ = O.trpo_expert_blend(cfg, th["actor"], zero, batch, 1.0, th)
        else:
            g_ref, mse_ref = O.ppo_expert_blend(cfg, th["actor"], zero, batch, 1.0, th)
            with torch.no_grad():
                c = O.gaussian_sample(cfg, th["actor"], torch.as_tensor(batch["sE"], dtype=torch.float64), batch["u3"], th)
            clipped += int((c.abs() > 0.6).sum())
        assert rel(grad[i, :L.na], O.flat(g_ref).numpy()) < 1e-3, i
        assert np.all(grad[i, L.na:] == 0)
        assert abs(mse[i] - float(mse_ref)) <= 1e-4 * abs(float(mse_ref))
    if num_models == 1:
        assert clipped > 0
    pop.close()


@pytest.mark.parametrize("per_state_std", [True, False])
def test_grad_blend_norms_and_clip(per_state_std):
    """saceo_grad_blend: (1 - eps) a + eps b with two rounded products and one rounded sum (bit-exact against NumPy fp32),
    the reference's norm bookkeeping (sums of per-tensor L2 norms, trpo.py:160-163) and tf.clip_by_global_norm."""
    cfg, pop, probs = _setup_expert(per_state_std, ("tanh", "tanh"), 2)
    n, L = 2, pop.L
    g = torch.Generator().manual_seed(3)
    a = torch.zeros(n, L.na_stride); b = torch.zeros(n, L.na_stride)
    a[:, :L.na] = torch.randn(n, L.na, generator=g) * 0.1; b[:, :L.na] = torch.randn(n, L.na, generator=g)
    eps = np.array([0.3, 0.0], np.float32)
    for max_norm in (None, 0.5):
        out, stats = pop.grad_blend(a.cuda(), b.cuda(), eps, max_norm)
        out, stats = out.cpu().numpy(), stats.cpu().numpy()
        an, bn = a.numpy(), b.numpy()
        for i in range(n):
            ref = (np.float32(1.0 - float(eps[i])) * an[i]).astype(np.float32) + (eps[i] * bn[i]).astype(np.float32)
            shapes = [w.shape for w in probs[i][0]["actor"]]
            sizes = np.cumsum([0] + [int(np.prod(sh)) for sh in shapes])
            n_pg = sum(np.linalg.norm(an[i, sizes[k]:sizes[k + 1]].astype(np.float64)) for k in range(len(shapes)))
            n_ms = sum(np.linalg.norm(bn[i, sizes[k]:sizes[k + 1]].astype(np.float64)) for k in range(len(shapes)))
            assert abs(stats[i, 6] - n_pg) < 1e-5 * n_pg and abs(stats[i, 7] - n_ms) < 1e-5 * n_ms
            gn = np.linalg.norm(ref[:L.na].astype(np.float64))
            assert abs(stats[i, 4] - gn) < 1e-5 * gn
            if max_norm is None:
                assert np.array_equal(out[i], ref)
                assert abs(stats[i, 5] - gn) < 1e-5 * gn
            else:
                scale = max_norm / max(gn, max_norm)
                assert rel(out[i], ref * scale) < 1e-6 and abs(stats[i, 5] - gn * scale) < 1e-5 * gn
    pop.close()


@pytest.mark.parametrize("per_state_std,eps", [(True, 0.5), (False, 0.2)])
def test_trpo_update_with_expert_blend_matches_the_oracle(per_state_std, eps):
    cfg, pop, probs = _setup_expert(per_state_std, ("tanh", "tanh"), 2, n=3, N=128)
    L = pop.L
    act = np.stack([p[2] for p in probs]); adv = np.stack([p[3] for p in probs])
    before = pop.t["actor"].cpu().numpy().copy()
    logs = pop.trpo_update(act, adv, delta=0.02, cg_iters=5, trust_damp=0.01, kl_maxfactor=1.5, expert_eps=eps)
    after = pop.t["actor"].cpu().numpy()
    for i, (st, s, a, ad, batch) in enumerate(probs):
        th = O.to_torch_state(st, torch.float64)
        new, log, _, _ = O.trpo_update(cfg, th["actor"], s, a, ad, th, delta=0.02, cg_iters=5, trust_damp=0.01,
                                       kl_maxfactor=1.5, expert=batch, eps=eps)
        plain, _, _, _ = O.trpo_update(cfg, th["actor"], s, a, ad, th, delta=0.02, cg_iters=5, trust_damp=0.01, kl_maxfactor=1.5)
        assert abs(logs[i]["adj"] - log["adj"]) < 1e-6, (logs[i], log)
        step_ref = (O.flat(new) - O.flat(th["actor"])).numpy()
        step_plain = (O.flat(plain) - O.flat(th["actor"])).numpy()
        assert rel(step_ref, step_plain) > 1e-2                                    # the expert term really moves the step
        if log["adj"] != 0:
            assert rel(after[i, :L.na] - before[i, :L.na], step_ref) < 1e-2
        for key in ("ent", "tv_pre", "kl_pre", "tv", "kl", "improve"):
            assert abs(logs[i][key] - log[key]) <= 2e-2 * max(abs(log[key]), 1e-3), (key, logs[i], log)
        assert logs[i]["mse"] > 0 and logs[i]["norm_MSE"] > 0 and logs[i]["norm_pg"] > 0
    pop.close()


def test_onpolicy_classes_take_expert_reg():
    """``TRPO.update(rollout, expert_reg)`` / ``PPO.update(rollout, expert_reg)`` through the mirror classes: same RNG
    consumption as the reference (rng.shuffle + two actor.sample draws; one draw per PPO minibatch step), the expert
    weight changes the update, epsilon = 0 reproduces the plain update bit for bit."""
    from sac_expert_b200.sac_eo.actors.init_actor import init_actor
    from sac_expert_b200.sac_eo.algs.model_free.ppo import PPO
    from sac_expert_b200.sac_eo.algs.model_free.trpo import TRPO
    from sac_expert_b200.sac_eo.envs.synthetic import SyntheticEnv
    from sac_expert_b200.sac_eo.models.init_world_models import init_world_models
    env = SyntheticEnv(7, 2)
    rng0 = np.random.default_rng(1)
    N, E = 80, 10
    s = rng0.standard_normal((N, 7)).astype(np.float32); a = rng0.standard_normal((N, 2)).astype(np.float32)
    adv = rng0.standard_normal(N).astype(np.float32)
    sE = rng0.standard_normal((E, 7)).astype(np.float32); spE = (sE + 0.1 * rng0.standard_normal((E, 7))).astype(np.float32)
    tkw = dict(adv_center=True, adv_scale=True, delta_trpo=0.02, cg_it=5, trust_sub=1, trust_damp=0.01, kl_maxfactor=1.5,
               ent_reg=False, ent_targ=0.0, alpha_lr=0.01)
    pkw = dict(adv_center=True, adv_scale=True, eps_ppo=0.2, max_grad_norm=0.5, actor_lr=3e-4, actor_update_it=2,
               actor_nminibatch=2, adaptlr=False, adapt_factor=0.03, adapt_minthresh=0.0, adapt_maxthresh=1.0,
               ent_reg=False, ent_targ=0.0, alpha_lr=0.01)

    def fresh():
        np.random.seed(0)
        actor = init_actor(env, [32, 32], ["tanh"], 0.01, 1.0, "orthogonal", False, None, actor_per_state_std=True,
                           actor_squash=True)
        models = init_world_models(env, [32, 32], ["relu"], 1.0, 1.0, None, [32, 32], ["relu"], 1.0, None, 2, False,
                                   dict(separate_reward_nn=False, reward_loss_coef=1.0, scale_model_loss=False,
                                        delta_clip_loss=None, reward_clip_loss=None, delta_clip_pred=3.0,
                                        reward_clip_pred=None))
        return actor, models

    def flat(actor):
        return np.concatenate([w.ravel() for w in actor.get_weights()])

    for cls, kw in ((TRPO, tkw), (PPO, pkw)):
        res = {}
        for eps in (None, 0.0, 0.6):
            actor, models = fresh()
            w0 = flat(actor)
            alg = cls(actor, kw)
            np.random.seed(5)
            reg = None if eps is None else (sE, None, spE, eps, models, False, np.random.default_rng(9))
            log = alg.update((s, a, adv, None, None, None), expert_reg=reg)
            res[eps] = (flat(actor) - w0, np.random.random(), log)            # the next global draw = how much was consumed
        if cls is TRPO:      # the draws of the expert branch do not feed back into the TRPO step: epsilon = 0 is the plain update
            assert rel(res[0.0][0], res[None][0]) < 1e-6, rel(res[0.0][0], res[None][0])
        # (PPO: the actor.sample draws advance the global stream that also shuffles the minibatches, so epsilon = 0
        # follows other minibatches than the plain update - as in the reference)
        assert rel(res[0.6][0], res[0.0][0]) > 1e-2
        assert np.isfinite(res[0.6][0]).all()
        if cls is TRPO:
            assert res[0.6][2]["epsilon"] == 0.6 and res[0.6][2]["norm_MSE"] > 0
        # the expert branch consumes the global NumPy stream (actor.sample), the plain one does not (TRPO) / less (PPO)
        assert res[0.6][1] != res[None][1]
        assert res[0.6][1] == res[0.0][1]
