"""CPU tests of the ORACLE's internal consistency (it is the checker for the CUDA path; the pins against the reference's
own code are in tests/test_reference_pin.py, see oracle/sac_eo_oracle.py header): fp32 vs fp64 twin, autograd vs the
hand-derived backward, both Fisher-vector forms, the committed golden fixtures, and one test per reference
quirk (SURVEY.md §8a closing list)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import sac_eo_oracle as O
from oracle.analytic import analytic_update, fvp_gn

GOLD = os.path.join(os.path.dirname(__file__), "golden")

CFGS = {
    "saceo2": O.NetCfg(S=11, A=3, actor_hidden=(64, 64), critic_hidden=(64, 64), model_hidden=(96, 96)),
    "one_model_state_indep": O.NetCfg(S=7, A=2, actor_hidden=(32, 48), critic_hidden=(40, 32), model_hidden=(64, 32),
                                      per_state_std=False, num_models=1, actor_acts=("tanh", "tanh"),
                                      critic_acts=("elu", "tanh"), model_acts=("relu", "elu"), delta_clip_pred=0.05),
    "plain": O.NetCfg(S=5, A=2, actor_hidden=(32, 32), critic_hidden=(32, 32), num_models=0),
}


def rel(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def problem(cfg, B=48, E=10, seed=1):
    st, replay, expert, hyper = O.make_problem(cfg, B=B, E=E, N=400, seed=seed, perturb=0.05)
    hyper["eps"] = 0.3
    return st, replay, expert, hyper, O.draw_batch(cfg, replay, expert, B, seed=seed + 1)


@pytest.mark.parametrize("name", list(CFGS))
def test_fp32_vs_fp64_and_analytic(name):
    cfg = CFGS[name]
    st, replay, expert, hyper, batch = problem(cfg)
    o64 = O.sac_eo_update(cfg, O.to_torch_state(st, torch.float64), batch, hyper)
    o32 = O.sac_eo_update(cfg, O.to_torch_state(st, torch.float32), batch, hyper)
    a64 = analytic_update(cfg, st, batch, hyper, dtype=np.float64)
    for k in ("y", "L_q1", "L_q2", "L_pi", "mse", "p_loss", "alpha_loss", "g_alpha"):
        assert rel(a64[k], o64[k].numpy()) < 1e-12, k
        assert rel(o32[k].numpy(), o64[k].numpy()) < 2e-5, k
    for k in ("g_q1", "g_q2", "g_actor"):
        for x, y, z in zip(a64[k], o64[k], o32[k]):
            assert rel(x, y.numpy()) < 1e-11, k
            assert rel(z.numpy(), y.numpy()) < 5e-5, k
    for k in ("q1", "q2", "actor", "t1", "t2"):
        for x, y in zip(a64["new"][k], o64["new"][k]):
            assert rel(x, y.numpy()) < 1e-12, k


@pytest.mark.parametrize("per_state_std", [True, False])
def test_fisher_vector_forms_agree(per_state_std):
    cfg = O.NetCfg(S=6, A=2, actor_hidden=(32, 24), critic_hidden=(8, 8), num_models=0, per_state_std=per_state_std,
                   actor_acts=("tanh", "elu"), std_mult=0.6)
    st, replay, _, _ = O.make_problem(cfg, 8, 2, 200, seed=4, perturb=0.1)
    th = O.to_torch_state(st, torch.float64)
    s_all = replay["s"][:64]
    F1 = O.make_F(cfg, th["actor"], s_all, th, 0.01)
    F2 = O.make_F_gn(cfg, th["actor"], s_all, th, 0.01)
    x = torch.randn(sum(w.numel() for w in th["actor"]), dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    ref = F1(x).numpy()
    assert rel(F2(x).numpy(), ref) < 1e-12
    assert rel(fvp_gn(cfg, st["actor"], x.numpy(), s_all, st, 0.01, dtype=np.float64), ref) < 1e-12
    # F is symmetric positive definite (+damp): x.F(y) == y.F(x), CG reduces the residual, x.F(x) > 0
    y2 = torch.randn_like(x)
    assert abs(float(x.dot(F1(y2)) - y2.dot(F1(x)))) < 1e-9 * float(x.norm() * y2.norm())
    b = torch.randn_like(x) * 0.1
    sol, vFv, step = O.trpo_step(F1, b, delta=0.02, cg_iters=40)
    assert float(vFv) > 0 and rel(F1(sol).numpy(), b.numpy()) < 0.2      # 40 CG iterations: residual well down


def test_golden_update_fixtures():
    pend = np.load(os.path.join(GOLD, "pendulum_templog0.npz"))
    aw = [pend[f"actor_{i}"] for i in range(7)]
    assert [w.shape for w in aw] == [(3, 64), (64,), (64, 64), (64,), (64, 1), (1,), (1, 1)]
    for tag, nm in (("sac", 0), ("saceo", 2)):
        gold = np.load(os.path.join(GOLD, f"golden_pendulum_{tag}.npz"))
        cfg = O.NetCfg(S=3, A=1, actor_hidden=(64, 64), critic_hidden=(64, 64), model_hidden=(64, 64),
                       actor_acts=("tanh", "tanh"), critic_acts=("tanh", "tanh"), per_state_std=False, num_models=nm)
        st, replay, expert, hyper = O.make_problem(cfg, B=32, E=8, N=300, seed=123, perturb=0.02)
        st["actor"] = [w.copy() for w in aw]
        st["adam_actor"] = dict(m=[np.zeros_like(w) for w in aw], v=[np.zeros_like(w) for w in aw], t=0)
        st.update(s_mean=pend["s_mean"], s_std=pend["s_std"], a_mean=pend["a_mean"], a_std=pend["a_std"],
                  ret_std=pend["ret_std"], m_s_mean=pend["s_mean"], m_s_std=pend["s_std"], m_a_mean=pend["a_mean"],
                  m_a_std=pend["a_std"], m_d_mean=pend["d_mean"], m_d_std=pend["d_std"])
        hyper["eps"] = 0.25
        batch = O.draw_batch(cfg, replay, expert, 32, seed=321)
        for dt, tol in ((torch.float64, 1e-12), (torch.float32, 3e-4)):
            o = O.sac_eo_update(cfg, O.to_torch_state(st, dt), batch, hyper)
            for k in ("y", "L_q1", "L_q2", "L_pi", "mse", "p_loss", "alpha_loss", "g_alpha"):
                assert rel(o[k].numpy(), gold[k]) < tol, (tag, k)
            for k in ("g_q1", "g_q2", "g_actor"):
                assert rel(O.flat(o[k]).numpy(), gold[k]) < tol, (tag, k)
            assert rel(O.flat(o["new"]["actor"]).numpy(), gold["new_actor"]) < tol
            assert rel(O.flat(o["new"]["t1"]).numpy(), gold["new_t1"]) < tol


def test_golden_fvp_fixture():
    pend = np.load(os.path.join(GOLD, "pendulum_templog0.npz"))
    gold = np.load(os.path.join(GOLD, "golden_pendulum_fvp.npz"))
    aw = [pend[f"actor_{i}"] for i in range(7)]
    cfg = O.NetCfg(S=3, A=1, actor_hidden=(64, 64), critic_hidden=(8, 8), actor_acts=("tanh", "tanh"),
                   per_state_std=False, num_models=0)
    st = dict(actor=aw, s_mean=pend["s_mean"], s_std=pend["s_std"])
    th = O.to_torch_state(st, torch.float64)
    F = O.make_F(cfg, th["actor"], gold["states"], th, damp=0.01)
    assert rel(F(torch.from_numpy(gold["x"])).numpy(), gold["Fx"]) < 1e-12
    assert rel(fvp_gn(cfg, aw, gold["x"], gold["states"], st, 0.01, dtype=np.float64), gold["Fx"]) < 1e-10
    sol = O.cg(F, torch.from_numpy(gold["b"]), cg_iters=20)
    assert rel(sol.numpy(), gold["cg_x"]) < 1e-9


# ---------------------------------------------------------------------------------- quirks
def test_quirk_keras_adam_epsilon_outside_bias_correction():
    th, g = [torch.tensor([1.0, -2.0])], [torch.tensor([1e-6, 3e-7])]
    new, m, v, t = O.keras_adam(th, g, [torch.zeros(2)], [torch.zeros(2)], 0, 1e-3)
    lr_t = 1e-3 * math.sqrt(1 - 0.999) / (1 - 0.9)
    expect = th[0] - lr_t * (0.1 * g[0]) / (torch.sqrt(0.001 * g[0] ** 2) + 1e-7)
    assert torch.allclose(new[0], expect, rtol=1e-6) and t == 1
    # torch.optim.Adam (eps inside the corrected denominator) gives a different step at this gradient scale
    p = torch.nn.Parameter(th[0].clone()); p.grad = g[0].clone()
    torch.optim.Adam([p], lr=1e-3, eps=1e-7).step()
    assert not torch.allclose(p.detach() - th[0], new[0] - th[0], rtol=1e-2, atol=0)


def test_quirk_raw_alpha_and_mixed_normalisation():
    cfg = CFGS["plain"]
    st, replay, expert, hyper, batch = problem(cfg)
    assert st["alpha"] > 0          # perturbed problem
    st["alpha"] = np.float32(math.log(0.1))       # SAC_expert.py:106 - raw, negative, used without exp
    o = O.sac_eo_update(cfg, O.to_torch_state(st, torch.float64), batch, hyper)
    T = O.to_torch_state(st, torch.float64)
    sp = torch.as_tensor(batch["sp"]).double()
    a1, nlp1 = O.head(cfg, T["actor"], sp, torch.as_tensor(batch["u1"]), T)
    minq = torch.minimum(O.q_value(cfg, T["t1"], sp, a1, T), O.q_value(cfg, T["t2"], sp, a1, T))
    y = torch.as_tensor(batch["r"]).double() + hyper["gamma"] * (1 - torch.as_tensor(batch["d"])) * (minq + math.log(0.1) * nlp1)
    assert rel(o["y"].numpy(), y.numpy()) < 1e-6
    # critic loss: NORMALISED prediction vs DENORMALISED target (ret_std != 1 makes the two differ)
    q = O.q_forward(cfg, T["q1"], torch.as_tensor(batch["s"]).double(), torch.as_tensor(batch["a"]).double(), T)
    assert rel(o["L_q1"].numpy(), (0.5 * (q[:, 0] - o["y"]) ** 2).mean().numpy()) < 1e-12
    assert float(T["ret_std"]) != 1.0
    # alpha is clamped at 1e-5 after its step (SAC_expert.py:348)
    assert float(o["new"]["alpha"]) == pytest.approx(1e-5)


def test_quirk_head_ignores_std_mult_and_clips_logstd():
    cfg = O.NetCfg(S=4, A=2, actor_hidden=(8, 8), critic_hidden=(8, 8), num_models=0, std_mult=0.3)
    st, *_ = O.make_problem(cfg, 8, 2, 50, seed=0, identity_norm=True)
    T = O.to_torch_state(st, torch.float64)
    T["actor"][5] = torch.tensor([0.0, 0.0, 9.0, -9.0], dtype=torch.float64)     # logstd biases far outside [-5, 2]
    x = torch.zeros(3, 4, dtype=torch.float64); u = torch.ones(3, 2, dtype=torch.float64)
    pi, nlp = O.head(cfg, T["actor"], x, u, T)
    z = torch.atanh(pi)
    mean = O.mlp(T["actor"], x, cfg.actor_acts)[:, :2]
    assert torch.allclose(z - mean, torch.tensor([math.exp(2.0), math.exp(-5.0)], dtype=torch.float64).expand(3, 2), rtol=1e-6)
    cfg2 = O.NetCfg(S=4, A=2, actor_hidden=(8, 8), critic_hidden=(8, 8), num_models=0, std_mult=1.0)
    pi2, _ = O.head(cfg2, T["actor"], x, u, T)
    assert torch.equal(pi, pi2)      # std_mult / logstd_init play no role in evaluate()/sample()


def test_quirk_reduce_min_splits_ties_and_clip_passes_gradient_on_the_boundary():
    a = torch.tensor([1.0, 2.0], requires_grad=True); b = torch.tensor([1.0, 3.0], requires_grad=True)
    torch.minimum(a, b).sum().backward()
    assert a.grad.tolist() == [0.5, 1.0] and b.grad.tolist() == [0.5, 0.0]
    x = torch.tensor([-5.0, 2.0, 2.1], requires_grad=True)
    torch.clamp(x, O.MIN_LOG_STD, O.MAX_LOG_STD).sum().backward()
    assert x.grad.tolist() == [1.0, 1.0, 0.0]


def test_quirk_polyak_gate_and_fp32_rounding():
    cfg = CFGS["plain"]
    st, replay, expert, hyper, batch = problem(cfg)
    hyper["do_polyak"] = False
    o = O.sac_eo_update(cfg, O.to_torch_state(st), batch, hyper)
    assert all(torch.equal(a, torch.as_tensor(b)) for a, b in zip(o["new"]["t1"], st["t1"]))
    hyper["do_polyak"] = True
    o = O.sac_eo_update(cfg, O.to_torch_state(st), batch, hyper)
    tau = np.float32(hyper["tau"]); om = np.float32(1.0 - hyper["tau"])
    want = st["t1"][0] * om + o["new"]["q1"][0].numpy() * tau           # NumPy fp32: two products, one sum
    assert np.array_equal(o["new"]["t1"][0].numpy(), want)


def test_quirk_two_models_need_even_expert_rows():
    cfg = CFGS["saceo2"]
    st, replay, expert, hyper = O.make_problem(cfg, B=16, E=7, N=100, seed=0)
    batch = O.draw_batch(cfg, replay, expert, 16, seed=1)
    with pytest.raises(RuntimeError):       # unequal halves cannot be added (SAC_expert.py:329-332)
        O.sac_eo_update(cfg, O.to_torch_state(st), batch, hyper)


def test_adaptive_epsilon():
    assert O.adaptive_epsilon(1e-3) == 1e-3
    e = O.adaptive_epsilon(0.5, scale_by_true_mse=True, mse_cf=4.0)
    assert e == pytest.approx(1 / 3)
    e2 = O.adaptive_epsilon(0.5, scale_by_true_mse=True, mse_cf=4.0, j_cur=50.0, j_exp=100.0, min_mult=True, mult_coeff=1.0)
    assert e2 == pytest.approx((1 / 3) * 0.5)
    e3 = O.adaptive_epsilon(0.5, scale_by_true_mse=True, mse_cf=4.0, j_cur=50.0, j_exp=100.0, exp_mult=True, mult_coeff=2.0)
    assert e3 == pytest.approx((1 / 3) * math.exp(-1.0))
    assert O.adaptive_epsilon(2.0, disc_mode="max", disc=np.array([1.0, 3.0])) == pytest.approx(1 / 7)
    assert O.adaptive_epsilon(2.0, disc_mode="median", disc=np.array([1.0, 3.0, 5.0])) == pytest.approx(1 / 7)
    assert O.adaptive_epsilon(2.0, disc_mode="total", disc=np.array([1.0, 3.0])) == pytest.approx(1 / 9)


def test_gather_semantics_with_replacement():
    rng = np.random.default_rng(0)
    rep = dict(s=rng.standard_normal((9, 3)).astype(np.float32), a=rng.standard_normal((9, 2)).astype(np.float32),
               sp=rng.standard_normal((9, 3)).astype(np.float32), r=rng.standard_normal(9).astype(np.float32),
               d=(rng.random(9) < 0.5).astype(np.float64))
    idx = np.array([8, 8, 0, 3, 3, 3])
    s, a, sp, r, d = O.gather(rep, idx)
    assert s.shape == (6, 3) and d.dtype == np.float64 and np.array_equal(s[0], s[1]) and np.array_equal(r[3:], rep["r"][[3, 3, 3]])


def test_model_fit_restatement():
    """apply_model_grads: joint Adam over both models, clip_by_global_norm semantics, batch index shapes
    (mbrl_onpolicy_alg.py:301-319, SAC_expert.py:524-540)."""
    from oracle.sac_eo_oracle import (NetCfg, apply_model_grads, make_problem, model_fit_batches, model_loss,
                                      to_torch_state)
    cfg = NetCfg(S=5, A=2, model_hidden=(16, 16), model_acts=("tanh", "tanh"))
    st, replay, expert, hyper = make_problem(cfg, 8, 4, 103, seed=4, perturb=0.05)
    bs = model_fit_batches(103, 2, 20, True, np.random.default_rng(0))
    assert len(bs) == 5 and all(b.shape == (2, 20) for b in bs)            # ragged tail batch dropped
    assert not np.array_equal(bs[0][0], bs[0][1])                          # per-model shuffles
    shared = model_fit_batches(100, 2, 20, False, np.random.default_rng(0))
    assert len(shared) == 5 and np.array_equal(shared[0][0], shared[0][1])  # one shuffle tiled over the models
    for dt in (torch.float32, torch.float64):
        T = to_torch_state(st, dt)
        models = [T["m1"], T["m2"]]
        zeros = lambda: [[torch.zeros_like(w) for w in m] for m in models]
        b = [{k: torch.as_tensor(replay[k][bs[0][i]]).to(dt) for k in ("s", "a", "sp", "r")} for i in range(2)]
        free = apply_model_grads(cfg, models, dict(m=zeros(), v=zeros(), t=0), b, T, dict(model_lr=1e-3))
        clip = apply_model_grads(cfg, models, dict(m=zeros(), v=zeros(), t=0), b, T,
                                 dict(model_lr=1e-3, model_max_grad_norm=0.25))
        gn = float(free["gnorm"])
        assert gn > 0.5                                                     # clipping active: norm -> 0.25 * num_models
        got = float(torch.sqrt(sum((g ** 2).sum() for gl in clip["grads"] for g in gl)))
        assert abs(got - 0.5) < 1e-4
        big = apply_model_grads(cfg, models, dict(m=zeros(), v=zeros(), t=0), b, T,
                                dict(model_lr=1e-3, model_max_grad_norm=1e6))
        for a_, b_ in zip(big["grads"][0], free["grads"][0]):               # norm below the clip: unchanged
            assert torch.allclose(a_, b_, rtol=1e-6, atol=0)
        assert free["t"] == 1
        # model k's gradient only depends on model k's loss (sum of independent losses)
        l0 = model_loss(cfg, [w.clone().requires_grad_(True) for w in models[0]], b[0]["s"], b[0]["a"], b[0]["sp"], b[0]["r"], T)
        assert abs(float(l0.detach()) - float(free["losses"][0])) < 1e-5 * abs(float(l0.detach()))
    # fp32 vs fp64 gradients
    T32, T64 = to_torch_state(st), to_torch_state(st, torch.float64)
    outs = []
    for T in (T32, T64):
        models = [T["m1"], T["m2"]]
        zeros = lambda: [[torch.zeros_like(w) for w in m] for m in models]
        b = [{k: torch.as_tensor(replay[k][bs[0][i]]) for k in ("s", "a", "sp", "r")} for i in range(2)]
        outs.append(apply_model_grads(cfg, models, dict(m=zeros(), v=zeros(), t=0), b, T,
                                      dict(model_lr=1e-3, reward_loss_coef=0.5, delta_clip_loss=1.0, reward_clip_loss=1.0)))
    for g32, g64 in zip(outs[0]["grads"][1], outs[1]["grads"][1]):
        assert float((g32.double() - g64).norm() / g64.norm()) < 1e-5


def test_bc_is_the_unit_expert_weight_slice():
    """BC._update_actor (BC.py:309-363) == the actor step of SAC_exp._update_actor_and_alpha with epsilon = 1
    (SAC_expert.py:299-338): same MSE, same gradient, same Adam step; critics / temperature untouched by construction."""
    from oracle.sac_eo_oracle import NetCfg, bc_update, draw_batch, make_problem, sac_eo_update, to_torch_state
    for nm, per_state in ((2, True), (1, False)):
        cfg = NetCfg(S=5, A=2, actor_hidden=(16, 16), critic_hidden=(16, 16), model_hidden=(16, 16), num_models=nm,
                     per_state_std=per_state)
        st, replay, expert, hyper = make_problem(cfg, 16, 4, 100, seed=1, perturb=0.05)
        b = draw_batch(cfg, replay, expert, 16, seed=3)
        T = to_torch_state(st, torch.float64)
        o = bc_update(cfg, T, b, hyper)
        h = dict(hyper); h["eps"] = 1.0
        o2 = sac_eo_update(cfg, T, b, h)
        assert abs(float(o["mse"]) - float(o2["mse"])) < 1e-12
        for g1, g2 in zip(o["g_actor"], o2["g_actor"]):
            assert float((g1 - g2).abs().max()) < 1e-12
        for w1, w2 in zip(o["new"]["actor"], o2["new"]["actor"]):
            assert float((w1 - w2).abs().max()) < 1e-14
        assert o["new"]["adam_actor"]["t"] == st["adam_actor"]["t"] + 1


def test_gaussian_model_loss_gradients():
    """GaussianModel.get_loss (continuous_models.py:101-131): closed-form gradient w.r.t. the unclipped logstd and the
    stop-gradient of scale_model_loss."""
    from oracle.sac_eo_oracle import NetCfg, make_problem, model_loss, to_torch_state
    cfg = NetCfg(S=4, A=2, model_hidden=(8, 8), model_acts=("tanh", "tanh"))
    st, replay, _, _ = make_problem(cfg, 8, 4, 50, seed=2, perturb=0.05)
    T = to_torch_state(st, torch.float64)
    th = [w.clone().requires_grad_(True) for w in T["m1"]]
    ls = (0.3 * torch.randn(1, 4, dtype=torch.float64)).requires_grad_(True)
    b = {k: torch.as_tensor(replay[k][:20]).double() for k in ("s", "a", "sp", "r")}
    for scale in (False, True):
        L = model_loss(cfg, th, b["s"], b["a"], b["sp"], b["r"], T, 0.5, 0.0, 0.0, ls, scale)
        (g,) = torch.autograd.grad(L, ls)
        with torch.no_grad():
            from oracle.sac_eo_oracle import mlp, normalize
            sa = torch.cat([normalize(b["s"], T["m_s_mean"], T["m_s_std"]), normalize(b["a"], T["m_a_mean"], T["m_a_std"])], -1)
            e = normalize(b["sp"] - b["s"], T["m_d_mean"], T["m_d_std"]) - mlp(th, sa, cfg.model_acts)[:, :-1]
            sc = (torch.exp(ls) ** 2).mean() if scale else 1.0
            ref = sc * (1.0 - e ** 2 * torch.exp(-2 * ls)).mean(0, keepdim=True)
        assert float((g - ref).abs().max()) < 1e-12


# ----------------------------------------------------------------------------------------------------------
# TRPO surrogate / line search restatement (trpo.py:36-198, :229-317)
# ----------------------------------------------------------------------------------------------------------
def _trpo_problem(per_state_std, acts=("tanh", "tanh"), N=64, seed=3, std_mult=0.8):
    cfg = O.NetCfg(S=6, A=3, actor_hidden=(32, 24), critic_hidden=(8, 8), num_models=0, per_state_std=per_state_std,
                   actor_acts=acts, std_mult=std_mult)
    st, replay, _, _ = O.make_problem(cfg, 8, 2, 200, seed=seed, perturb=0.2)
    rng = np.random.default_rng(seed)
    return cfg, st, replay["s"][:N], replay["a"][:N], rng.standard_normal(N).astype(np.float32)


@pytest.mark.parametrize("per_state_std", [True, False])
def test_trpo_surrogate_gradient_matches_the_closed_form_output_cotangent(per_state_std):
    """The CUDA head (csrc/trpo.cuh::k_trpo_rows) uses d/dmean = -adv ratio q / std / N and
    d/dlogstd = (adv ratio (1 - q^2) - alpha) / N; check them against autograd of the restated tape in fp64."""
    cfg, st, s, a, adv = _trpo_problem(per_state_std)
    th = O.to_torch_state(st, torch.float64)
    theta = th["actor"]
    N, A = len(s), cfg.A
    adv_n = O.trpo_normalise_adv(adv)
    rng = np.random.default_rng(0)
    with torch.no_grad():
        m0, l0 = O.gaussian_forward(cfg, theta, torch.as_tensor(s, dtype=torch.float64), th)
        nlp_old = O.gaussian_neglogp(m0, l0, torch.as_tensor(a, dtype=torch.float64)) + torch.as_tensor(rng.normal(size=N) * 0.1)
    alpha = 0.37
    neg_pg, alpha_grad, _ = O.trpo_surrogate_grad(cfg, theta, s, a, adv_n, nlp_old, alpha, -1.5, th)
    # closed form through the (mean, logstd) outputs
    tt = [t.clone().requires_grad_(True) for t in theta]
    mean, ls = O.gaussian_forward(cfg, tt, torch.as_tensor(s, dtype=torch.float64), th)
    with torch.no_grad():
        sd = torch.exp(ls)
        q = (torch.as_tensor(a, dtype=torch.float64) - mean) / sd
        ratio = torch.exp(nlp_old - O.gaussian_neglogp(mean, ls, torch.as_tensor(a, dtype=torch.float64)))
        w = (torch.as_tensor(adv_n, dtype=torch.float64) * ratio / N)[:, None]
        g_mean = -w * q / sd
        g_ls = w * (1 - q * q) - alpha / N
    res = torch.autograd.grad([mean, ls], tt, grad_outputs=[g_mean, g_ls * torch.ones_like(ls)], allow_unused=True)
    res = [r if r is not None else torch.zeros_like(p) for r, p in zip(res, tt)]
    assert rel(O.flat(res), O.flat(neg_pg)) < 1e-12
    ent = float(O.gaussian_entropy(ls.detach()).mean())
    assert abs(float(alpha_grad) + (ent - (-1.5))) < 1e-12


def test_trpo_update_line_search_semantics():
    cfg, st, s, a, adv = _trpo_problem(True, ("relu", "tanh"))
    th64 = O.to_torch_state(st, torch.float64)
    new, log, pg, eta_v = O.trpo_update(cfg, th64["actor"], s, a, adv, th64, delta=0.02, cg_iters=5)
    assert log["adj"] == 0 or min(abs(log["adj"] - 2 ** (-k / 2)) for k in range(11)) < 1e-12
    assert log["kl"] <= 1.5 * 0.02 + 1e-12 and log["improve"] >= 0
    # the accepted step is adj * eta * v
    assert rel(O.flat(new) - O.flat(th64["actor"]), eta_v) < 1e-12
    # a trust region no step can satisfy: ten shrinks, then the parameters are restored and adj = 0 (:292-301)
    new0, log0, _, _ = O.trpo_update(cfg, th64["actor"], s, a, adv, th64, delta=0.02, cg_iters=5, kl_maxfactor=1e-9)
    assert log0["adj"] == 0 and abs(log0["kl"]) < 1e-15 and rel(O.flat(new0), O.flat(th64["actor"])) == 0
    # fp32 twin of the same update agrees with fp64
    th32 = O.to_torch_state(st, torch.float32)
    new32, log32, _, _ = O.trpo_update(cfg, th32["actor"], s, a, adv, th32, delta=0.02, cg_iters=5)
    assert log32["adj"] == log["adj"]
    assert rel(O.flat(new32) - O.flat(th32["actor"]), O.flat(new) - O.flat(th64["actor"])) < 1e-2


def test_trpo_state_independent_logstd_floor_on_increment():
    cfg, st, s, a, adv = _trpo_problem(False)
    th = O.to_torch_state(st, torch.float64)["actor"]
    step = torch.zeros(sum(t.numel() for t in th), dtype=torch.float64)
    step[-cfg.A:] = -100.0
    new = O.actor_increment(cfg, th, step)
    assert torch.all(new[-1] == math.log(1e-3)) and rel(O.flat(new[:-1]), O.flat(th[:-1])) == 0


# ----------------------------------------------------------------------------------------------------------
# PPO restatement (ppo.py:41-119, :122-237)
# ----------------------------------------------------------------------------------------------------------
def test_ppo_gradient_is_the_surrogate_gradient_with_clipped_rows_masked():
    """csrc/trpo.cuh::k_trpo_rows (clip_eps >= 0) zeroes the row weight where the clipped branch is selected."""
    cfg, st, s, a, adv = _trpo_problem(True)
    th = O.to_torch_state(st, torch.float64)
    theta = th["actor"]
    N = len(s)
    adv_n = O.trpo_normalise_adv(adv)
    rng = np.random.default_rng(4)
    with torch.no_grad():
        m0, l0 = O.gaussian_forward(cfg, theta, torch.as_tensor(s, dtype=torch.float64), th)
        cur = O.gaussian_neglogp(m0, l0, torch.as_tensor(a, dtype=torch.float64))
        nlp_old = cur + torch.as_tensor(rng.normal(size=N) * 0.3)
        ratio = torch.exp(nlp_old - cur)
        rc = torch.clamp(ratio, 0.8, 1.2)
        advt = torch.as_tensor(adv_n, dtype=torch.float64)
        keep = (-ratio * advt >= -rc * advt)
    assert 0 < int(keep.sum()) < N                       # both branches occur
    g, _, pre, post = O.ppo_actor_grad(cfg, theta, s, a, adv_n, nlp_old, 0.2, 0.0, 0.2, None, th)
    g_ref, _, _ = O.trpo_surrogate_grad(cfg, theta, s, a, adv_n * keep.numpy(), nlp_old, 0.2, 0.0, th)
    assert rel(O.flat(g), O.flat(g_ref)) < 1e-12 and pre == post
    gc, _, pre2, post2 = O.ppo_actor_grad(cfg, theta, s, a, adv_n, nlp_old, 0.2, 0.0, 0.2, 0.5 * pre, th)
    assert abs(pre2 - pre) < 1e-12 and abs(post2 - 0.5 * pre) < 1e-12 and rel(O.flat(gc), 0.5 * O.flat(g)) < 1e-12


def test_ppo_update_fp32_vs_fp64_and_rng_consumption():
    cfg, st, s, a, adv = _trpo_problem(False, N=96)
    out = {}
    for dt in (torch.float64, torch.float32):
        th = O.to_torch_state(st, dt)
        adam = {"m": [torch.zeros_like(t) for t in th["actor"]], "v": [torch.zeros_like(t) for t in th["actor"]], "t": 0}
        np.random.seed(7)
        new, log = O.ppo_update(cfg, th["actor"], adam, s, a, adv, th, actor_update_it=2, actor_nminibatch=3)
        out[dt] = (O.flat(new) - O.flat(th["actor"]), log, adam["t"], np.random.randint(1 << 30))
    assert out[torch.float64][2] == 6 and out[torch.float64][3] == out[torch.float32][3]
    assert rel(out[torch.float32][0], out[torch.float64][0]) < 2e-3
    for k, v in out[torch.float64][1].items():
        assert abs(out[torch.float32][1][k] - v) <= 1e-3 * max(abs(v), 1e-3), k


# ----------------------------------------------------------------------------------------------------------
# Pinned against OUTPUTS OF THE REFERENCE: the TRPO log the reference recorded in sac_eo/logs/TEMPLOG_0
# (tests/golden/templog0_trpo_log.json, extracted by tests/golden/make_golden_trpo_log.py)
# ----------------------------------------------------------------------------------------------------------
def _ref_trpo_log():
    import json
    return json.load(open(os.path.join(GOLD, "templog0_trpo_log.json")))


def test_reference_log_pins_the_gaussian_entropy_and_logstd_parameterisation():
    """Update 0 of the reference run starts from the freshly initialised state-independent-std actor (logstd variable
    0, continuous_actors.py:56-57; logstd_init = log(std_mult) = 0): its logged mean entropy must be what the restated
    ``gaussian_forward`` + ``gaussian_entropy`` give for those weights, and the final logstd the run saved must be
    consistent with the entropy / KL it logged on the way."""
    ref = _ref_trpo_log()
    A = ref["actor"]["a_dim"]
    assert not ref["actor"]["per_state_std"] and A == 1
    cfg = O.NetCfg(S=3, A=A, actor_hidden=(64, 64), critic_hidden=(8, 8), num_models=0, per_state_std=False,
                   actor_acts=("tanh", "tanh"), std_mult=ref["actor"]["std_mult"])
    gold = np.load(os.path.join(GOLD, "pendulum_templog0.npz"))
    theta = [torch.from_numpy(gold[f"actor_{i}"]) for i in range(7)]
    st = {"s_mean": torch.from_numpy(gold["s_mean"]), "s_std": torch.from_numpy(gold["s_std"])}
    s = torch.from_numpy(np.random.default_rng(0).standard_normal((50, 3)).astype(np.float32))
    theta0 = theta[:6] + [torch.zeros(1, A)]                                   # the initial logstd variable
    _, ls0 = O.gaussian_forward(cfg, theta0, s, st)
    assert abs(float(O.gaussian_entropy(ls0).mean()) - ref["ent"][0]) < 5e-7   # reference: 1.4189382 (fp32)
    # entropy is affine in the logstd variable (A = 1): the logged entropies give the variable before update 1
    ls_before_last = ref["ent"][1] - ref["ent"][0]
    _, ls_fin = O.gaussian_forward(cfg, theta, s, st)
    assert abs(float(ls_fin[0, 0]) - ref["actor"]["final_logstd"][0]) < 1e-7
    # the last update moved the logstd variable from ls_before_last to final_logstd; the KL it logged for that step
    # (forward KL to the old policy) can be no smaller than its logstd-only part (mean shift = 0)
    ref_info_ls = torch.full((1, A), ls_before_last)
    kl_floor = float(O.kl_forward(torch.zeros(1, A), ls_fin[:1], torch.zeros(1, A), ref_info_ls))
    assert 0 < kl_floor < ref["kl"][1]
    # and the accepted step is the shrunk one: the full step would have moved logstd 1/adj times as far
    assert ref["adj"][1] < 1 and ref["actor"]["final_logstd"][0] < ls_before_last < 0


def test_reference_log_pins_the_line_search_rule():
    """Replays the reference's own (kl, improve) trajectory through the restated control flow: it must take the same
    decisions - accept at once in update 0 (adj 1, tv/kl == tv_pre/kl_pre), shrink exactly once by sqrt(2) in update 1
    (kl_pre 0.0553 > kl_maxfactor * delta = 0.03, then 0.0203 <= 0.03) - and report the same adj to the last bit."""
    ref = _ref_trpo_log()
    hp = ref["hyper"]
    for u in range(2):
        seq = [dict(kl=ref["kl_pre"][u], tv=ref["tv_pre"][u])]
        if ref["adj"][u] != 1.0:
            seq.append(dict(kl=ref["kl"][u], tv=ref["tv"][u]))
        calls = []

        def trial(step):
            calls.append(float(step))
            e = seq[min(len(calls) - 1, len(seq) - 1)]
            return None, e, ref["improve"][u]

        _, e, improve, adj, step, tv_pre, kl_pre = O.backtrack(trial, 1.0, hp["kl_maxfactor"], hp["delta_trpo"])
        assert adj == ref["adj"][u]                                              # bit-identical (1.0, 0.7071067811865475)
        assert len(calls) == len(seq) and step == calls[-1] == ref["adj"][u]
        assert (tv_pre, kl_pre) == (ref["tv_pre"][u], ref["kl_pre"][u]) and (e["tv"], e["kl"]) == (ref["tv"][u], ref["kl"][u])
        assert e["kl"] <= hp["kl_maxfactor"] * hp["delta_trpo"] and improve >= 0
    assert ref["kl_pre"][1] > hp["kl_maxfactor"] * hp["delta_trpo"]             # why the reference shrank
    assert all(a == 0.0 for a in ref["alpha"]) and not hp["ent_reg"]             # temperature untouched without ent_reg


# ----------------------------------------------------------------------------------------------------------
# TRPO two-model expert blend (trpo.py:113-165) - checker for the device path of a later round
# ----------------------------------------------------------------------------------------------------------
def _blend_problem(per_state_std):
    cfg = O.NetCfg(S=6, A=3, actor_hidden=(32, 24), critic_hidden=(8, 8), model_hidden=(40, 40), num_models=2,
                   per_state_std=per_state_std, actor_acts=("tanh", "tanh"), std_mult=0.8)
    st, replay, expert, _ = O.make_problem(cfg, 8, 10, 200, seed=5, perturb=0.1)
    batch = O.draw_batch(cfg, replay, expert, 8, seed=6)
    rng = np.random.default_rng(5)
    return cfg, st, replay["s"][:64], replay["a"][:64], rng.standard_normal(64).astype(np.float32), batch


@pytest.mark.parametrize("per_state_std", [True, False])
def test_trpo_expert_blend_gradient_and_norm_bookkeeping(per_state_std):
    cfg, st, s, a, adv, batch = _blend_problem(per_state_std)
    th = O.to_torch_state(st, torch.float64)
    theta = th["actor"]
    neg_pg = [torch.randn_like(t) * 0.1 for t in theta]
    g0, mse, n_pg, n_mse = O.trpo_expert_blend(cfg, theta, neg_pg, batch, 0.0, th)
    g1, _, _, _ = O.trpo_expert_blend(cfg, theta, neg_pg, batch, 1.0, th)
    gh, _, _, _ = O.trpo_expert_blend(cfg, theta, neg_pg, batch, 0.3, th)
    assert rel(O.flat(g0), O.flat(neg_pg)) == 0                                   # epsilon = 0: the plain gradient
    assert rel(O.flat(gh), 0.7 * O.flat(neg_pg) + 0.3 * O.flat(g1)) < 1e-14       # linear in epsilon
    assert abs(n_pg - sum(float(t.norm()) for t in neg_pg)) < 1e-12               # sum of per-tensor norms, not the global norm
    assert abs(n_mse - sum(float(t.norm()) for t in g1)) < 1e-12 and float(mse) > 0
    # closed form of the unsquashed sample head: dL/dmean = dL/da, dL/dlogstd = dL/da * std * u (floor mask on logstd)
    tt = [t.clone().requires_grad_(True) for t in theta]
    dt = torch.float64
    sE, spE = torch.as_tensor(batch["sE"]).to(dt), torch.as_tensor(batch["spE"]).to(dt)
    res = [torch.zeros_like(t) for t in theta]
    for I, u, mk in ((batch["I1"], batch["u3"], "m1"), (batch["I2"], batch["u4"], "m2")):
        I = torch.as_tensor(np.asarray(I), dtype=torch.long)
        mean, ls = O.gaussian_forward(cfg, tt, sE[I], th)
        act = (mean + torch.exp(ls) * torch.as_tensor(u).to(dt)).detach().requires_grad_(True)
        pred = O.model_sample(cfg, th[mk], sE[I], act, th)
        loss = (0.5 * ((spE[I] - pred) ** 2).sum(-1)).sum() / len(I)
        (dLda,) = torch.autograd.grad(loss, act)
        g_ls = dLda * torch.exp(ls.detach()) * torch.as_tensor(u).to(dt)
        part = torch.autograd.grad([mean, ls], tt, grad_outputs=[dLda, g_ls], allow_unused=True)
        res = [r + (p if p is not None else 0) for r, p in zip(res, part)]
    assert rel(O.flat(res), O.flat(g1)) < 1e-12


def test_trpo_update_with_expert_blend_runs_the_line_search():
    cfg, st, s, a, adv, batch = _blend_problem(True)
    th = O.to_torch_state(st, torch.float64)
    plain, log_p, pg_p, _ = O.trpo_update(cfg, th["actor"], s, a, adv, th, delta=0.02, cg_iters=5)
    same, log_s, pg_s, _ = O.trpo_update(cfg, th["actor"], s, a, adv, th, delta=0.02, cg_iters=5, expert=batch, eps=0.0)
    assert rel(pg_s, pg_p) == 0 and log_s == log_p
    mixed, log_m, pg_m, _ = O.trpo_update(cfg, th["actor"], s, a, adv, th, delta=0.02, cg_iters=5, expert=batch, eps=0.5)
    assert rel(pg_m, pg_p) > 1e-3 and log_m["kl"] <= 1.5 * 0.02 + 1e-12
    # odd expert count: the two halves differ in length and the row-wise sum of their losses (:147-150) cannot be formed
    odd = dict(batch); odd["I2"] = batch["I2"][:-1]; odd["u4"] = batch["u4"][:-1]
    with pytest.raises(RuntimeError):
        O.trpo_expert_blend(cfg, th["actor"], [torch.zeros_like(t) for t in th["actor"]], odd, 0.5, th)
