"""Pins against the REFERENCE'S OWN CODE: tests/golden/ref_*.npz were produced by the reference's unmodified
``SAC_exp._update`` / ``SAC._update`` (SAC_expert.py:463-477, SAC.py:236-250) executed over oracle/tfemu
(tests/golden/make_golden_reference.py; oracle/tfemu/README.md says what that does and does not pin).

CPU tier: the oracle restatement replays the recorded draws and must reproduce the reference's TD target, the gradients
handed to every optimiser, the logged losses and every parameter / target / alpha value after each of K consecutive
updates (fp32 rounding only), with the Polyak gate alternating.  GPU tier: the CUDA path through the C ABI against the
same vectors.  Nothing here reads /root/reference."""
import os

import numpy as np
import pytest
import torch

from oracle.sac_eo_oracle import NetCfg, gather, make_problem, sac_eo_update, to_torch_state

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# must equal CASES of tests/golden/make_golden_reference.py (the meta / checksum entries of each file verify it)
CFGS = dict(
    saceo2_relu=NetCfg(S=5, A=2, actor_hidden=(32, 24), critic_hidden=(32, 24), model_hidden=(32, 24), num_models=2),
    saceo1_tanh_elu_sis=NetCfg(S=6, A=3, actor_hidden=(24, 32), critic_hidden=(32, 32), model_hidden=(24, 24),
                               actor_acts=("tanh", "tanh"), critic_acts=("elu", "elu"), model_acts=("tanh", "tanh"),
                               per_state_std=False, num_models=1, delta_clip_pred=0.05),
    sac_plain_relu=NetCfg(S=5, A=2, actor_hidden=(32, 24), critic_hidden=(32, 24), model_hidden=(8, 8), num_models=0),
    saceo2_hopper_256=NetCfg(S=11, A=3),
    saceo2_sepreward=NetCfg(S=5, A=2, actor_hidden=(32, 24), critic_hidden=(32, 24), model_hidden=(32, 24), num_models=2,
                            separate_reward_nn=True, model_acts=("elu", "tanh")),
)
NETS = ("actor", "q1", "q2", "t1", "t2")


def rel(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def flat(ws):
    return np.concatenate([np.asarray(w.numpy() if hasattr(w, "numpy") else w, np.float32).ravel() for w in ws])


def load_case(name):
    """-> (cfg, golden, problem pieces, meta).  Small cases take their inputs from the file; the benchmark-shaped case
    regenerates them from the seed and both verify the stored checksum."""
    import hashlib
    g = np.load(os.path.join(GOLD, f"ref_{name}.npz"))
    cfg = CFGS[name]
    S, A, B, E, N, seed, K, tui, nm = (int(x) for x in g["meta"])
    assert (S, A, nm) == (cfg.S, cfg.A, cfg.num_models)
    st, replay, expert, hyper = make_problem(cfg, B, E, N, seed=seed, perturb=0.05)
    hyper["eps"] = float(g["eps"])
    if "in_actor_0" in g.files:
        for k in ("actor", "q1", "q2", "t1", "t2", "m1", "m2"):
            st[k] = [g[f"in_{k}_{i}"] for i in range(len(st[k]))]
        for k in ("q1", "q2", "actor"):
            st["adam_" + k]["m"] = [g[f"in_adam_{k}_m_{i}"] for i in range(len(st[k]))]
            st["adam_" + k]["v"] = [g[f"in_adam_{k}_v_{i}"] for i in range(len(st[k]))]
        for k in ("s_mean", "s_std", "a_mean", "a_std", "ret_std", "m_s_mean", "m_s_std", "m_a_mean", "m_a_std", "m_d_mean",
                  "m_d_std", "alpha"):
            st[k] = g["in_" + k] if g["in_" + k].ndim else g["in_" + k][()]
        replay = {k: g["in_replay_" + k] for k in replay}
        expert = {k: g["in_expert_" + k] for k in expert}
    h = hashlib.sha256()
    for k in ("actor", "q1", "q2", "t1", "t2", "m1", "m2"):
        h.update(np.concatenate([np.asarray(w, np.float32).ravel() for w in st[k]]).tobytes())
    for k in ("s", "a", "sp", "r", "d"):
        h.update(np.ascontiguousarray(replay[k]).tobytes())
    assert bytes(g["input_sha256"]) == h.digest(), "inputs differ from those the reference was run on"
    return cfg, g, (st, replay, expert, hyper), dict(B=B, E=E, N=N, K=K, tui=tui, nm=nm)


def step_batch(g, step, replay, expert, m):
    """The draws the reference made in update ``step``, in the oracle's batch layout (SURVEY.md App. A order)."""
    idx = g[f"step{step}_idx"]
    s, a, sp, r, d = gather(replay, idx)
    nrm = [g[f"step{step}_normal{j}"].astype(np.float32) for j in range(3 + m["nm"])]
    b = dict(idx=idx, s=s, a=a, sp=sp, r=r, d=d, u1=nrm[0], u2=nrm[1], u5=nrm[-1])
    if m["nm"] == 2:
        b["I1"], b["I2"] = np.array_split(g[f"step{step}_perm"], 2)        # SAC_expert.py:301-309
        b["u3"], b["u4"] = nrm[2], nrm[3]
    elif m["nm"] == 1:
        b["I1"], b["u3"] = np.arange(m["E"]), nrm[2]
    if m["nm"]:
        b["sE"], b["spE"] = expert["sE"], expert["spE"]
    return b


def proj(seed, dim, vec):
    return np.random.default_rng(seed).standard_normal((dim, vec.size)) @ vec.astype(np.float64)


# ---------------------------------------------------------------------------------------------------------------------
# CPU tier: the oracle restatement against the reference's own code
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(CFGS))
def test_oracle_reproduces_reference_updates(name):
    cfg, g, (st, replay, expert, hyper), m = load_case(name)
    state = to_torch_state(st, torch.float32)
    small = "step0_theta_actor" in g.files
    for step in range(m["K"]):
        hyper["do_polyak"] = (step % m["tui"] == 0)                       # SAC_expert.py:475
        o = sac_eo_update(cfg, state, step_batch(g, step, replay, expert, m), hyper)
        assert rel(o["y"].numpy(), g[f"step{step}_y"]) < 2e-6
        assert abs(float(o["g_alpha"]) - float(g[f"step{step}_g_alpha"])) < 2e-6 * max(1.0, abs(float(o["g_alpha"])))
        if m["nm"]:
            assert abs(float(o["p_loss"]) - float(g[f"step{step}_p_loss"])) < 2e-6 * max(1.0, abs(float(o["p_loss"])))
            assert abs(float(o["alpha_loss"]) - float(g[f"step{step}_alpha_loss"])) < 2e-6
        assert abs(float(o["new"]["alpha"]) - float(g[f"step{step}_alpha"])) < 1e-7
        for k in ("q1", "q2", "actor"):
            got = flat(o["g_" + k])
            if small:
                assert rel(got, g[f"step{step}_g_{k}"]) < 5e-6, (k, step)
            else:
                assert abs(np.linalg.norm(got.astype(np.float64)) / float(g[f"step{step}_gnorm_{k}"]) - 1) < 5e-6
                want = g[f"step{step}_gproj_{k}"]
                assert rel(proj(54321, len(want), got), want) < 2e-5, (k, step)
        for k in NETS:
            got = flat(o["new"][k])
            if small:
                assert rel(got, g[f"step{step}_theta_{k}"]) < 1e-6, (k, step)
            else:
                d = got.astype(np.float64) - flat(st[k]).astype(np.float64)
                want = g[f"step{step}_proj_{k}"]
                assert rel(proj(12345, len(want), d), want) < 1e-3, (k, step)   # Δθ: Adam amplifies last-bit g noise
                assert abs(np.linalg.norm(d) / float(g[f"step{step}_dnorm_{k}"]) - 1) < 1e-4
        for k in NETS + ("alpha", "adam_q1", "adam_q2", "adam_actor", "adam_alpha"):
            state[k] = o["new"][k]
    if small:
        for k in ("q1", "actor"):
            assert rel(flat(state["adam_" + k]["m"]), g[f"final_adam_{k}_m"]) < 5e-6
            assert rel(flat(state["adam_" + k]["v"]), g[f"final_adam_{k}_v"]) < 5e-6


def test_polyak_gate_is_exercised():
    """target_update_int = 2 in the relu cases: the reference left the targets untouched in update 1 and moved them in
    updates 0 and 2."""
    g = np.load(os.path.join(GOLD, "ref_saceo2_relu.npz"))
    assert np.array_equal(g["step0_theta_t1"], g["step1_theta_t1"])
    assert not np.array_equal(g["step1_theta_t1"], g["step2_theta_t1"])
    assert not np.array_equal(g["in_t1_0"].ravel(), g["step0_theta_t1"][:g["in_t1_0"].size])


def test_rng_consumption_recorded():
    """The generator asserted the call order randint, normal x (3 + num_models) per update; here: the
    shapes of what was drawn, and that the expert permutation is a permutation drawn from alg.rng."""
    for name, cfg in CFGS.items():
        g = np.load(os.path.join(GOLD, f"ref_{name}.npz"))
        S, A, B, E, N, seed, K, tui, nm = (int(x) for x in g["meta"])
        assert g["step0_idx"].shape == (B,) and g["step0_idx"].max() < N
        shapes = [g[f"step0_normal{j}"].shape for j in range(3 + nm)]
        if nm == 2:
            assert shapes == [(B, A), (B, A), (E // 2, A), (E - E // 2, A), (B, A)]
            assert sorted(g["step0_perm"].tolist()) == list(range(E))
        elif nm == 1:
            assert shapes == [(B, A), (B, A), (E, A), (B, A)]
        else:
            assert shapes == [(B, A), (B, A), (B, A)]
        assert g["step0_normal0"].dtype == np.float64


# ---------------------------------------------------------------------------------------------------------------------
# GPU tier: the CUDA path against the reference's own outputs
# ---------------------------------------------------------------------------------------------------------------------
def _device_run(name, gemm_mode, use_graph, tol_g, tol_theta, tol_dtheta):
    from sac_expert_b200 import lib as _lib
    from tests.helpers import inject, spec_from_cfg
    from sac_expert_b200.population import Population
    cfg, g, (st, replay, expert, hyper), m = load_case(name)
    mode = getattr(_lib, gemm_mode)
    pop = Population(spec_from_cfg(cfg, 1, m["B"], m["E"], m["N"], gemm_mode=mode, use_graph=use_graph,
                                   target_update_int=m["tui"]))
    pop.load_agent(0, st, hyper)
    pop.append_rows(0, replay["s"], replay["a"], replay["r"], replay["sp"], replay["d"])
    if cfg.num_models > 0:
        pop.set_expert(0, expert["sE"], expert["spE"])
    L = pop.L
    small = "step0_theta_actor" in g.files
    worst = {}
    for step in range(m["K"]):
        inject(pop, cfg, [(st, replay, expert, hyper, step_batch(g, step, replay, expert, m))])
        pop.update(1, num_timesteps=step, use_device_rng=False)
        torch.cuda.synchronize()
        losses = pop.losses.cpu().numpy()[0]
        y = pop.debug("y").cpu().numpy().reshape(-1)[:m["B"]]
        assert rel(y, g[f"step{step}_y"]) < tol_g, ("y", step)
        if m["nm"]:
            assert abs(losses[4] - float(g[f"step{step}_p_loss"])) < tol_g * max(1.0, abs(float(g[f"step{step}_p_loss"])))
            assert abs(losses[5] - float(g[f"step{step}_alpha_loss"])) < tol_g * max(1.0, abs(float(g[f"step{step}_alpha_loss"])))
        assert abs(losses[6] - float(g[f"step{step}_alpha"])) < 1e-6
        g_q = pop.debug("g_q").cpu().numpy().reshape(2, L.nc_stride)
        g_a = pop.debug("g_actor").cpu().numpy().reshape(L.na_stride)
        for k, got_full in (("q1", g_q[0]), ("q2", g_q[1]), ("actor", g_a)):
            if small:
                want = g[f"step{step}_g_{k}"]
                e = rel(got_full[:want.size], want)
            else:
                want = g[f"step{step}_gproj_{k}"]
                n = sum(int(np.prod(w.shape)) for w in st[k])
                e = rel(proj(54321, len(want), got_full[:n]), want)
            worst["g_" + k] = max(worst.get("g_" + k, 0.0), e)
            assert e < tol_g, (k, step, e)
        for k in NETS:
            got = flat(pop.get_net(0, k))
            if small:
                e = rel(got, g[f"step{step}_theta_{k}"])
                assert e < tol_theta, (k, step, e)
                d0 = flat(st[k]).astype(np.float64)
                dref = g[f"step{step}_theta_{k}"].astype(np.float64) - d0
                e = rel(got.astype(np.float64) - d0, dref)
            else:
                d = got.astype(np.float64) - flat(st[k]).astype(np.float64)
                want = g[f"step{step}_proj_{k}"]
                e = rel(proj(12345, len(want), d), want)
            worst["dtheta_" + k] = max(worst.get("dtheta_" + k, 0.0), e)
            assert e < tol_dtheta, (k, step, e)
    return worst


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["saceo2_relu", "saceo1_tanh_elu_sis", "sac_plain_relu", "saceo2_sepreward"])
def test_cuda_fp32_engine_reproduces_reference_updates(name):
    """Three consecutive updates on the fp32 engine vs the reference's own outputs (gradients 2e-5, θ 1e-6, Δθ 2e-3)."""
    print("\n[%s] %s" % (name, {k: float("%.2e" % v) for k, v in _device_run(name, "GEMM_FP32_SIMT", False, 2e-5, 2e-6, 2e-3).items()}))


@pytest.mark.gpu
def test_cuda_tcgen05_engine_reproduces_reference_updates_hopper():
    """The benchmarked engine (tcgen05 fp16 hi/lo x3, fused kernels, CUDA graph) at the Hopper benchmark shape (2x256
    / 2x512, B = 256, E = 20), two consecutive updates vs the reference's own outputs through fixed random projections of
    the gradients and of Δθ.  Measured: gradients 1e-6 ... 1.4e-4 (q1: one ReLU unit on the other side of zero, DESIGN.md 4
    "ReLU masks"), Δθ <= 8e-5; bounds 5e-4 / 5e-3 as for every ReLU case on this engine."""
    print("\n[saceo2_hopper_256] %s" % {k: float("%.2e" % v) for k, v in
                                        _device_run("saceo2_hopper_256", "GEMM_TCGEN05_BF16X3", True, 5e-4, None, 5e-3).items()})


# ---------------------------------------------------------------------------------------------------------------------
# TRPO.update (trpo.py:36-198, :200-227, :229-317; update_utils.py:4-24) - rows a14 / a15 and the f4 remainder
# ---------------------------------------------------------------------------------------------------------------------
TRPO_CFGS = dict(
    trpo_psd_tanh=NetCfg(S=9, A=3, actor_hidden=(32, 24), critic_hidden=(8, 8), model_hidden=(24, 24), per_state_std=True,
                         actor_acts=("tanh", "tanh"), std_mult=0.7),
    trpo_sis_relu=NetCfg(S=9, A=3, actor_hidden=(32, 24), critic_hidden=(8, 8), model_hidden=(24, 24), per_state_std=False,
                         actor_acts=("relu", "relu"), std_mult=0.7),
)


def load_trpo_case(name):
    g = np.load(os.path.join(GOLD, f"ref_{name}.npz"))
    cfg = TRPO_CFGS[name]
    S, A, N, E, seed, cg_it, psd = (int(x) for x in g["meta"])
    eps, delta, klf, damp, std_mult = (float(x) for x in g["hyper"])
    assert (S, A, bool(psd), std_mult) == (cfg.S, cfg.A, cfg.per_state_std, cfg.std_mult)
    st, _, _, hyper = make_problem(cfg, 8, E, max(N, 300), seed=seed, perturb=0.2)
    for k in ("actor", "m1", "m2"):
        st[k] = [g[f"in_{k}_{i}"] for i in range(len(st[k]))]
    for k in ("s_mean", "s_std", "m_s_mean", "m_s_std", "m_a_mean", "m_a_std", "m_d_mean", "m_d_std"):
        st[k] = g["in_" + k]
    I1, I2 = np.array_split(g["perm"], 2)                                   # trpo.py:114-116
    batch = dict(sE=g["in_expert_sE"], spE=g["in_expert_spE"], I1=I1, I2=I2, u3=g["u3"].astype(np.float32),
                 u4=g["u4"].astype(np.float32))
    return cfg, g, st, hyper, batch, dict(N=N, E=E, cg_it=cg_it, eps=eps, delta=delta, klf=klf, damp=damp)


@pytest.mark.parametrize("name", list(TRPO_CFGS))
def test_oracle_reproduces_reference_trpo_update(name):
    """fp64 oracle vs the reference's fp32 run: every quantity to fp32-rounding-through-CG tolerances; the line search
    must take the same accept / shrink decisions (``adj`` exactly)."""
    from oracle import sac_eo_oracle as O
    cfg, g, st, hyper, batch, m = load_trpo_case(name)
    for dt, tol in ((torch.float64, 1.0), (torch.float32, 3.0)):
        th = O.to_torch_state(st, dt)
        x = torch.as_tensor(g["fvp_x"]).to(dt)
        Fx = O.make_F(cfg, th["actor"], g["s_all"], th, m["damp"], 1)(x)
        assert rel(Fx.numpy(), g["fvp_Fx"]) < 2e-5 * tol
        s_t, a_t = torch.as_tensor(g["s_all"]).to(dt), torch.as_tensor(g["a_all"]).to(dt)
        with torch.no_grad():
            mean, ls = O.gaussian_forward(cfg, th["actor"], s_t, th)
            assert rel(O.gaussian_neglogp(mean, ls, a_t).numpy(), g["neglogp_old"]) < 2e-6 * tol
            assert rel(O.gaussian_entropy(ls).numpy(), g["entropy"]) < 2e-6 * tol
        new, log, pg_vec, eta_v = O.trpo_update(cfg, th["actor"], g["s_all"], g["a_all"], g["adv_all"], th,
                                                delta=m["delta"], cg_iters=m["cg_it"], trust_damp=m["damp"],
                                                kl_maxfactor=m["klf"], alpha=0.0, ent_targ=-cfg.A, expert=batch, eps=m["eps"])
        assert rel(pg_vec.numpy(), g["pg_vec"]) < 1e-5 * tol
        assert log["adj"] == float(g["log_adj"])
        if name in CG_ILL_CONDITIONED and dt == torch.float64:
            continue
        # eta_v_flat as handed to _backtrack (before any shrink); the oracle returns the accepted (shrunk) step
        assert rel(eta_v.numpy() / max(log["adj"], 1e-30), g["eta_v_flat"]) < 2e-3 * tol
        assert rel(O.flat(new).numpy(), g["theta_new"]) < 2e-4 * tol
        for k in ("ent", "tv_pre", "kl_pre", "tv", "kl", "improve"):
            assert abs(log[k] - float(g["log_" + k])) <= 5e-3 * tol * max(abs(float(g["log_" + k])), 1e-3), (k, log[k], float(g["log_" + k]))


# The reference runs cg() in fp32 (update_utils.py:4-24 on a float32 pg_vec).  In trpo_sis_relu the 10-iteration solve is
# far from converged (residual 0.45 |b|) and fp32 round-off has already broken conjugacy: an fp64 solve of the same
# system lands 13 % away from the reference's v_flat while the fp32 oracle reproduces it to 2e-3 - so the fp64 twin is
# compared on the well-conditioned quantities only (Fisher-vector product, gradient, line-search decisions).
CG_ILL_CONDITIONED = {"trpo_sis_relu"}


def test_reference_trpo_line_search_shrinks_once():
    """trpo_sis_relu runs with kl_maxfactor 0.6: the reference's own _backtrack rejected the full step (kl_pre > 0.6 delta)
    and accepted after one sqrt(2) shrink."""
    g = np.load(os.path.join(GOLD, "ref_trpo_sis_relu.npz"))
    assert float(g["log_kl_pre"]) > 0.6 * float(g["hyper"][1]) >= float(g["log_kl"])
    assert abs(float(g["log_adj"]) - 1 / np.sqrt(2)) < 1e-12
    assert float(np.load(os.path.join(GOLD, "ref_trpo_psd_tanh.npz"))["log_adj"]) == 1.0
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(TRPO_CFGS))
def test_cuda_reproduces_reference_trpo_update(name):
    """saceo_fvp and the whole ``Population.trpo_update`` (surrogate gradient, two-model expert blend, CG, step length,
    line search) against what the reference's own TRPO.update produced.  The ill-conditioned case is compared on the
    Fisher-vector product, the gradient norms and the line-search decisions; its 10-iteration fp32 CG solution only to
    the spread two fp32 implementations of the same recurrence show (see CG_ILL_CONDITIONED)."""
    from sac_expert_b200.population import Population
    from tests.helpers import spec_from_cfg
    cfg, g, st, hyper, batch, m = load_trpo_case(name)
    B, E, N, A = 8, m["E"], m["N"], cfg.A
    pop = Population(spec_from_cfg(cfg, 1, B, E, 16, fvp_rows=N))
    L = pop.L
    pop.load_agent(0, st, hyper)
    pop.set_expert(0, batch["sE"], batch["spE"])
    pop.t["fvp_states"][0].copy_(torch.from_numpy(g["s_all"]))
    noise = np.zeros((1, 3 * B + E, A), np.float32)
    noise[0, 2 * B:2 * B + E] = np.concatenate([batch["u3"], batch["u4"]])
    pop.set_draws(noise=noise, perm=np.concatenate([batch["I1"], batch["I2"]]).astype(np.int32)[None])
    xd = torch.zeros(1, L.na_stride)
    xd[0, :L.na] = torch.from_numpy(g["fvp_x"])
    Fx = pop.fvp(xd, m["damp"]).cpu().numpy()[0, :L.na]
    assert rel(Fx, g["fvp_Fx"]) < 1e-4
    before = pop.t["actor"].cpu().numpy().copy()
    logs = pop.trpo_update(g["a_all"][None], g["adv_all"][None], delta=m["delta"], cg_iters=m["cg_it"],
                           trust_damp=m["damp"], kl_maxfactor=m["klf"], alpha=0.0, expert_eps=m["eps"])
    after = pop.t["actor"].cpu().numpy()
    log = logs[0]
    loose = name in CG_ILL_CONDITIONED
    assert abs(log["adj"] - float(g["log_adj"])) < 1e-6, (log, float(g["log_adj"]))
    assert abs(log["norm_pg"] - float(g["log_norm_pg"])) < 1e-4 * float(g["log_norm_pg"])
    assert abs(log["norm_MSE"] - float(g["log_norm_MSE"])) < 1e-4 * float(g["log_norm_MSE"])
    assert abs(log["ent"] - float(g["log_ent"])) < 1e-5 * abs(float(g["log_ent"]))
    step_ref = g["theta_new"].astype(np.float64) - flat(st["actor"]).astype(np.float64)
    e = rel(after[0, :L.na] - before[0, :L.na], step_ref)
    print(f"\n[{name}] Fx {rel(Fx, g['fvp_Fx']):.2e}  step {e:.2e}  log {log}")
    assert e < (5e-2 if loose else 5e-3), e
    for k in ("tv_pre", "kl_pre", "tv", "kl", "improve"):
        ref = float(g["log_" + k])
        assert abs(log[k] - ref) <= (5e-2 if loose else 1e-2) * max(abs(ref), 1e-3), (k, log[k], ref)
    pop.close()


# ---------------------------------------------------------------------------------------------------------------------
# SAC_exp._update_models (SAC_expert.py:478-621) + _expert_preprocess (:375-404) - rows f1 and a13
# ---------------------------------------------------------------------------------------------------------------------
FIT_CFG = NetCfg(S=5, A=2, actor_hidden=(16, 16), critic_hidden=(8, 8), model_hidden=(24, 24), num_models=2)


FIT_CFGS = dict(fit_mse_relu=FIT_CFG,
                fit_gauss_tanh=NetCfg(S=5, A=2, actor_hidden=(16, 16), critic_hidden=(8, 8), model_hidden=(24, 24), num_models=2,
                                      model_acts=("tanh", "relu")))


def load_fit_case(name="fit_mse_relu"):
    g = np.load(os.path.join(GOLD, f"ref_{name}.npz"))
    cfg = FIT_CFGS[name]
    S, A, n_rows, E, seed, epochs, mbs = (int(x) for x in g["meta"])
    st, replay, expert, hyper = make_problem(cfg, 8, E, n_rows, seed=seed, perturb=0.05)
    for k in ("actor", "m1", "m2"):
        st[k] = [g[f"in_{k}_{i}"] for i in range(len(st[k]))]
    for k in ("s_mean", "s_std", "m_s_mean", "m_s_std", "m_a_mean", "m_a_std", "m_d_mean", "m_d_std"):
        st[k] = g["in_" + k]
    replay = {k: g["in_replay_" + k] for k in replay}
    # minibatches the reference drew: per epoch one permutation per model, split every mbs columns, ragged tail dropped
    batches = []
    for ep in range(epochs):
        idx = g["shuffles"][2 * ep:2 * ep + 2]
        sec = np.array_split(idx, np.arange(0, n_rows, mbs)[1:], axis=1)
        batches += sec[:-1] if n_rows % mbs else sec
    fit = dict(model_lr=float(g["hyper"][1]), model_max_grad_norm=float(g["hyper"][0]))
    if "in_logstd" in g.files:
        coef, dclip, scale = (float(x) for x in g["setup"])
        fit.update(gaussian=True, reward_loss_coef=coef, delta_clip_loss=dclip, scale_model_loss=bool(scale))
    return g, st, replay, batches, dict(E=E, mgn=float(g["hyper"][0]), lr=float(g["hyper"][1]), fit=fit, cfg=cfg)


@pytest.mark.parametrize("name", list(FIT_CFGS))
def test_oracle_reproduces_reference_model_fitting(name):
    from oracle import sac_eo_oracle as O
    g, st, replay, batches, m = load_fit_case(name)
    cfg, fit = m["cfg"], m["fit"]
    gauss = bool(fit.get("gaussian"))
    th = O.to_torch_state(st, torch.float32)
    models = [list(th["m1"]), list(th["m2"])]
    if gauss:
        models = [ml + [torch.as_tensor(g["in_logstd"][k])[None]] for k, ml in enumerate(models)]
    adam = dict(m=[[torch.zeros_like(w) for w in ml] for ml in models], v=[[torch.zeros_like(w) for w in ml] for ml in models], t=0)
    clipped = 0
    for idx in batches:
        bs = [{k: torch.as_tensor(replay[k][idx[j]]) for k in ("s", "a", "sp", "r")} for j in range(2)]
        o = O.apply_model_grads(cfg, models, adam, bs, th, fit)
        clipped += float(o["gnorm"]) > m["mgn"] * 2
        models, adam = o["models"], dict(m=o["m"], v=o["v"], t=o["t"])
    assert len(batches) == 8 and 0 < clipped                                  # the global-norm clip was active
    for k, ml in zip(("m1", "m2"), models):
        d0 = flat(st[k]).astype(np.float64)
        assert rel(flat(ml[:6]).astype(np.float64) - d0, g["theta_" + k].astype(np.float64) - d0) < 2e-4, k
        if gauss:
            i = int(k[1]) - 1
            assert rel(ml[6].numpy().ravel() - g["in_logstd"][i], g["logstd_" + k] - g["in_logstd"][i]) < 2e-4
    th["m1"], th["m2"] = models[0][:6], models[1][:6]
    mse_e = O.model_mse_on_expert(cfg, th, g["in_expert_sE"], g["in_expert_aE"], g["in_expert_spE"], use_expert_actions=True)
    mse_c = O.model_mse_on_expert(cfg, th, g["in_expert_sE"], g["in_expert_aE"], g["in_expert_spE"],
                                  u=g["u_cf"].astype(np.float32))
    assert abs(float(mse_e) - float(g["mse_expert"])) < 1e-5 * float(g["mse_expert"])
    assert abs(float(mse_c) - float(g["mse_counterfactual"])) < 1e-5 * float(g["mse_counterfactual"])
    eps, coeff, j_cur, j_exp = (float(x) for x in g["adaptive"])
    got = O.adaptive_epsilon(eps, scale_by_true_mse=True, mse_cf=float(g["mse_counterfactual"]), j_cur=j_cur, j_exp=j_exp,
                             min_mult=True, exp_mult=True, mult_coeff=coeff)
    # np.float32 bookkeeping value x Python float: float32 arithmetic under NumPy >= 2, float64 under NumPy 1.x
    assert abs(got - float(g["epsilon_coef"])) < 1e-7 * got


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(FIT_CFGS))
def test_cuda_reproduces_reference_model_fitting(name):
    """saceo_model_fit over the eight minibatches the reference's _update_models drew (per-model shuffles, global-norm
    clip active, one joint Adam; MSE and Gaussian-NLL losses): final weights (and logstd) against the reference's."""
    from sac_expert_b200.population import Population
    from tests.helpers import spec_from_cfg
    g, st, replay, batches, m = load_fit_case(name)
    cfg, fit = m["cfg"], dict(m["fit"])
    gauss = bool(fit.pop("gaussian", False))
    mbs = batches[0].shape[1]
    pop = Population(spec_from_cfg(cfg, 1, 8, m["E"], len(replay["r"])))
    pop.fit_bind(mbs, use_grad_clip=True, gaussian=gauss, std_mult=1.0)
    if gauss:
        pop.t["model_logstd"].copy_(torch.as_tensor(g["in_logstd"])[None].to(pop.dev))
    _, _, _, hyper = make_problem(cfg, 8, m["E"], len(replay["r"]), seed=int(g["meta"][4]), perturb=0.05)
    pop.load_agent(0, st, hyper)
    pop.append_rows(0, replay["s"], replay["a"], replay["r"], replay["sp"], replay["d"])
    pop.set_fit_hyper(0, r_mean=0.0, r_std=1.0, **fit)
    pop.model_fit(np.stack(batches)[:, None])                        # [steps, agent, model, minibatch]
    torch.cuda.synchronize()
    worst = 0.0
    for i, k in enumerate(("m1", "m2")):
        d0 = flat(st[k]).astype(np.float64)
        e = rel(flat(pop.get_net(0, k)).astype(np.float64) - d0, g["theta_" + k].astype(np.float64) - d0)
        if gauss:
            e = max(e, rel(pop.t["model_logstd"][0, i].cpu().numpy() - g["in_logstd"][i], g["logstd_" + k] - g["in_logstd"][i]))
        worst = max(worst, e)
        assert e < 2e-3, (k, e)
    print(f"\n[{name}] dtheta vs reference {worst:.2e}")
    pop.close()


# ---------------------------------------------------------------------------------------------------------------------
# BC._update (BC.py:298-363) - row f4 (behaviour cloning from expert observations)
# ---------------------------------------------------------------------------------------------------------------------
BC_CFG = NetCfg(S=5, A=2, actor_hidden=(32, 24), critic_hidden=(8, 8), model_hidden=(24, 24), num_models=2)


def load_bc_case():
    g = np.load(os.path.join(GOLD, "ref_bc2_relu.npz"))
    S, A, E, seed, K = (int(x) for x in g["meta"])
    st, replay, expert, hyper = make_problem(BC_CFG, 8, E, 50, seed=seed, perturb=0.05)
    for k in ("actor", "m1", "m2"):
        st[k] = [g[f"in_{k}_{i}"] for i in range(len(st[k]))]
    st["adam_actor"] = dict(m=[g[f"in_adam_actor_m_{i}"] for i in range(len(st["actor"]))],
                            v=[g[f"in_adam_actor_v_{i}"] for i in range(len(st["actor"]))], t=int(g["in_adam_actor_t"]))
    for k in ("s_mean", "s_std", "m_s_mean", "m_s_std", "m_a_mean", "m_a_std", "m_d_mean", "m_d_std"):
        st[k] = g["in_" + k]
    expert = {k: g["in_expert_" + k] for k in expert}
    hyper["lr_pi"] = float(g["lr_pi"])

    def batch(step):
        I1, I2 = np.array_split(g[f"step{step}_perm"], 2)
        return dict(sE=expert["sE"], spE=expert["spE"], I1=I1, I2=I2, u3=g[f"step{step}_u3"].astype(np.float32),
                    u4=g[f"step{step}_u4"].astype(np.float32))
    return g, st, replay, expert, hyper, batch, E, K


def test_oracle_reproduces_reference_bc_updates():
    from oracle.sac_eo_oracle import bc_update
    g, st, replay, expert, hyper, batch, E, K = load_bc_case()
    state = to_torch_state(st, torch.float32)
    for step in range(K):
        o = bc_update(BC_CFG, state, batch(step), hyper)
        assert abs(float(o["mse"]) - float(g[f"step{step}_mse"])) < 2e-6 * float(o["mse"])
        assert rel(flat(o["g_actor"]), g[f"step{step}_g_actor"]) < 5e-6
        assert rel(flat(o["new"]["actor"]), g[f"step{step}_theta_actor"]) < 1e-6
        state["actor"], state["adam_actor"] = o["new"]["actor"], o["new"]["adam_actor"]


@pytest.mark.gpu
def test_cuda_reproduces_reference_bc_updates():
    from sac_expert_b200.population import Population
    from tests.helpers import spec_from_cfg
    g, st, replay, expert, hyper, batch, E, K = load_bc_case()
    B, A = 8, BC_CFG.A
    pop = Population(spec_from_cfg(BC_CFG, 1, B, E, 50))
    pop.load_agent(0, st, hyper)
    pop.append_rows(0, replay["s"], replay["a"], replay["r"], replay["sp"], replay["d"])
    pop.set_expert(0, expert["sE"], expert["spE"])
    worst = 0.0
    for step in range(K):
        b = batch(step)
        noise = np.zeros((1, 3 * B + E, A), np.float32)
        noise[0, 2 * B:2 * B + E] = np.concatenate([b["u3"], b["u4"]])
        pop.set_draws(np.zeros((1, B), np.int64), noise, np.concatenate([b["I1"], b["I2"]]).astype(np.int32)[None])
        losses = pop.bc_update(1, use_device_rng=False).cpu().numpy()
        torch.cuda.synchronize()
        assert abs(losses[0, 3] - float(g[f"step{step}_mse"])) < 1e-5 * float(g[f"step{step}_mse"])
        want = g[f"step{step}_g_actor"]
        e = rel(pop.debug("g_actor").cpu().numpy().reshape(-1)[:want.size], want)
        assert e < 2e-5, e
        d0 = flat(st["actor"]).astype(np.float64)
        e = rel(flat(pop.get_net(0, "actor")).astype(np.float64) - d0, g[f"step{step}_theta_actor"].astype(np.float64) - d0)
        worst = max(worst, e)
        assert e < 2e-3, e
    print(f"\n[bc2_relu] dtheta vs reference {worst:.2e}")
    pop.close()


# ---------------------------------------------------------------------------------------------------------------------
# the reference's pure-NumPy host classes, run as they are (no emulation): rows a1 - a3
# ---------------------------------------------------------------------------------------------------------------------
def _replay_host_case(buf, nz, g):
    for t in range(3):
        tr = {k: g[f"traj{t}_{k}"] for k in ("s", "a", "r", "sp", "d")}
        buf.add(tr["s"], tr["a"], tr["r"], tr["sp"], tr["d"])
        if nz is not None:
            nz.update_rms(tr["s"], tr["a"], tr["r"], tr["sp"])
        assert [buf.current_size, buf.traj_total, buf.steps_total] == g[f"after{t}_size"].tolist()


def test_mirror_buffer_and_normalizers_match_reference_bitwise():
    """The mirror TrajectoryBuffer (pre-allocated sliding window) and RunningNormalizers against the reference's own
    classes: identical bytes in every exposed array after appends that overflow buffer_size, identical statistics."""
    from sac_expert_b200.sac_eo.common.buffers import TrajectoryBuffer
    from sac_expert_b200.sac_eo.common.normalizer import RunningNormalizers
    g = np.load(os.path.join(GOLD, "ref_host_buffers_normalizers.npz"))
    S, A, cap = (int(x) for x in g["meta"])
    buf, nz = TrajectoryBuffer(S, A, 0.99, 0.95, cap), RunningNormalizers(S, A, 0.99)
    _replay_host_case(buf, nz, g)
    for k in ("s_all", "a_all", "r_all", "sp_all", "d_all", "idx_all"):
        got, want = np.asarray(getattr(buf, k)), g["buf_" + k]
        assert got.dtype == want.dtype and got.shape == want.shape and got.tobytes() == want.tobytes(), k
    np.random.seed(7)
    idx = np.random.randint(buf.current_size, size=16)                    # buffers.py:135
    assert buf.s_all[idx].tobytes() == g["off_s"].tobytes() and buf.d_all[idx].tobytes() == g["off_d"].tobytes()
    for nm_, r_ in zip(("s", "a", "r", "delta", "ret"), nz.get_rms()):
        assert int(r_.t_last) == int(g[f"rms_{nm_}_t"])
        for stat in ("mean", "var", "std"):
            got, want = np.asarray(getattr(r_, stat)), g[f"rms_{nm_}_{stat}"]
            assert got.dtype == want.dtype, (nm_, stat, got.dtype, want.dtype)
            assert np.array_equal(got, want), (nm_, stat)
    assert np.array_equal(np.asarray(nz.s_rms.normalize(g["norm_x"])), g["norm_y"])
    assert np.array_equal(np.asarray(nz.delta_rms.denormalize(g["norm_x"])), g["denorm_y"])


@pytest.mark.gpu
def test_device_replay_gather_matches_reference_bitwise():
    """The device ring behind the mirror buffer: the reference's seeded get_offmodel_info / get_model_info draws after
    the buffer overflowed, byte for byte."""
    from sac_expert_b200.population import Population
    from sac_expert_b200.sac_eo.common.buffers import TrajectoryBuffer
    from tests.helpers import spec_from_cfg
    g = np.load(os.path.join(GOLD, "ref_host_buffers_normalizers.npz"))
    S, A, cap = (int(x) for x in g["meta"])
    cfg = NetCfg(S=S, A=A, actor_hidden=(16, 16), critic_hidden=(16, 16), model_hidden=(16, 16), num_models=0)
    pop = Population(spec_from_cfg(cfg, 2, 16, 0, cap))
    buf = TrajectoryBuffer(S, A, 0.99, 0.95, cap)
    buf.attach(pop, agent=1)
    _replay_host_case(buf, None, g)
    np.random.seed(7)
    for k, v in zip(("s", "a", "sp", "r", "d"), buf.get_offmodel_info(batch_size=16)):
        want = g["off_" + k]
        assert v.dtype == want.dtype and v.tobytes() == want.tobytes(), k
    for k, v in zip(("s", "a", "sp", "r"), buf.get_model_info(batch_size=8)):
        assert np.asarray(v).tobytes() == g["mod_" + k].tobytes(), k
    pop.close()


def test_keras_adam_one_minus_beta_rounding_variants_are_bounded():
    """oracle/tfemu implements the TF >= 2.11 Keras formula (Python-double ``1 - beta`` rounded to fp32); the fused
    ApplyAdam kernel of TF <= 2.10 forms it in fp32.  The two differ by 1.3e-5 relative in v and < 5e-5 in 20 accumulated steps -
    far inside every Δθ tolerance used against the reference vectors, so the pin does not hinge on the TF version."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_tfemu_tensorflow", os.path.join(os.path.dirname(GOLD), "..", "oracle", "tfemu", "tensorflow", "__init__.py"))
    tf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tf)
    rng = np.random.default_rng(0)
    w0 = rng.standard_normal(4096).astype(np.float32)
    out = {}
    for legacy in (False, True):
        v = tf.Variable(w0.copy(), dtype=tf.float32)
        opt = tf.keras.optimizers.Adam(learning_rate=3e-4, legacy_one_minus_beta=legacy)
        r = np.random.default_rng(1)
        for _ in range(20):
            opt.apply_gradients([(tf.constant(r.standard_normal(4096).astype(np.float32)), v)])
        out[legacy] = (v.numpy() - w0, opt.get_slot_arrays(v)[1])
    assert 5e-6 < rel(out[True][1], out[False][1]) < 2e-5          # v: (1 - 0.999) rounding
    assert rel(out[True][0], out[False][0]) < 5e-5                 # accumulated step (incl. the fp32 rounding of theta itself)


# ---------------------------------------------------------------------------------------------------------------------
# PPO.update (ppo.py:41-119, 121-238 with expert_reg = None) - row f4
# ---------------------------------------------------------------------------------------------------------------------
def test_oracle_reproduces_reference_ppo_update():
    from oracle import sac_eo_oracle as O
    g = np.load(os.path.join(GOLD, "ref_ppo_psd_tanh.npz"))
    S, A, N, seed, update_it, nminibatch, psd = (int(x) for x in g["meta"])
    eps_ppo, mgn, lr, std_mult = (float(x) for x in g["hyper"])
    cfg = NetCfg(S=S, A=A, actor_hidden=(32, 24), critic_hidden=(8, 8), model_hidden=(8, 8), per_state_std=bool(psd),
                 actor_acts=("tanh", "tanh"), std_mult=std_mult, num_models=0)
    st, _, _, _ = make_problem(cfg, 8, 4, max(N, 300), seed=seed, perturb=0.2)
    st["actor"] = [g[f"in_actor_{i}"] for i in range(len(st["actor"]))]
    st["s_mean"], st["s_std"] = g["in_s_mean"], g["in_s_std"]

    class Replay:                                  # np.random.shuffle, replayed: the permutations the reference drew
        def __init__(self):
            self.k = 0

        def shuffle(self, idx):
            idx[:] = g["shuffles"][self.k]
            self.k += 1
    th = O.to_torch_state(st, torch.float32)
    adam = {"m": [torch.zeros_like(t) for t in th["actor"]], "v": [torch.zeros_like(t) for t in th["actor"]], "t": 0}
    new, log = O.ppo_update(cfg, th["actor"], adam, g["s_all"], g["a_all"], g["adv_all"], th, actor_lr=lr,
                            actor_update_it=update_it, actor_nminibatch=nminibatch, eps_ppo=eps_ppo, max_grad_norm=mgn,
                            alpha=0.0, ent_targ=-A, np_rng=Replay())
    d0 = flat(st["actor"]).astype(np.float64)
    assert rel(O.flat(new).numpy().astype(np.float64) - d0, g["theta_new"].astype(np.float64) - d0) < 1e-3
    assert rel(flat(adam["m"]), g["adam_m"]) < 1e-4 and rel(flat(adam["v"]), g["adam_v"]) < 1e-4
    assert adam["t"] == update_it * nminibatch
    assert float(g["log_actor_grad_norm_pre"]) > mgn and abs(float(g["log_actor_grad_norm"]) - mgn) < 1e-6   # clip active
    for k in ("ent", "tv", "kl", "outside_clip", "actor_grad_norm_pre", "actor_grad_norm"):
        assert abs(log[k] - float(g["log_" + k])) <= 2e-3 * max(abs(float(g["log_" + k])), 1e-3), (k, log[k], float(g["log_" + k]))


# ---------------------------------------------------------------------------------------------------------------------
# the drop-in boundary itself: this repository's reference-named classes, driven like the reference's, same seeds
# ---------------------------------------------------------------------------------------------------------------------
def _build_mirror(cfg, st, replay, expert, hyper, m):
    """This repository's reference-named classes around a single-agent device population, loaded with a golden problem."""
    from sac_expert_b200 import lib as _lib
    from sac_expert_b200.sac_eo.actors.init_actor import init_actor
    from sac_expert_b200.sac_eo.algs.init_alg import init_alg
    from sac_expert_b200.sac_eo.critics.init_critic import init_critics
    from sac_expert_b200.sac_eo.envs.synthetic import SyntheticEnv
    from sac_expert_b200.sac_eo.models.init_world_models import init_world_models
    env = SyntheticEnv(cfg.S, cfg.A)
    setup = dict(separate_reward_nn=False, reward_loss_coef=1.0, scale_model_loss=False, delta_clip_loss=None,
                 reward_clip_loss=None, delta_clip_pred=cfg.delta_clip_pred or None, reward_clip_pred=None)
    actor = init_actor(env, list(cfg.actor_hidden), list(cfg.actor_acts), 0.01, 1.0, "orthogonal", False, None,
                       cfg.per_state_std, True, False)
    critics, q_targets, q_critics = init_critics(env, list(cfg.critic_hidden), list(cfg.critic_acts), 1.0, None, 2, False,
                                                 "orthogonal", False)
    models = init_world_models(env, list(cfg.model_hidden), list(cfg.model_acts), 0.01, 1.0, None, list(cfg.model_hidden),
                               list(cfg.model_acts), 0.01, None, max(m["nm"], 1), False, setup)
    kw = dict(alg_type="sac_imit" if m["nm"] else "sac", sac_batch_size=m["B"], expert_buffer_size=m["E"], gamma=hyper["gamma"],
              soft_tau=hyper["tau"], q_crit_lr=hyper["lr_q"], mbpo_actor_lr=hyper["lr_pi"], mbpo_alpha_lr=hyper["lr_alpha"],
              alg_seed=0, epsilon=hyper["eps"], target_update_int=m["tui"], device_replay_capacity=m["N"],
              env_buffer_size=m["N"], only_model_normalizer=True, gemm_mode=_lib.GEMM_FP32_SIMT)
    alg = init_alg(0, env, env, env, actor, critics, q_targets, q_critics, models, kw, {}, None, None)
    for nz, pre in ((alg.normalizer, ""), (alg.model_normalizer, "m_")):
        nz.s_rms.mean, nz.s_rms.std = st[pre + "s_mean"].copy(), st[pre + "s_std"].copy()
        nz.a_rms.mean, nz.a_rms.std = st[pre + "a_mean"].copy(), st[pre + "a_std"].copy()
        for r_ in nz.get_rms():
            r_.version += 1
    alg.normalizer.ret_rms.std = np.float32(st["ret_std"])
    alg.model_normalizer.delta_rms.mean, alg.model_normalizer.delta_rms.std = st["m_d_mean"].copy(), st["m_d_std"].copy()
    alg._set_rms()
    alg.pop.load_agent(0, st, hyper)                    # weights, Adam slots at t = 7, alpha; same normaliser record again
    alg.env_data.add(replay["s"], replay["a"], replay["r"], replay["sp"], replay["d"])
    return alg


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["saceo2_relu", "saceo1_tanh_elu_sis", "sac_plain_relu"])
def test_mirror_classes_with_the_same_seeds_reproduce_the_reference(name):
    """No injected draws here: the mirror ``init_actor / init_critics / init_world_models / init_alg`` objects are loaded
    with the golden problem, the global NumPy RNG is seeded the way the generator seeded it for the reference's run
    (``np.random.seed(1000 + seed)``, ``alg_seed = 0``), and ``alg._update(step[, expert_reg])`` is called K times - the
    call a user of the reference makes.  Parameters after every update must equal what the reference's own classes
    produced, which also proves that the two RNG streams are consumed identically."""
    cfg, g, (st, replay, expert, hyper), m = load_case(name)
    seed = int(g["meta"][5])
    alg = _build_mirror(cfg, st, replay, expert, hyper, m)
    expert_reg = (expert["sE"], expert["aE"], expert["spE"], hyper["eps"], False)
    np.random.seed(1000 + seed)
    worst = 0.0
    for step in range(m["K"]):
        if m["nm"]:
            alg._update(step, expert_reg)
        else:
            alg._update(step)
        assert np.array_equal(alg.last_idx, g[f"step{step}_idx"])          # same minibatch as the reference drew
        for k, net in (("actor", alg.actor), ("q1", alg.q_critics[0]), ("q2", alg.q_critics[1]),
                       ("t1", alg.q_targets[0]), ("t2", alg.q_targets[1])):
            got = flat(net.get_weights())
            assert rel(got, g[f"step{step}_theta_{k}"]) < 2e-6, (k, step)
            d0 = flat(st[k]).astype(np.float64)
            e = rel(got.astype(np.float64) - d0, g[f"step{step}_theta_{k}"].astype(np.float64) - d0)
            worst = max(worst, e)
            assert e < 2e-3, (k, step, e)
        assert abs(alg.alpha - float(g[f"step{step}_alpha"])) < 1e-6
    print(f"\n[{name}] mirror classes, same seeds: worst dtheta vs reference {worst:.2e}")


# ---------------------------------------------------------------------------------------------------------------------
# SAC_exp._calc_disc / disagreement-scaled expert weight (SAC_expert.py:405-460) - row a13
# ---------------------------------------------------------------------------------------------------------------------
def test_oracle_reproduces_reference_model_disagreement():
    from oracle import sac_eo_oracle as O
    cfg, _, (st, replay, expert, hyper), m = load_case("saceo2_relu")
    g = np.load(os.path.join(GOLD, "ref_disc_saceo2.npz"))
    th = O.to_torch_state(st, torch.float32)
    sE = torch.as_tensor(expert["sE"])
    with torch.no_grad():
        for u, check_eps in ((g["u_pre"], True), (g["u_disc"], False)):
            act, _ = O.head(cfg, th["actor"], sE, torch.as_tensor(u.astype(np.float32)), th)
            act = torch.clamp(act, -1.0, 1.0)                                   # tf_clip to the action limits (:438)
            disc = torch.linalg.norm(O.model_sample(cfg, th["m1"], sE, act, th) - O.model_sample(cfg, th["m2"], sE, act, th), dim=1).numpy()
            if check_eps:
                assert abs(1 / (float(g["epsilon"]) * float(disc.max()) + 1) - float(g["epsilon_coef"])) < 1e-6
            else:
                assert rel(disc / disc.sum(), g["disc_ratio"]) < 1e-5
                assert abs(disc.max() - float(g["max_disc"])) < 1e-6 and abs(np.median(disc) - float(g["median_disc"])) < 1e-6
                assert abs(disc.sum() - float(g["total_disc"])) < 1e-5


@pytest.mark.gpu
def test_mirror_expert_preprocess_disagreement_matches_reference():
    """``alg._expert_preprocess()`` with ``scale_max_disc`` and ``alg._calc_disc`` of the reference-named classes (device
    forwards of the actor and both models), seeded like the reference's run."""
    cfg, _, (st, replay, expert, hyper), m = load_case("saceo2_relu")
    g = np.load(os.path.join(GOLD, "ref_disc_saceo2.npz"))
    alg = _build_mirror(cfg, st, replay, expert, hyper, m)
    E = m["E"]
    alg.expert_data.add(expert["sE"], expert["aE"], np.zeros(E, np.float32), expert["spE"], np.zeros(E))
    alg.scale_max_disc, alg.epsilon, alg.expert_batch_size = True, float(g["epsilon"]), None
    np.random.seed(6000 + 11)
    reg = alg._expert_preprocess()
    ratio, mx, med, tot = alg._calc_disc(expert["sE"], expert["aE"], expert["spE"])
    assert abs(reg[3] - float(g["epsilon_coef"])) < 1e-5 and reg[4] is False
    assert rel(ratio, g["disc_ratio"]) < 1e-4
    for got, key in ((mx, "max_disc"), (med, "median_disc"), (tot, "total_disc")):
        assert abs(got - float(g[key])) < 1e-4 * float(g[key]), key
    print(f"\n[disc_saceo2] eps {reg[3]:.6f} vs {float(g['epsilon_coef']):.6f}, ratio {rel(ratio, g['disc_ratio']):.2e}")


# ---------------------------------------------------------------------------------------------------------------------
# the callers either side of the path: trajectory_sampler and the command line (pure Python in the reference)
# ---------------------------------------------------------------------------------------------------------------------
class _StubEnv:
    """Same deterministic environment as tests/golden/make_golden_reference.py::StubEnv."""

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


class _StubActor:
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


def test_mirror_trajectory_sampler_matches_reference_bitwise():
    """Time-limit truncation (last done forced False), early termination, eval return: the reference's own
    trajectory_sampler on a stub environment vs the mirror's, byte for byte."""
    from sac_expert_b200.sac_eo.common.samplers import trajectory_sampler
    g = np.load(os.path.join(GOLD, "ref_host_buffers_normalizers.npz"))
    S, A, _ = (int(x) for x in g["meta"])
    for tag, (horizon, term_at, ev) in dict(trunc=(6, None, True), term=(9, 4, True), plain=(5, None, False)).items():
        res = trajectory_sampler(_StubEnv(S, term_at), _StubActor(A), horizon, eval=ev)
        assert len(res) == (6 if ev else 5)
        for k, v in zip(("s", "a", "r", "sp", "d", "J"), res):
            want = g[f"samp_{tag}_{k}"]
            got = np.asarray(v)
            assert got.dtype == want.dtype and got.shape == want.shape and got.tobytes() == want.tobytes(), (tag, k)
    assert len(g["samp_term_r"]) == 4 and bool(g["samp_term_d"][-1]) and not bool(g["samp_trunc_d"][-1])


def test_mirror_command_line_has_every_reference_flag_with_its_default():
    """Every one of the reference parser's 135 flags exists in the mirror parser with the same default, and the kwargs
    groups ``gather_inputs`` builds (train_parser.py: all_kwargs) hold the same keys - ``train.py``'s argument set."""
    import json
    from sac_expert_b200.sac_eo.common.train_parser import all_kwargs, create_train_parser
    ref = json.load(open(os.path.join(GOLD, "ref_train_parser.json")))
    mine = vars(create_train_parser().parse_args([]))
    missing = sorted(set(ref["defaults"]) - set(mine))
    assert not missing, missing
    diff = {k: (mine[k], v) for k, v in ref["defaults"].items()
            if (mine[k] if isinstance(mine[k], (int, float, str, bool, list, type(None))) else repr(mine[k])) != v}
    assert not diff, diff
    for grp, keys in ref["groups"].items():
        assert grp in all_kwargs and set(keys) <= set(all_kwargs[grp]), grp


def test_reference_quirks_observed_by_running_it():
    """What DESIGN.md and the oracle docstrings say the reference does in its corner cases, as recorded from its own
    code (tests/golden/ref_quirks.json), and the oracle's matching behaviour."""
    import json
    q = json.load(open(os.path.join(GOLD, "ref_quirks.json")))
    assert q["trpo_update_without_expert_reg"] == "UnboundLocalError"       # grad_final exists only in the expert branches
    assert q["trpo_update_one_model_branch"] == "TypeError"                # epsilon * None (MSE_alpha_grad), trpo.py:111
    assert q["trpo_update_two_model_branch"] is None
    assert q["saceo_update_odd_expert_rows_two_models"] is not None        # halves of different length cannot be added
    assert q["saceo_update_even_expert_rows_two_models"] is None
    assert q["alpha_after_update_from_minus_3"] == float(np.float32(1e-5))  # raw alpha clamped AFTER its Adam step (:348)
    cfg, g, (st, replay, expert, hyper), m = load_case("saceo2_relu")
    b = step_batch(g, 0, replay, expert, m)
    state = to_torch_state(st, torch.float32)
    state["alpha"] = torch.tensor(-3.0)
    assert float(sac_eo_update(cfg, state, b, hyper)["new"]["alpha"]) == float(np.float32(1e-5))
    b["I2"], b["u4"] = b["I2"][:-1], b["u4"][:-1]                           # 4 + 3 expert rows
    with pytest.raises(RuntimeError):
        sac_eo_update(cfg, to_torch_state(st, torch.float32), b, hyper)
