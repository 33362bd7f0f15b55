"""Dynamics-model fitting through the C ABI vs the oracle restatement of
MBRLOnPolicyAlg._apply_model_grads (mbrl_onpolicy_alg.py:301-319) / MSEModel.get_loss
(continuous_models.py:280-302): losses, gradients, global-norm clip, joint Keras Adam."""
import numpy as np
import pytest
import torch

from oracle.sac_eo_oracle import NetCfg, apply_model_grads, make_problem, model_fit_batches, to_torch_state
from sac_expert_b200 import lib as L
from sac_expert_b200.lib import SaceoError
from tests.helpers import rel, spec_from_cfg
from sac_expert_b200.population import Population, unpack_flat

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _setup(cfg, n_agents, mb, N, seed, fit, gemm_mode, capacity=None, **kw):
    pop = Population(spec_from_cfg(cfg, n_agents, 32, 4, capacity or N, gemm_mode=gemm_mode, **kw))
    pop.fit_bind(mb, use_grad_clip=bool(fit.get("model_max_grad_norm")), gaussian=bool(fit.get("gaussian")), std_mult=0.7)
    if fit.get("gaussian"):       # per-column, per-model logstd values instead of the constant initial value
        g = torch.Generator().manual_seed(seed)
        pop.t["model_logstd"].copy_((0.4 * torch.randn(n_agents, 2, cfg.S, generator=g) - 0.3).to(pop.dev))
    probs = []
    for i in range(n_agents):
        st, replay, expert, hyper = make_problem(cfg, 32, 4, N, seed=seed + 13 * i, perturb=0.05)
        st["m_r_mean"] = np.float32(0.1 * (i + 1))
        st["m_r_std"] = np.float32(1.5 + 0.1 * i)
        pop.load_agent(i, st, hyper)
        pop.append_rows(i, replay["s"], replay["a"], replay["r"], replay["sp"], replay["d"])
        f = dict(fit)
        f.pop("gaussian", None)
        f["model_lr"] = fit.get("model_lr", 1e-3) * (1 + 0.5 * i)       # per-agent hyper-parameters differ
        pop.set_fit_hyper(i, r_mean=st["m_r_mean"], r_std=st["m_r_std"], **f)
        if fit.get("gaussian"):
            f["gaussian"] = True
        probs.append((st, replay, f))
    return pop, probs


def _run(cfg, pop, probs, mb, steps, seed, tol=TOL, window=None, verbose=False):
    n, nm = pop.spec.n_agents, cfg.num_models
    rng = np.random.default_rng(seed)
    worst = {}

    def upd(k, v):
        worst[k] = max(worst.get(k, 0.0), float(v))

    state = []
    for st, replay, f in probs:
        T = to_torch_state(st)
        models = [T["m%d" % (k + 1)] for k in range(nm)]
        if probs[0][2].get("gaussian"):
            ai = len(state)
            models = [m + [pop.t["model_logstd"][ai, k].cpu().clone()[None]] for k, m in enumerate(models)]
        adam = dict(m=[[torch.zeros_like(w) for w in m] for m in models],
                    v=[[torch.zeros_like(w) for w in m] for m in models], t=0)
        state.append([T, models, adam])
    for step in range(steps):
        idx = np.stack([model_fit_batches(len(p[1]["r"]) if window is None else window, nm, mb, True, rng)[0] for p in probs])
        losses = pop.model_fit(idx)
        torch.cuda.synchronize()
        g = pop.debug("g_model").cpu().numpy().reshape(n, 2, pop.L.nm_stride)
        for i, (st, replay, f) in enumerate(probs):
            T, models, adam = state[i]
            rows = {k: (v if window is None else v[-window:]) for k, v in replay.items()}
            b = [{k: torch.as_tensor(rows[k][idx[i, m]]) for k in ("s", "a", "sp", "r")} for m in range(nm)]
            o = apply_model_grads(cfg, models, adam, b, T, f)
            for m in range(nm):
                upd("loss", abs(float(losses[0, i, m]) - float(o["losses"][m])) / abs(float(o["losses"][m])))
                gl = o["grads"][m][:-1] if f.get("gaussian") else o["grads"][m]
                ref = np.concatenate([x.numpy().ravel() for x in gl])
                scale = float(pop.debug("fit_gscale").cpu()[i]) if f.get("model_max_grad_norm") else 1.0
                upd("grad", rel(g[i, m, :ref.size] * scale, ref))
                if f.get("gaussian"):
                    S_ = cfg.S
                    gls = pop.debug("g_model_logstd").cpu().numpy().reshape(n, 2, S_)[i, m] * scale
                    upd("grad_logstd", rel(gls, o["grads"][m][-1].numpy().ravel()))
                    upd("logstd", rel(pop.t["model_logstd"][i, m].cpu().numpy(), o["models"][m][-1].numpy().ravel()))
                    upd("logstd_m", rel(pop.t["model_logstd_m"][i, m].cpu().numpy(), o["m"][m][-1].numpy().ravel()))
                    upd("logstd_v", rel(pop.t["model_logstd_v"][i, m].cpu().numpy(), o["v"][m][-1].numpy().ravel()))
                name = "m%d" % (m + 1)
                for ti, (gw, nw, ow) in enumerate(zip(pop.get_net(i, name), o["models"][m], models[m])):
                    upd("theta", rel(gw, nw.numpy()))
                    upd("dtheta", rel(gw - ow.numpy(), nw.numpy() - ow.numpy()))
                    if verbose:
                        upd("dtheta_t%d" % ti, rel(gw - ow.numpy(), nw.numpy() - ow.numpy()))
                        upd("grad_t%d" % ti, rel(unpack_flat(g[i, m] * scale, pop.shapes["model"])[ti], o["grads"][m][ti].numpy()))
                for gw, nw in zip(pop.get_net(i, name, table="model_m"), o["m"][m]):
                    upd("adam_m", rel(gw, nw.numpy()))
                for gw, nw in zip(pop.get_net(i, name, table="model_v"), o["v"][m]):
                    upd("adam_v", rel(gw, nw.numpy()))
            if f.get("model_max_grad_norm"):
                upd("gnorm", abs(float(pop.debug("fit_gnorm").cpu()[i]) - float(o["gnorm"])) / float(o["gnorm"]))
            assert int(pop.t["model_t"][i]) == o["t"]
            # continue from the DEVICE state so that errors do not compound through the oracle's own trajectory
            ex = (lambda tab, m: [pop.t[tab][i, m].cpu().clone()[None]]) if f.get("gaussian") else (lambda tab, m: [])
            state[i][1] = [[torch.from_numpy(w.copy()) for w in pop.get_net(i, "m%d" % (m + 1))] + ex("model_logstd", m) for m in range(nm)]
            state[i][2] = dict(m=[[torch.from_numpy(w.copy()) for w in pop.get_net(i, "m%d" % (m + 1), table="model_m")] + ex("model_logstd_m", m) for m in range(nm)],
                               v=[[torch.from_numpy(w.copy()) for w in pop.get_net(i, "m%d" % (m + 1), table="model_v")] + ex("model_logstd_v", m) for m in range(nm)],
                               t=o["t"])
        if verbose:
            print(step, {k: "%.1e" % v for k, v in worst.items()})
    assert max(worst.values()) < tol, worst
    return worst


@pytest.mark.parametrize("gemm_mode", [L.GEMM_FP32_SIMT, L.GEMM_TCGEN05_BF16X3])
@pytest.mark.parametrize("acts,clip", [(("relu", "relu"), 0.0), (("tanh", "tanh"), 0.5), (("elu", "relu"), 100.0)])
def test_model_fit_small(gemm_mode, acts, clip):
    cfg = NetCfg(S=5, A=2, model_hidden=(64, 48), model_acts=acts)
    fit = dict(model_lr=1e-3, reward_loss_coef=0.7, model_max_grad_norm=clip)
    pop, probs = _setup(cfg, 3, 24, 120, seed=3, fit=fit, gemm_mode=gemm_mode)
    _run(cfg, pop, probs, 24, steps=3, seed=5)
    pop.close()


@pytest.mark.parametrize("gemm_mode", [L.GEMM_FP32_SIMT, L.GEMM_TCGEN05_BF16X3])
@pytest.mark.parametrize("scale,clip", [(False, 0.0), (True, 0.6)])
def test_model_fit_gaussian_nll(gemm_mode, scale, clip):
    """GaussianModel.get_loss (continuous_models.py:101-131): trainable unclipped logstd in the joint optimiser,
    optional scale_model_loss (stop-gradient mean variance) and the global-norm clip over weights AND logstd."""
    cfg = NetCfg(S=7, A=3, model_hidden=(64, 64), model_acts=("tanh", "relu"))
    fit = dict(model_lr=1e-3, reward_loss_coef=0.5, model_max_grad_norm=clip, gaussian=True, scale_model_loss=scale,
               delta_clip_loss=2.5)
    pop, probs = _setup(cfg, 2, 40, 160, seed=12, fit=fit, gemm_mode=gemm_mode)
    w = _run(cfg, pop, probs, 40, steps=3, seed=2)
    assert "grad_logstd" in w and "logstd" in w
    pop.close()


def test_model_fit_gaussian_full_size():
    cfg = NetCfg(S=27, A=8, model_acts=("tanh", "tanh"))
    fit = dict(model_lr=1e-3, gaussian=True, scale_model_loss=True)
    pop, probs = _setup(cfg, 2, 200, 1000, seed=5, fit=fit, gemm_mode=L.GEMM_TCGEN05_BF16X3)
    w = _run(cfg, pop, probs, 200, steps=2, seed=3)
    assert w["grad"] < 2e-4 and w["grad_logstd"] < 2e-4, w
    pop.close()


def test_model_fit_loss_clips_and_single_model():
    cfg = NetCfg(S=6, A=3, model_hidden=(32, 32), num_models=1)
    fit = dict(model_lr=2e-3, reward_loss_coef=1.0, delta_clip_loss=0.8, reward_clip_loss=0.5)
    pop, probs = _setup(cfg, 2, 16, 64, seed=8, fit=fit, gemm_mode=L.GEMM_FP32_SIMT)
    _run(cfg, pop, probs, 16, steps=2, seed=1)
    pop.close()


def test_model_fit_ring_buffer_window():
    """Rows appended past the capacity: logical index 0 is the oldest surviving row (buffers.py:60-66)."""
    cfg = NetCfg(S=4, A=2, model_hidden=(32, 32))
    pop, probs = _setup(cfg, 2, 10, 90, seed=2, fit=dict(), gemm_mode=L.GEMM_FP32_SIMT, capacity=64)
    _run(cfg, pop, probs, 10, steps=2, seed=4, window=64)
    pop.close()


@pytest.mark.parametrize("mode", [L.GEMM_FP32_SIMT, L.GEMM_TCGEN05_BF16X3])
@pytest.mark.parametrize("act", ["relu", "tanh"])
@pytest.mark.parametrize("shape", ["hopper", "ant"])
def test_model_fit_full_size(shape, act, mode):
    """Reference-sized models (2x512, --model_batch_size 200), both engines.
    Tolerances: 1e-3 (L2-relative per tensor) except ReLU on the tensor-core engine, 3e-3: tensor-core fp32
    accumulation truncates (forward error ~2e-6 vs ~2e-7 for FMA chains), so roughly one pre-activation in
    5e5 lands on the other side of zero than in the oracle; ONE flipped ReLU mask moves the L2-relative
    error of the affected gradient tensor by ~1/sqrt(rows*width) ~ 1e-3.  Smooth
    activations on the same engine are held to 2e-4 on gradients."""
    S, A = {"hopper": (11, 3), "ant": (27, 8)}[shape]
    cfg = NetCfg(S=S, A=A, model_acts=(act, act))
    fit = dict(model_lr=1e-3, model_max_grad_norm=10.0)
    pop, probs = _setup(cfg, 2, 200, 1000, seed=21, fit=fit, gemm_mode=mode)
    tc_relu = mode == L.GEMM_TCGEN05_BF16X3 and act == "relu"
    w = _run(cfg, pop, probs, 200, steps=2, seed=9, tol=3e-3 if tc_relu else TOL)
    if not tc_relu:
        assert w["grad"] < 2e-4, w
    pop.close()


def test_model_fit_multi_step_call_matches_single_steps():
    cfg = NetCfg(S=5, A=2, model_hidden=(32, 32))
    rng = np.random.default_rng(0)
    out = []
    for mode in ("single", "multi"):
        pop, probs = _setup(cfg, 2, 12, 60, seed=6, fit=dict(), gemm_mode=L.GEMM_FP32_SIMT)
        idx = np.stack([np.stack([np.stack([np.random.default_rng(100 + s * 7 + i * 3 + m).permutation(60)[:12] for m in range(2)])
                                  for i in range(2)]) for s in range(4)])
        if mode == "single":
            ls = torch.cat([pop.model_fit(idx[s]) for s in range(4)])
        else:
            ls = pop.model_fit(idx)
        torch.cuda.synchronize()
        out.append((ls.cpu().numpy(), pop.t["model"].cpu().numpy().copy()))
        pop.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


@pytest.mark.parametrize("gemm_mode", [L.GEMM_FP32_SIMT, L.GEMM_TCGEN05_BF16X3])
def test_model_fit_separate_reward_nn(gemm_mode):
    """separate_reward_nn (base_world_model.py:34-38, 72-74; continuous_models.py:216-219): the model net predicts the S
    delta columns, a second net the reward; both nets of both models sit in ONE joint optimiser with ONE global-norm
    clip.  Three steps against the oracle: losses, gradients of both nets, the clip norm, parameters and Adam slots."""
    from oracle.sac_eo_oracle import init_net
    big = gemm_mode == L.GEMM_TCGEN05_BF16X3
    cfg = NetCfg(S=11 if big else 5, A=3 if big else 2, model_hidden=(512, 512) if big else (64, 48), model_acts=("tanh", "relu"),
                 separate_reward_nn=True, reward_hidden=(256, 256) if big else (32, 40), reward_acts=("relu", "tanh"))
    n, mb, N, S, A = 2, 200 if big else 24, 600, cfg.S, cfg.A
    fit = dict(model_lr=2e-3, reward_loss_coef=0.7, delta_clip_loss=3.0, reward_clip_loss=2.0, model_max_grad_norm=0.4)
    pop = Population(spec_from_cfg(cfg, n, 32, 4, N, gemm_mode=gemm_mode, reward_hidden=cfg.reward_hidden, reward_acts=cfg.reward_acts))
    pop.fit_bind(mb, use_grad_clip=True)
    rng = np.random.default_rng(77)
    agents = []
    for i in range(n):
        st, replay, expert, hyper = make_problem(cfg, 32, 4, N, seed=50 + 13 * i, perturb=0.05)
        st["m_r_mean"], st["m_r_std"] = np.float32(0.2 * i), np.float32(1.3)
        pop.load_agent(i, st, hyper)
        pop.append_rows(i, replay["s"], replay["a"], replay["r"], replay["sp"], replay["d"])
        pop.set_fit_hyper(i, r_mean=st["m_r_mean"], r_std=st["m_r_std"], **fit)
        rw = []
        for k in range(2):
            w = [x + (0.05 * rng.standard_normal(x.shape)).astype(np.float32) for x in init_net(rng, S + A, cfg.reward_hidden, 1, 0.01)]
            pop.set_net(i, "r%d" % (k + 1), w)
            rw.append(w)
        T = to_torch_state(st)
        models = [T["m%d" % (k + 1)] + [torch.from_numpy(x) for x in rw[k]] for k in range(2)]
        adam = dict(m=[[torch.zeros_like(w) for w in m] for m in models], v=[[torch.zeros_like(w) for w in m] for m in models], t=0)
        agents.append([T, replay, models, adam])
    worst = {}

    def upd(k, v):
        worst[k] = max(worst.get(k, 0.0), float(v))

    for step in range(3):
        idx = np.stack([model_fit_batches(N, 2, mb, True, rng)[0] for _ in range(n)])
        losses = pop.model_fit(idx)
        torch.cuda.synchronize()
        g = pop.debug("g_model").cpu().numpy().reshape(n, 2, pop.L.nm_stride)
        gr = pop.debug("g_reward").cpu().numpy().reshape(n, 2, pop.nr_stride)
        for i, (T, replay, models, adam) in enumerate(agents):
            b = [{k: torch.as_tensor(replay[k][idx[i, m]]) for k in ("s", "a", "sp", "r")} for m in range(2)]
            o = apply_model_grads(cfg, models, adam, b, T, fit)
            scale = float(pop.debug("fit_gscale").cpu()[i])
            upd("gnorm", abs(float(pop.debug("fit_gnorm").cpu()[i]) - float(o["gnorm"])) / float(o["gnorm"]))
            for m in range(2):
                upd("loss", abs(float(losses[0, i, m]) - float(o["losses"][m])) / abs(float(o["losses"][m])))
                ref = np.concatenate([x.numpy().ravel() for x in o["grads"][m][:6]])
                upd("grad", rel(g[i, m, :ref.size] * scale, ref))
                ref_r = np.concatenate([x.numpy().ravel() for x in o["grads"][m][6:]])
                upd("grad_reward", rel(gr[i, m, :ref_r.size] * scale, ref_r))
                for name, sl, tabs in (("m%d" % (m + 1), slice(0, 6), ("model_m", "model_v")),
                                       ("r%d" % (m + 1), slice(6, 12), ("reward_m", "reward_v"))):
                    for gw, nw, ow in zip(pop.get_net(i, name), o["models"][m][sl], models[m][sl]):
                        upd("dtheta_" + name[0], rel(gw - ow.numpy(), nw.numpy() - ow.numpy()))
                    for gw, nw in zip(pop.get_net(i, name, table=tabs[0]), o["m"][m][sl]):
                        upd("adam_m_" + name[0], rel(gw, nw.numpy()))
                    for gw, nw in zip(pop.get_net(i, name, table=tabs[1]), o["v"][m][sl]):
                        upd("adam_v_" + name[0], rel(gw, nw.numpy()))
            # continue from the device state
            agents[i][2] = [[torch.from_numpy(w.copy()) for w in pop.get_net(i, "m%d" % (m + 1)) + pop.get_net(i, "r%d" % (m + 1))] for m in range(2)]
            agents[i][3] = dict(m=[[torch.from_numpy(w.copy()) for w in pop.get_net(i, "m%d" % (m + 1), table="model_m") + pop.get_net(i, "r%d" % (m + 1), table="reward_m")] for m in range(2)],
                                v=[[torch.from_numpy(w.copy()) for w in pop.get_net(i, "m%d" % (m + 1), table="model_v") + pop.get_net(i, "r%d" % (m + 1), table="reward_v")] for m in range(2)],
                                t=o["t"])
    pop.close()
    assert max(worst.values()) < (3e-3 if big else TOL), worst


def test_model_fit_errors():
    cfg = NetCfg(S=4, A=2, model_hidden=(32, 32), separate_reward_nn=True)
    pop = Population(spec_from_cfg(cfg, 1, 16, 4, 32))
    ft = L.FitTables()                       # separate_reward_nn without the reward tables
    for name in ("model",):
        setattr(ft, name, pop.t["model"].data_ptr())
    z = torch.zeros(1, 2, pop.L.nm_stride, device="cuda"); zt = torch.zeros(1, dtype=torch.int32, device="cuda"); zh = torch.zeros(1, 8, device="cuda")
    ft.model_m, ft.model_v, ft.model_t, ft.fit_hyper = z.data_ptr(), z.data_ptr(), zt.data_ptr(), zh.data_ptr()
    import ctypes as C
    with pytest.raises(SaceoError):
        L.check(pop.lib.saceo_fit_bind(pop.ctx, C.byref(ft), 8, 0))
    pop.close()
    cfg = NetCfg(S=4, A=2, model_hidden=(32, 32))
    pop = Population(spec_from_cfg(cfg, 1, 16, 4, 32))
    pop.model_batch = 8
    with pytest.raises(SaceoError):
        pop.model_fit(np.zeros((1, 1, 2, 8), np.int64))
    pop.close()
