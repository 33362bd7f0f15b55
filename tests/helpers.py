"""Shared helpers of the parity tests: build a device population from oracle problems, inject the
draws, and compare against the CPU oracle (oracle/ is the checker, never the thing under test)."""
import numpy as np
import torch

from oracle.sac_eo_oracle import NetCfg, draw_batch, make_problem, sac_eo_update, to_torch_state
from sac_expert_b200.population import Population, PopulationSpec


def rel(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def spec_from_cfg(cfg: NetCfg, n_agents, B, E, N, **kw) -> PopulationSpec:
    # tests default to the exact-fp32 engine launched kernel by kernel; the tensor-core engine / graphs are opted into
    from sac_expert_b200 import lib as _lib
    kw.setdefault("gemm_mode", _lib.GEMM_FP32_SIMT)
    kw.setdefault("use_graph", False)
    return PopulationSpec(n_agents=n_agents, S=cfg.S, A=cfg.A, actor_hidden=cfg.actor_hidden,
                          critic_hidden=cfg.critic_hidden, model_hidden=cfg.model_hidden,
                          actor_acts=cfg.actor_acts, critic_acts=cfg.critic_acts, model_acts=cfg.model_acts,
                          per_state_std=cfg.per_state_std, separate_reward_nn=cfg.separate_reward_nn,
                          num_models=cfg.num_models, delta_clip_pred=cfg.delta_clip_pred, B=B, E=E,
                          replay_capacity=N, std_mult=cfg.std_mult, **kw)


def build(cfg: NetCfg, n_agents, B, E, N, seed=0, perturb=0.05, eps=0.3, identity_norm=False, **kw):
    """-> (population, problems) with problems[i] = (state_np, replay, expert, hyper, batch)."""
    pop = Population(spec_from_cfg(cfg, n_agents, B, E, N, **kw))
    probs = []
    for i in range(n_agents):
        st, replay, expert, hyper = make_problem(cfg, B, E, N, seed=seed + 17 * i, perturb=perturb,
                                                 identity_norm=identity_norm)
        hyper["eps"] = eps
        hyper["gamma"] = 0.995 - 0.01 * i       # per-agent hyper-parameters differ
        batch = draw_batch(cfg, replay, expert, B, seed=seed + 1000 + i)
        pop.load_agent(i, st, hyper)
        pop.append_rows(i, replay["s"], replay["a"], replay["r"], replay["sp"], replay["d"])
        if cfg.num_models > 0:
            pop.set_expert(i, expert["sE"], expert["spE"])
        probs.append((st, replay, expert, hyper, batch))
    inject(pop, cfg, probs)
    return pop, probs


def inject(pop, cfg, probs):
    B, E, A = pop.spec.B, (pop.spec.E if cfg.num_models > 0 else 0), cfg.A
    n = len(probs)
    idx = np.stack([p[4]["idx"] for p in probs])
    noise = np.zeros((n, 3 * B + E, A), np.float32)
    perm = np.zeros((n, max(E, 1)), np.int32)
    for i, p in enumerate(probs):
        b = p[4]
        noise[i, :B] = b["u1"]
        noise[i, B:2 * B] = b["u2"]
        if cfg.num_models == 1:
            noise[i, 2 * B:2 * B + E] = b["u3"]
            perm[i] = np.arange(E)
        elif cfg.num_models == 2:
            noise[i, 2 * B:2 * B + E] = np.concatenate([b["u3"], b["u4"]])
            perm[i] = np.concatenate([b["I1"], b["I2"]])
        noise[i, 2 * B + E:] = b["u5"]
    pop.set_draws(idx, noise, perm if E > 0 else None)


def oracle_update(cfg, prob, dtype=torch.float32):
    st, replay, expert, hyper, batch = prob
    return sac_eo_update(cfg, to_torch_state(st, dtype), batch, hyper)


def compare_update(pop, cfg, probs, tol=1e-3, verbose=False):
    """Runs ONE injected-draw update on the device and compares every tensor with the fp32 oracle.
    Returns the dict of worst relative errors."""
    L = pop.L
    pop.update(1, num_timesteps=0, use_device_rng=False)
    torch.cuda.synchronize()
    losses = pop.losses.cpu().numpy()
    g_q = pop.debug("g_q").cpu().numpy().reshape(pop.spec.n_agents, 2, L.nc_stride)
    g_a = pop.debug("g_actor").cpu().numpy().reshape(pop.spec.n_agents, L.na_stride)
    y = pop.debug("y").cpu().numpy().reshape(pop.spec.n_agents, pop.spec.B)
    worst = {}

    def upd(k, v):
        worst[k] = max(worst.get(k, 0.0), v)

    for i, prob in enumerate(probs):
        o = oracle_update(cfg, prob)
        o64 = oracle_update(cfg, prob, torch.float64)
        upd("oracle32_vs_64_g_actor", rel(np.concatenate([g.numpy().ravel() for g in o["g_actor"]]),
                                         np.concatenate([g.numpy().ravel() for g in o64["g_actor"]])))
        upd("y", rel(y[i], o["y"].numpy()))
        for j, k in enumerate(("L_q1", "L_q2", "L_pi", "mse", "p_loss", "alpha_loss")):
            ref = float(o[k])
            upd(k, abs(losses[i, j] - ref) / max(abs(ref), 1e-12) if ref != 0 else abs(losses[i, j]))
        upd("alpha", abs(losses[i, 6] - float(o["new"]["alpha"])) / abs(float(o["new"]["alpha"])))
        for net, key in ((0, "g_q1"), (1, "g_q2")):
            ref = np.concatenate([g.numpy().ravel() for g in o[key]])
            upd(key, rel(g_q[i, net, :ref.size], ref))
        ref = np.concatenate([g.numpy().ravel() for g in o["g_actor"]])
        upd("g_actor", rel(g_a[i, :ref.size], ref))
        upd("g_alpha", abs(g_a[i, -1] - float(o["g_alpha"])) / max(abs(float(o["g_alpha"])), 1e-12))
        for name in ("q1", "q2", "t1", "t2", "actor"):
            got = pop.get_net(i, name)
            new, old = o["new"][name], prob[0][name]
            for gw, nw, ow in zip(got, new, old):
                upd("theta_" + name, rel(gw, nw.numpy()))
                dref = nw.numpy() - np.asarray(ow)
                if np.linalg.norm(dref) > 0:
                    upd("dtheta_" + name, rel(gw - np.asarray(ow), dref))
        for name, (tm, tv) in (("q1", ("q_m", "q_v")), ("q2", ("q_m", "q_v")), ("actor", ("actor_m", "actor_v"))):
            for gw, nw in zip(pop.get_net(i, name, table=tm), o["new"]["adam_" + name]["m"]):
                upd("adam_m_" + name, rel(gw, nw.numpy()))
            for gw, nw in zip(pop.get_net(i, name, table=tv), o["new"]["adam_" + name]["v"]):
                upd("adam_v_" + name, rel(gw, nw.numpy()))
    if verbose:
        for k, v in worst.items():
            print(f"  {k:28s} {v:.3e}")
    return worst
