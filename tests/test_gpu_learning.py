"""End-to-end sanity beyond single-step parity: many consecutive device updates actually LEARN.
(1) plain SAC on a one-step problem (done = 1, reward = -|a - a*|^2): the critics regress the reward and the
    deterministic policy action moves to a*;  (2) dynamics-model fitting on linear dynamics: the loss collapses."""
import numpy as np
import pytest
import torch

from sac_expert_b200 import lib as L
from sac_expert_b200.population import Population, PopulationSpec

pytestmark = pytest.mark.gpu


def _init(pop, rng, names, shapes_key, gain):
    from oracle.sac_eo_oracle import init_net
    spec = pop.spec
    for i in range(spec.n_agents):
        for nme in names:
            n_in = spec.S if shapes_key == "actor" else spec.S + spec.A
            hidden = {"actor": spec.actor_hidden, "q": spec.critic_hidden, "model": spec.model_hidden}[shapes_key]
            n_out = {"actor": pop.L.Ao, "q": 1, "model": pop.L.model_out}[shapes_key]
            pop.set_net(i, nme, init_net(rng, n_in, hidden, n_out, gain))


def test_sac_learns_one_step_problem():
    n, S, A, N = 4, 3, 2, 4000
    spec = PopulationSpec(n_agents=n, S=S, A=A, actor_hidden=(64, 64), critic_hidden=(64, 64), B=256, E=0, num_models=0,
                          replay_capacity=N, gemm_mode=L.GEMM_FP32_SIMT)
    pop = Population(spec)
    rng = np.random.default_rng(0)
    _init(pop, rng, ["actor"], "actor", 0.01)
    _init(pop, rng, ["q1", "q2"], "q", 1.0)
    targets = np.stack([rng.uniform(-0.6, 0.6, A) for _ in range(n)]).astype(np.float32)
    for i in range(n):
        pop.set_net(i, "t1", pop.get_net(i, "q1"))
        pop.set_net(i, "t2", pop.get_net(i, "q2"))
        pop.t["alpha"][i] = float(np.log(0.1))                       # raw log-initialised temperature (SAC_expert.py:106)
        pop.set_hyper(i, gamma=0.99, tau=5e-3, lr_q=1e-3, lr_pi=1e-3, lr_alpha=1e-4, eps=0.0, target_entropy=float(-A))
        s = rng.standard_normal((N, S)).astype(np.float32)
        a = rng.uniform(-1, 1, (N, A)).astype(np.float32)
        r = -((a - targets[i]) ** 2).sum(-1).astype(np.float32)
        pop.append_rows(i, s, a, r, rng.standard_normal((N, S)).astype(np.float32), np.ones(N))
    obs = torch.from_numpy(rng.standard_normal((n, 64, S)).astype(np.float32))
    mean_action = lambda: (lambda o: o[0] if isinstance(o, tuple) else o)(pop.actor_forward(obs, None)).cpu().numpy()
    a0 = mean_action()
    first = pop.update(1, 0, True, seed=3).cpu().numpy().copy()
    pop.update(2500, 1, True, seed=3)
    last = pop.losses.cpu().numpy().copy()
    a1 = mean_action()
    err0 = np.abs(a0 - targets[:, None, :]).mean(axis=(1, 2))
    err1 = np.abs(a1 - targets[:, None, :]).mean(axis=(1, 2))
    assert np.all(np.isfinite(last))
    assert np.all(last[:, 0] < 0.1 * first[:, 0]) and np.all(last[:, 1] < 0.1 * first[:, 1]), (first[:, :2], last[:, :2])   # critic losses
    assert np.all(err1 < 0.1) and np.all(err1 < 0.35 * err0), (err0, err1)
    assert np.all(last[:, 6] >= 1e-5)                                  # temperature clamp (SAC_expert.py:348)
    pop.close()


def test_model_fit_learns_linear_dynamics():
    n, S, A, N, mb = 3, 6, 2, 3000, 200
    spec = PopulationSpec(n_agents=n, S=S, A=A, actor_hidden=(32, 32), critic_hidden=(32, 32), model_hidden=(64, 64),
                          model_acts=("tanh", "tanh"), B=64, E=4, num_models=2, replay_capacity=N, gemm_mode=L.GEMM_FP32_SIMT)
    pop = Population(spec)
    rng = np.random.default_rng(1)
    _init(pop, rng, ["m1", "m2"], "model", 0.1)
    for i in range(n):
        M = (0.3 * rng.standard_normal((S + A, S))).astype(np.float32)
        w = (0.5 * rng.standard_normal(S + A)).astype(np.float32)
        s = rng.standard_normal((N, S)).astype(np.float32)
        a = rng.uniform(-1, 1, (N, A)).astype(np.float32)
        sa = np.concatenate([s, a], 1)
        pop.append_rows(i, s, a, sa @ w, s + sa @ M, np.zeros(N))
    pop.fit_bind(mb)
    steps = 600
    idx = rng.integers(0, N, size=(steps, n, 2, mb))
    losses = pop.model_fit(idx).cpu().numpy()
    assert np.all(np.isfinite(losses))
    start, end = losses[:5].mean(0), losses[-20:].mean(0)
    assert np.all(end < 0.05 * start), (start, end)
    assert int(pop.t["model_t"][0]) == steps
    pop.close()
