"""Behaviour cloning from expert observations (BC._update_actor, BC.py:309-363) through the C ABI vs the oracle:
the expert-observation MSE alone drives one actor Adam step; everything else must stay bit-identical."""
import numpy as np
import pytest
import torch

from oracle.sac_eo_oracle import NetCfg, bc_update, to_torch_state
from sac_expert_b200 import lib as L
from tests.helpers import build, rel

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _check(cfg, B, E, gemm_mode, n_agents=2, seed=3):
    pop, probs = build(cfg, n_agents=n_agents, B=B, E=E, N=600, seed=seed, gemm_mode=gemm_mode)
    frozen = {k: pop.t[k].clone() for k in ("q", "q_m", "q_v", "qt", "alpha", "alpha_m", "alpha_v", "model")}
    t0 = pop.t["adam_t"].clone()
    losses = pop.bc_update(1, use_device_rng=False).cpu().numpy()
    torch.cuda.synchronize()
    g_a = pop.debug("g_actor").cpu().numpy().reshape(n_agents, pop.L.na_stride)
    for k, v in frozen.items():
        assert torch.equal(pop.t[k], v), k                                   # untouched, bit for bit
    t1 = pop.t["adam_t"]
    assert torch.equal(t1[:, [0, 1, 3]], t0[:, [0, 1, 3]]) and torch.equal(t1[:, 2], t0[:, 2] + 1)
    worst = {}
    for i, (st, replay, expert, hyper, batch) in enumerate(probs):
        o = bc_update(cfg, to_torch_state(st), batch, hyper)
        ref = np.concatenate([g.numpy().ravel() for g in o["g_actor"]])
        worst["g_actor"] = max(worst.get("g_actor", 0), rel(g_a[i, :ref.size], ref))
        worst["mse"] = max(worst.get("mse", 0), abs(losses[i, 3] - float(o["mse"])) / abs(float(o["mse"])))
        for gw, nw, ow in zip(pop.get_net(i, "actor"), o["new"]["actor"], st["actor"]):
            d = nw.numpy() - np.asarray(ow)
            if np.linalg.norm(d) > 0:
                worst["dtheta"] = max(worst.get("dtheta", 0), rel(gw - np.asarray(ow), d))
        for gw, nw in zip(pop.get_net(i, "actor", table="actor_m"), o["new"]["adam_actor"]["m"]):
            worst["adam_m"] = max(worst.get("adam_m", 0), rel(gw, nw.numpy()))
        for gw, nw in zip(pop.get_net(i, "actor", table="actor_v"), o["new"]["adam_actor"]["v"]):
            worst["adam_v"] = max(worst.get("adam_v", 0), rel(gw, nw.numpy()))
    pop.close()
    assert max(worst.values()) < TOL, worst
    return worst


@pytest.mark.parametrize("num_models,per_state_std,acts", [(2, True, ("tanh", "tanh")), (1, True, ("relu", "relu")),
                                                           (2, False, ("elu", "tanh"))])
def test_bc_small(num_models, per_state_std, acts):
    cfg = NetCfg(S=6, A=3, actor_hidden=(32, 32), critic_hidden=(32, 32), model_hidden=(48, 48), actor_acts=acts,
                 model_acts=("relu", "tanh"), num_models=num_models, per_state_std=per_state_std, delta_clip_pred=3.0)
    _check(cfg, B=32, E=8, gemm_mode=L.GEMM_FP32_SIMT)


@pytest.mark.parametrize("gemm_mode", [L.GEMM_FP32_SIMT, L.GEMM_TCGEN05_BF16X3])
def test_bc_full_size(gemm_mode):
    cfg = NetCfg(S=27, A=8, actor_acts=("tanh", "tanh"))
    w = _check(cfg, B=256, E=20, gemm_mode=gemm_mode, seed=11)
    assert w["g_actor"] < 2e-4, w


def test_bc_device_rng_reduces_mse():
    """Many BC steps with in-kernel draws: the expert-observation MSE goes down and stays finite."""
    cfg = NetCfg(S=6, A=3, actor_hidden=(32, 32), critic_hidden=(32, 32), model_hidden=(32, 32), actor_acts=("tanh", "tanh"),
                 model_acts=("tanh", "tanh"))
    pop, probs = build(cfg, n_agents=3, B=32, E=8, N=200, seed=5, gemm_mode=L.GEMM_FP32_SIMT, perturb=0.3)
    for i in range(3):
        pop.set_hyper(i, lr_pi=3e-3)
    first = pop.bc_update(20, True, seed=1).cpu().numpy()[:, 3].copy()
    last = pop.bc_update(400, True, seed=2).cpu().numpy()[:, 3].copy()
    assert np.all(np.isfinite(last)) and np.all(last < first), (first, last)
    pop.close()
