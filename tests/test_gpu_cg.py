"""GPU: Fisher-vector product and CG solve vs the oracle (trpo.py:200-227, update_utils.py:4-24)."""
import numpy as np
import pytest
import torch

from oracle.sac_eo_oracle import NetCfg, cg, make_F, make_problem, to_torch_state, flat
from sac_expert_b200.population import Population
from tests.helpers import rel, spec_from_cfg

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("per_state_std,acts", [(True, ("relu", "relu")), (False, ("tanh", "tanh")), (True, ("elu", "tanh"))])
def test_fvp_and_cg(per_state_std, acts):
    cfg = NetCfg(S=9, A=3, actor_hidden=(64, 48), critic_hidden=(16, 16), num_models=0,
                 per_state_std=per_state_std, actor_acts=acts, std_mult=0.7)
    n, N = 2, 96
    pop = Population(spec_from_cfg(cfg, n, 8, 0, 16, fvp_rows=N))
    L = pop.L
    xs, refs_F, refs_cg, refs_vfv = [], [], [], []
    rng = np.random.default_rng(1)
    bs = []
    for i in range(n):
        st, replay, _, hyper = make_problem(cfg, 8, 2, 200, seed=10 + i, perturb=0.1)
        pop.load_agent(i, st, hyper)
        states = replay["s"][:N]
        pop.t["fvp_states"][i].copy_(torch.from_numpy(states))
        th = to_torch_state(st, torch.float64)
        F = make_F(cfg, th["actor"], states, th, damp=0.01)
        x = rng.standard_normal(L.na)
        b = rng.standard_normal(L.na) * 0.1
        xs.append(x); bs.append(b)
        refs_F.append(F(torch.from_numpy(x)).numpy())
        sol = cg(F, torch.from_numpy(b), cg_iters=10)
        refs_cg.append(sol.numpy())
        refs_vfv.append(float(sol.dot(F(sol))))
    xd = torch.zeros(n, L.na_stride); bd = torch.zeros(n, L.na_stride)
    for i in range(n):
        xd[i, :L.na] = torch.from_numpy(xs[i]).float(); bd[i, :L.na] = torch.from_numpy(bs[i]).float()
    Fx = pop.fvp(xd, 0.01).cpu().numpy()
    sol, vfv = pop.cg_solve(bd, iters=10, tol=1e-10, damp=0.01)
    sol, vfv = sol.cpu().numpy(), vfv.cpu().numpy()
    for i in range(n):
        assert rel(Fx[i, :L.na], refs_F[i]) < 1e-3
        assert rel(sol[i, :L.na], refs_cg[i]) < 5e-3     # 10 fp32 CG iterations vs the fp64 oracle
        assert abs(vfv[i] - refs_vfv[i]) / abs(refs_vfv[i]) < 5e-3
