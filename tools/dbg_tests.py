import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.sac_eo_oracle import NetCfg, cg, make_F, make_problem, to_torch_state
from sac_expert_b200 import lib as L
from sac_expert_b200.population import Population
from tests.helpers import build, compare_update, oracle_update, rel, spec_from_cfg
cfg = NetCfg(S=11, A=3)
for ws in (True, False):
    pop, probs = build(cfg, n_agents=3, B=256, E=20, N=1500, seed=77, gemm_mode=L.GEMM_TCGEN05_BF16X3, use_graph=True, ws_kernels=ws)
    pop.update(1, num_timesteps=0, use_device_rng=False); torch.cuda.synchronize()
    g_a = pop.debug("g_actor").cpu().numpy().reshape(3, pop.L.na_stride)
    for i in range(3):
        o = oracle_update(cfg, probs[i]); off = 0; errs = []
        for g in o["g_actor"]:
            n = g.numel(); errs.append("%.1e/%.1e" % (rel(g_a[i, off:off + n], g.numpy()), float(np.linalg.norm(g.numpy())))); off += n
        print("hopper ws=%s agent %d per-tensor rel/norm:" % (ws, i), errs)
    pop.close()
# CG
for ws in (True, False):
    cfgc = NetCfg(S=376, A=17, actor_hidden=(256, 256), critic_hidden=(256, 256), num_models=0, per_state_std=False, std_mult=0.8)
    n, N, iters, damp = 1, 1024, 20, 0.01
    pop = Population(spec_from_cfg(cfgc, n, 128, 0, 16, fvp_rows=N, gemm_mode=L.GEMM_TCGEN05_BF16X3, ws_kernels=ws))
    Lo = pop.L; rng = np.random.default_rng(5)
    st, replay, _, hyper = make_problem(cfgc, 128, 2, N + 8, seed=60, perturb=0.1)
    pop.load_agent(0, st, hyper); states = replay["s"][:N]
    pop.t["fvp_states"][0].copy_(torch.from_numpy(states))
    x = rng.standard_normal(Lo.na); b = rng.standard_normal(Lo.na) * 0.1
    xd = torch.zeros(n, Lo.na_stride); bd = torch.zeros(n, Lo.na_stride)
    xd[0, :Lo.na] = torch.from_numpy(x).float(); bd[0, :Lo.na] = torch.from_numpy(b).float()
    th64 = to_torch_state(st, torch.float64)
    F64 = make_F(cfgc, th64["actor"], states, th64, damp=damp)
    ref = F64(torch.from_numpy(x)).numpy()
    Fx = pop.fvp(xd, damp).cpu().numpy()[0, :Lo.na]
    sol64 = cg(F64, torch.from_numpy(b), cg_iters=iters).numpy()
    sol, vfv = pop.cg_solve(bd, iters=iters, tol=1e-10, damp=damp)
    sol = sol.cpu().numpy()[0, :Lo.na]
    off = 0; errs = []
    for w in st["actor"]:
        nn = np.asarray(w).size; errs.append("%.1e" % rel(Fx[off:off + nn], ref[off:off + nn])); off += nn
    print("CG ws=%s: Fx rel %.3e per-tensor %s ; cg rel %.3e" % (ws, rel(Fx, ref), errs, rel(sol, sol64)))
    pop.close()
