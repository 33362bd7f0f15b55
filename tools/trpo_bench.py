"""Whole-TRPO-update and PPO-minibatch-step throughput for a population (Humanoid-shaped by default, N = 1024 rollout
rows, 20 CG iterations) - wall clock around the host line search (it reads a [n, 8] statistics block per trial), plus
the device time of the surrogate gradient alone.  python tools/trpo_bench.py [shape] [n_agents]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sac_expert_b200.population import Population, PopulationSpec
from sac_expert_b200.synth import fill_synthetic, SHAPES

shape = sys.argv[1] if len(sys.argv) > 1 else "humanoid"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
S, A, B = SHAPES[shape]
N = 1024
pop = Population(PopulationSpec(n_agents=n, S=S, A=A, B=64, E=2, num_models=0, replay_capacity=64, fvp_rows=N, gemm_mode=1))
fill_synthetic(pop, seed=3)
g = torch.Generator(device="cuda").manual_seed(0)
pop.t["fvp_states"].copy_(torch.randn(n, N, S, device="cuda", generator=g))
info = pop.trpo_eval(want_kl_info=True)["kl_info"]
act = info[..., 0] + torch.exp(info[..., 1]) * torch.randn(n, N, A, device="cuda", generator=g)     # rollout actions ~ policy
adv = np.random.default_rng(0).standard_normal((n, N)).astype(np.float32)
theta0 = pop.t["actor"].clone()


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def one_update():
    pop.t["actor"].copy_(theta0)
    return pop.trpo_update(act, adv, delta=0.01, cg_iters=20, trust_damp=0.01)


logs = one_update()
ms_update = timed(one_update, 3)
nlp = pop.trpo_eval(act=act, want_nlp=True)["nlp"]
advd = torch.from_numpy(adv).cuda()
ms_grad = timed(lambda: pop.trpo_grad(act, advd, nlp, None), 10)
ms_eval = timed(lambda: pop.trpo_eval(act, advd, nlp, info), 10)
pop.set_hyper(0, lr_pi=3e-4)


def ppo_step():
    gr, _ = pop.ppo_grad(act, advd, nlp, None, 0.2, 0.5)
    pop.actor_adam(gr)


ms_ppo = timed(ppo_step, 10)
adj = [l["adj"] for l in logs]
print(json.dumps({"shape": shape, "n_agents": n, "rollout_rows": N, "cg_iters": 20,
                  "trpo_update_ms": round(ms_update, 2), "trpo_updates_per_s": round(n / (ms_update * 1e-3), 1),
                  "surrogate_grad_ms": round(ms_grad, 3), "line_search_eval_ms": round(ms_eval, 3),
                  "ppo_minibatch_step_ms": round(ms_ppo, 3), "ppo_steps_per_s": round(n / (ms_ppo * 1e-3), 1),
                  "adj_hist": {str(a): adj.count(a) for a in sorted(set(adj))},
                  "kl_mean": float(np.mean([l["kl"] for l in logs])), "improve_min": float(min(l["improve"] for l in logs)),
                  "timing": "host wall clock around stream-synchronised calls (the line search is host-driven)"}))
