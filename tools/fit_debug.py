import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.sac_eo_oracle import NetCfg, make_problem, model_fit_batches
from sac_expert_b200 import lib as L
from sac_expert_b200.population import Population
from tests.helpers import rel, spec_from_cfg

cfg = NetCfg(S=11, A=3)
mb = 200
NA = 2
probs = [make_problem(cfg, 32, 4, 1000, seed=21 + 13 * i, perturb=0.05) for i in range(NA)]
rng = np.random.default_rng(9)
idx0 = np.stack([model_fit_batches(1000, 2, mb, True, rng)[0] for _ in range(NA)])
idx1 = np.stack([model_fit_batches(1000, 2, mb, True, rng)[0] for _ in range(NA)])
def mk(mode):
    pop = Population(spec_from_cfg(cfg, NA, 32, 4, 1000, gemm_mode=mode))
    pop.fit_bind(mb, True)
    for i, (st, replay, expert, hyper) in enumerate(probs):
        pop.load_agent(i, st, hyper)
        pop.set_fit_hyper(i, model_max_grad_norm=10.0, model_lr=1e-3 * (1 + 0.5 * i), r_mean=0.1 * (i + 1), r_std=1.5 + 0.1 * i)
        pop.append_rows(i, replay["s"], replay["a"], replay["r"], replay["sp"], replay["d"])
    return pop
names = ("fit_X", "fit_T", "fit_H1", "fit_H2", "fit_Out", "fit_dOut", "fit_dH2", "fit_dH1", "g_model")
def cmp(a, b, tag):
    for i in range(NA):
        print(tag, "agent", i, " ".join("%s=%.1e" % (n, rel(a.debug(n).cpu().numpy().reshape(NA, -1)[i], b.debug(n).cpu().numpy().reshape(NA, -1)[i])) for n in names))
ref = mk(L.GEMM_FP32_SIMT); tc = mk(L.GEMM_TCGEN05_BF16X3)
ref.model_fit(idx0); tc.model_fit(idx0); torch.cuda.synchronize()
cmp(tc, ref, "step0")
for k in ("model", "model_m", "model_v", "model_t"):
    ref.t[k].copy_(tc.t[k])
ref.model_fit(idx1); tc.model_fit(idx1); torch.cuda.synchronize()
cmp(tc, ref, "step1")
for n in ("fit_H1", "fit_H2"):
    h_ref = ref.debug(n).cpu().numpy(); h_tc = tc.debug(n).cpu().numpy()
    print(n, "mask flips:", int(((h_ref > 0) != (h_tc > 0)).sum()), " max|dH|", float(np.abs(h_ref - h_tc).max()), "max|H|", float(np.abs(h_ref).max()))
