"""Per-tensor gradient deviation of one model-fit step vs the oracle (both GEMM engines)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.sac_eo_oracle import NetCfg, apply_model_grads, make_problem, model_fit_batches, to_torch_state
from sac_expert_b200 import lib as L
from sac_expert_b200.population import Population, unpack_flat
from tests.helpers import rel, spec_from_cfg

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for acts in (("relu", "relu"), ("tanh", "tanh")):
    for mode, name in ((L.GEMM_FP32_SIMT, "fp32"), (L.GEMM_TCGEN05_BF16X3, "tc")):
        cfg = NetCfg(S=11, A=3, model_acts=acts)
        pop = Population(spec_from_cfg(cfg, 1, 32, 4, 1000, gemm_mode=mode))
        pop.fit_bind(mb)
        st, replay, expert, hyper = make_problem(cfg, 32, 4, 1000, seed=21, perturb=0.05)
        pop.load_agent(0, st, hyper)
        pop.append_rows(0, replay["s"], replay["a"], replay["r"], replay["sp"], replay["d"])
        idx = model_fit_batches(1000, 2, mb, True, np.random.default_rng(9))[0][None]
        losses = pop.model_fit(idx)
        torch.cuda.synchronize()
        T = to_torch_state(st)
        T64 = to_torch_state(st, torch.float64)
        models = [T["m1"], T["m2"]]
        adam = dict(m=[[torch.zeros_like(w) for w in m] for m in models], v=[[torch.zeros_like(w) for w in m] for m in models], t=0)
        b = [{k: torch.as_tensor(replay[k][idx[0, m]]) for k in ("s", "a", "sp", "r")} for m in range(2)]
        o = apply_model_grads(cfg, models, adam, b, T, dict())
        m64 = [T64["m1"], T64["m2"]]
        adam64 = dict(m=[[torch.zeros_like(w) for w in m] for m in m64], v=[[torch.zeros_like(w) for w in m] for m in m64], t=0)
        o64 = apply_model_grads(cfg, m64, adam64, b, T64, dict())
        g = pop.debug("g_model").cpu().numpy().reshape(2, pop.L.nm_stride)
        out = pop.debug("fit_Out").cpu().numpy()
        for m in range(2):
            gs = unpack_flat(g[m], pop.shapes["model"])
            print(acts[0], name, "model", m, "loss", float(losses[0, 0, m]), float(o["losses"][m]),
                  " ".join("%s=%.1e/%.1e" % (nm_, rel(a, r.numpy()), rel(r32.numpy(), r.numpy()))
                           for nm_, a, r, r32 in zip(("W0", "b0", "W1", "b1", "W2", "b2"), gs, o64["grads"][m], o["grads"][m])))
        pop.close()
