"""Per-tensor error of the TRPO surrogate gradient and the Fisher-vector product on the tensor-core engine vs the fp64
oracle at the reference sizes (diagnostic; run on a B200: python tools/trpo_tc_check.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import sac_eo_oracle as O
from sac_expert_b200 import lib
from tests.test_gpu_trpo import _setup
from tests.helpers import rel

for mode, perturb in ((lib.GEMM_TCGEN05_BF16X3, 0.2), (lib.GEMM_FP32_SIMT, 0.2), (lib.GEMM_TCGEN05_BF16X3, 0.02), (lib.GEMM_FP32_SIMT, 0.02)):
    cfg, pop, probs = _setup(True, ("tanh", "tanh"), gemm_mode=mode, hidden=(256, 256), S=27, A=8, N=256, perturb=perturb)
    L = pop.L
    act = np.stack([p[2] for p in probs]); adv = np.stack([O.trpo_normalise_adv(p[3]) for p in probs]).astype(np.float32)
    g, _ = pop.trpo_grad(act, adv, None, None)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, L.na_stride)).astype(np.float32); x[:, L.na:] = 0
    Fx = pop.fvp(torch.from_numpy(x), 0.01).cpu().numpy()
    g = g.cpu().numpy()
    sizes = [27 * 256, 256, 65536, 256, 256 * 16, 16]
    for i, (st, s, a, _) in enumerate(probs):
        th = O.to_torch_state(st, torch.float64)
        e0 = O.trpo_eval(cfg, th["actor"], s, a, adv[i], np.zeros(len(s)), None, th)
        gr, _, _ = O.trpo_surrogate_grad(cfg, th["actor"], s, a, adv[i], e0["nlp"], 0.0, 0.0, th)
        gr = O.flat(gr).numpy()
        Fr = O.make_F(cfg, th["actor"], s, th, damp=0.01)(torch.from_numpy(x[i, :L.na].astype(np.float64))).numpy()
        o = 0; parts = []
        for n_ in sizes:
            parts.append(("%.1e/%.1e" % (rel(g[i, o:o + n_], gr[o:o + n_]), rel(Fx[i, o:o + n_], Fr[o:o + n_]))))
            o += n_
        print("mode", mode, "perturb", perturb, "agent", i, "grad/fvp rel per tensor [W0 b0 W1 b1 W2 b2]:", " ".join(parts),
              "| total %.1e/%.1e" % (rel(g[i, :L.na], gr), rel(Fx[i, :L.na], Fr)), flush=True)
    pop.close()
