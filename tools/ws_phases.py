"""Per-CTA phase timing of the warp-specialised fused kernels inside one real (un-graphed) update: every fused launch
of the step is stamped in turn (saceo_test_set_ws_debug) and summarised.  usage: ws_phases.py [n_agents] [S] [A]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from sac_expert_b200 import lib as L
from sac_expert_b200.population import Population, PopulationSpec
from sac_expert_b200.synth import fill_synthetic
lib = L.load()
lib.saceo_test_set_ws_debug.argtypes = [C.c_void_p, C.c_int32]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 27
A = int(sys.argv[3]) if len(sys.argv) > 3 else 8
pop = Population(PopulationSpec(n_agents=n, S=S, A=A, B=256, E=20, replay_capacity=2000, gemm_mode=1, use_graph=False, fork_actor=False))   # one stream: the launch order below
fill_synthetic(pop, seed=1)
for w in range(3):
    pop.update(1, num_timesteps=w, use_device_rng=True, seed=3)
torch.cuda.synchronize()
NC = 3 * 2 * n
labels = ["fwd actor(sp)", "fwd Qt(sp,a')", "fwd Q(s,a)+save", "bwd Q (grads)", "fwd actor(s,sE)+save", "fwd Q(s,pi)+save",
          "bwd Q (dXa)", "bwd actor (grads)", "fwd actor(s) alpha"]
for sel in range(9):
    dbg = torch.zeros(NC * 16, dtype=torch.int64, device="cuda")
    lib.saceo_test_set_ws_debug(dbg.data_ptr(), sel)
    pop.update(1, num_timesteps=5, use_device_rng=True, seed=3)
    torch.cuda.synchronize()
    lib.saceo_test_set_ws_debug(None, 0)
    t = dbg.cpu().numpy().reshape(NC, 16).astype(np.float64)
    t = t[t[:, 0] > 0]
    if not len(t):
        print(sel, "no stamps"); continue
    span = (t[:, 7].max() - t[:, 0].min()) / 1e3
    bwd = labels[sel].startswith("bwd")
    def m(a, b):
        ok = (t[:, a] > 0) & (t[:, b] > 0)
        return ((t[ok, b] - t[ok, a]).mean() / 1e3) if ok.any() else float("nan")
    print("%-24s CTAs %4d span %7.1f us  CTA mean %6.2f us" % (labels[sel], len(t), span, m(0, 7)))
    if not bwd:
        print("    X convert %.2f | wait D0 %.2f | epi0 %.2f | W2 stage+bar %.2f | wait D1 %.2f | epi1 %.2f | reduce/out %.2f || MMA: L0 done at %.2f, last a_part at %.2f, L1 done at %.2f, TMA issued at %.2f"
              % (m(0, 1), m(1, 2), m(2, 3), m(3, 4), m(4, 5), m(5, 6), m(6, 7), m(0, 8), m(0, 9), m(0, 10), m(0, 11)))
    else:
        print("    W2 stage+scale %.2f | epi0 %.2f | W0a stage %.2f | wait D(h0) %.2f | epi1(h0) %.2f | tail %.2f || D half0 at %.2f, half1 at %.2f; warp4: wait until %.2f, epi1 done %.2f"
              % (m(0, 1), m(1, 2), m(2, 3), m(3, 4), m(4, 5), m(5, 7), m(0, 8), m(0, 9), m(0, 12), m(0, 13)))
