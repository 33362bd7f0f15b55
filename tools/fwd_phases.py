"""Per-CTA phase timing of the fused forward kernel (actor inference pass over 256 agents x 256 rows)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from sac_expert_b200 import lib as L
from sac_expert_b200.population import Population, PopulationSpec
from sac_expert_b200.synth import fill_synthetic
lib = L.load()
lib.saceo_test_set_tc_debug.argtypes = [C.c_void_p]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 27
A = int(sys.argv[3]) if len(sys.argv) > 3 else 8
pop = Population(PopulationSpec(n_agents=n, S=S, A=A, B=256, E=20, replay_capacity=1000, gemm_mode=1))
fill_synthetic(pop, seed=1)
obs = torch.randn(n, 256, S, device="cuda")
for _ in range(3):
    pop.actor_forward(obs, None)
torch.cuda.synchronize()
nct = n * 2
dbg = torch.zeros(2 * nct * 8, dtype=torch.int64, device="cuda")
lib.saceo_test_set_tc_debug(dbg.data_ptr())
pop.actor_forward(obs, None)
torch.cuda.synchronize()
lib.saceo_test_set_tc_debug(None)
raw = dbg.cpu().numpy()
t = raw[:nct * 8].reshape(nct, 8).astype(np.float64)

names = ["setup+L0 (X,W0 load, MMA)", "epilogue0", "misc", "L1 stream (8 slabs)", "epilogue1 + W2 stage", "L2 MMA", "epilogue2"]
d = np.diff(t[:, :7], axis=1)
print("kernel span %.1f us, CTA mean %.2f us" % ((t[:, 6].max() - t[:, 0].min()) / 1e3, (t[:, 6] - t[:, 0]).mean() / 1e3))
for i, nme in enumerate(["setup+L0", "epilogue0", "L1 stream (8 slabs)", "epilogue1+W2 stage", "L2 MMA", "epilogue2"]):
    print("  %-22s mean %7.2f us  p10 %7.2f  p90 %7.2f" % (nme, d[:, i].mean() / 1e3, np.percentile(d[:, i], 10) / 1e3, np.percentile(d[:, i], 90) / 1e3))
