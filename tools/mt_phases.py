"""Per-CTA phase times of the tensor-core expert-term kernel (k_model_term_mma) inside one real un-graphed update.
usage: mt_phases.py [n_agents]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from sac_expert_b200 import lib as L
from sac_expert_b200.population import Population, PopulationSpec
from sac_expert_b200.synth import fill_synthetic
lib = L.load()
lib.saceo_test_set_mt_debug.argtypes = [C.c_void_p]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pop = Population(PopulationSpec(n_agents=n, S=27, A=8, B=256, E=20, replay_capacity=2000, gemm_mode=1, use_graph=False))
fill_synthetic(pop, seed=1)
for w in range(3):
    pop.update(1, num_timesteps=w, use_device_rng=True, seed=3)
torch.cuda.synchronize()
dbg = torch.zeros(2 * n * 8, dtype=torch.int64, device="cuda")
lib.saceo_test_set_mt_debug(dbg.data_ptr())
pop.update(1, num_timesteps=5, use_device_rng=True, seed=3)
torch.cuda.synchronize()
lib.saceo_test_set_mt_debug(None)
t = dbg.cpu().numpy().reshape(2 * n, 8).astype(np.float64)
names = ["layer 0", "layer 1 (mma)", "layer 2", "loss", "layer 2^T", "layer 1^T (mma)", "layer 0^T"]
print("CTAs %d  span %.1f us  CTA mean %.2f us" % (len(t), (t[:, 7].max() - t[:, 0].min()) / 1e3, (t[:, 7] - t[:, 0]).mean() / 1e3))
print(" | ".join("%s %.2f" % (names[i], (t[:, i + 1] - t[:, i]).mean() / 1e3) for i in range(7)))
