"""Per-launch device times of one un-graphed single-stream update (saceo_profile_step), every launch listed."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sac_expert_b200.population import Population, PopulationSpec
from sac_expert_b200.synth import fill_synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pop = Population(PopulationSpec(n_agents=n, S=27, A=8, B=256, E=20, replay_capacity=2000, gemm_mode=1))
fill_synthetic(pop, seed=1)
for w in range(3):
    prof = pop.profile_step(w, True, seed=9)
tot = sum(us for _, us in prof)
for name, us in prof:
    print("%-24s %8.1f" % (name, us))
print("total %.1f us, %d launches" % (tot, len(prof)))
