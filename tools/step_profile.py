"""Per-launch device times of one update (saceo_profile_step), any shape:
    python tools/step_profile.py [shape] [agents] [gemm_mode]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sac_expert_b200.population import Population, PopulationSpec
from sac_expert_b200.synth import SHAPES, fill_synthetic
shape = sys.argv[1] if len(sys.argv) > 1 else "ant"
agents = int(sys.argv[2]) if len(sys.argv) > 2 else 256
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
S, A, B = SHAPES[shape]
pop = Population(PopulationSpec(n_agents=agents, S=S, A=A, B=B, E=20, num_models=2, replay_capacity=20000, gemm_mode=mode))
fill_synthetic(pop, seed=1)
for w in range(3):
    prof = pop.profile_step(w, True, seed=3)
tot = sum(u for _, u in prof)
for i, (n, u) in enumerate(prof):
    print("%3d %-18s %9.1f us %5.1f%%" % (i, n, u, 100 * u / tot))
print("total %.1f us, %d launches" % (tot, len(prof)))
