"""One update of the bench population between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --set full --clock-control none -o gpurun_out/r2_step python tools/step_traffic.py [n]
The kernels are the ones bench.py times (same spec, same synthetic tables); un-graphed so that every launch is a plain
kernel launch for the profiler.  tools/traffic_from_ncu.py turns the raw CSV page into profiles/r2_step_traffic.json."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sac_expert_b200.population import Population, PopulationSpec
from sac_expert_b200.synth import fill_synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pop = Population(PopulationSpec(n_agents=n, S=27, A=8, B=256, E=20, num_models=2, replay_capacity=100000, gemm_mode=1,
                                use_graph=False))
fill_synthetic(pop, seed=1234)
for w in range(4):
    pop.update(1, num_timesteps=w, use_device_rng=True, seed=3)
torch.cuda.synchronize()
torch.cuda.profiler.start()
pop.update(1, num_timesteps=4, use_device_rng=True, seed=3)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(pop.losses.sum()))
