"""Runs the README quick start."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sac_expert_b200.population import Population, PopulationSpec
from sac_expert_b200.synth import fill_synthetic
pop = Population(PopulationSpec(n_agents=256, S=27, A=8))
fill_synthetic(pop, seed=0)
losses = pop.update(n_steps=100, num_timesteps=0, use_device_rng=True, seed=1)
pop.fit_bind(model_batch=200)
idx = torch.randint(0, 100_000, (50, 256, 2, 200), device="cuda")
fit_losses = pop.model_fit(idx)
act = pop.actor_forward(torch.randn(256, 1, 27, device="cuda"))
torch.cuda.synchronize()
print(losses.shape, fit_losses.shape, (act[0] if isinstance(act, tuple) else act).shape, float(fit_losses[0].mean()), float(fit_losses[-1].mean()))
