"""CG / Fisher-vector solve throughput (BASELINE config 4: Humanoid-shaped, N = 1024 states, 20 iterations)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sac_expert_b200.population import Population, PopulationSpec
from sac_expert_b200.synth import fill_synthetic, SHAPES

shape = sys.argv[1] if len(sys.argv) > 1 else "humanoid"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
S, A, B = SHAPES[shape]
N = 1024
pop = Population(PopulationSpec(n_agents=n, S=S, A=A, B=64, E=2, num_models=0, replay_capacity=64, fvp_rows=N, gemm_mode=1))
fill_synthetic(pop, seed=3)
L = pop.L
b = torch.randn(n, L.na_stride, device="cuda") * 0.01
for _ in range(2):
    x, vfv = pop.cg_solve(b, iters=20, tol=1e-10, damp=0.01)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(pop.stream)
reps = 3
for _ in range(reps):
    x, vfv = pop.cg_solve(b, iters=20, tol=1e-10, damp=0.01)
e1.record(pop.stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
Pa = S * 256 + 256 * 256 + 256 * 2 * A
flops = 21 * 2 * N * (2 * Pa + 256 * 256 + 256 * 2 * A) + 2 * N * Pa
print(f"{shape}: {n} agents, N={N}, 20 CG iterations + vFv: {ms:.2f} ms per population solve, {n / (ms * 1e-3):.0f} solves/s, "
      f"{flops * n / (ms * 1e-3) / 1e12:.1f} algorithmic TFLOP/s, vFv[0]={float(vfv[0]):.4e} finite={bool(torch.isfinite(x).all())}")
