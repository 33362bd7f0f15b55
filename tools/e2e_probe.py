"""Where the end-to-end step time goes: host index generation vs staging copies vs device time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sac_expert_b200.population import Population, PopulationSpec
from sac_expert_b200.synth import fill_synthetic
n, B = 256, 256
pop = Population(PopulationSpec(n_agents=n, S=27, A=8, B=B, E=20, num_models=2, replay_capacity=100000, gemm_mode=1))
fill_synthetic(pop, seed=1)
rng = np.random.default_rng(7)
sizes = pop._host_size
expert = np.stack([pop.t["expert_s"].cpu().numpy(), pop.t["expert_sp"].cpu().numpy()], 0)
def timeit(f, k=20):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(k): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3
print("rng.integers broadcast  %.2f ms" % timeit(lambda: rng.integers(0, sizes[:, None], size=(n, B)).astype(np.int64)))
print("rng.integers scalar     %.2f ms" % timeit(lambda: rng.integers(0, int(sizes[0]), size=(n, B))))
idx = rng.integers(0, sizes[:, None], size=(n, B)).astype(np.int64)
print("update_host sync        %.2f ms" % timeit(lambda: pop.update_host(0, 5, idx, expert)))
print("update_host sync no exp %.2f ms" % timeit(lambda: pop.update_host(0, 5, idx, None)))
print("update (device rng)     %.2f ms" % timeit(lambda: pop.update(1, 0, True, 5)))
st = [0]
def pipelined():
    s = st[0]; st[0] += 1
    pop.update_host_async(s, 5, idx, expert, slot=s & 1)
    if s: pop.wait_host((s - 1) & 1)
print("async precomputed idx   %.2f ms" % timeit(pipelined))
def pipelined_gen():
    s = st[0]; st[0] += 1
    ii = rng.integers(0, sizes[:, None], size=(n, B)).astype(np.int64)
    pop.update_host_async(s, 5, ii, expert, slot=s & 1)
    pop.wait_host((s - 1) & 1)
print("async + gen broadcast   %.2f ms" % timeit(pipelined_gen))
def pipelined_gen2():
    s = st[0]; st[0] += 1
    ii = rng.integers(0, int(sizes[0]), size=(n, B))
    pop.update_host_async(s, 5, ii, expert, slot=s & 1)
    pop.wait_host((s - 1) & 1)
print("async + gen scalar      %.2f ms" % timeit(pipelined_gen2))
