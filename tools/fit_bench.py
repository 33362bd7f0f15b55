"""Throughput of dynamics-model fitting: agent-fit-steps/s for a population, CUDA-event timed,
with the HBM roofline of the step and the oracle (torch CPU) timed beside it.
    python tools/fit_bench.py [--agents 256] [--shape ant] [--steps 20] [--gemm-mode 1]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sac_expert_b200 import lib as L
from sac_expert_b200.population import Population, PopulationSpec
from sac_expert_b200.synth import SHAPES, fill_synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--agents", type=int, default=256)
ap.add_argument("--shape", default="ant")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--mb", type=int, default=200)
ap.add_argument("--rows", type=int, default=20000)
ap.add_argument("--gemm-mode", type=int, default=L.GEMM_TCGEN05_BF16X3)
ap.add_argument("--clip", type=float, default=0.0)
ap.add_argument("--cpu-seconds", type=float, default=10.0)
a = ap.parse_args()
S, A = SHAPES[a.shape][:2]
spec = PopulationSpec(n_agents=a.agents, S=S, A=A, B=256, E=20, num_models=2, replay_capacity=a.rows, gemm_mode=a.gemm_mode)
pop = Population(spec)
fill_synthetic(pop, seed=0, replay_rows=a.rows)
pop.fit_bind(a.mb, use_grad_clip=a.clip > 0)
if a.clip > 0:
    pop.t["fit_hyper"][:, 4] = a.clip
g = torch.Generator(device="cuda").manual_seed(0)
idx = torch.randint(0, a.rows, (a.steps, a.agents, 2, a.mb), device="cuda", generator=g)
pop.model_fit(idx[: a.warmup], want_losses=False)
torch.cuda.synchronize()
l0 = pop.launches
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(pop.stream):
    e0.record(pop.stream)
    losses = pop.model_fit(idx)
    e1.record(pop.stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
nm = pop.L.nm
alg_bytes = 2 * nm * 4 * 7 + 2 * a.mb * pop.L.row_words * 4          # theta,m,v read+write, grad read; minibatch rows
flops = 2 * 3 * 2 * a.mb * (nm - 2 * 512 - (S + 1))                  # fwd + dX + dW per weight, 2 models
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
peak = float(peaks.get("hbm_gbs", 6553.3)) if isinstance(peaks, dict) else 6553.3
ach = alg_bytes * a.agents / (ms * 1e-3) / 1e9
out = dict(metric="agent-fit-steps/sec", value=a.agents / (ms * 1e-3), ms_per_step=ms, agents=a.agents, shape=a.shape,
           model_batch=a.mb, loss_first=float(losses[0].mean()), loss_last=float(losses[-1].mean()),
           launches_per_step=(pop.launches - l0) / a.steps,
           roofline=dict(bound="hbm", achieved=ach, peak=peak, unit="GB/s", frac=ach / peak,
                         algorithmic_bytes_per_agent_step=alg_bytes, algorithmic_tflops=flops * a.agents / (ms * 1e-3) / 1e12))
# CPU baseline: the oracle's apply_model_grads on one agent
if a.cpu_seconds > 0:
    from oracle.sac_eo_oracle import NetCfg, apply_model_grads, make_problem, to_torch_state
    cfg = NetCfg(S=S, A=A)
    st, replay, _, _ = make_problem(cfg, 32, 4, 4000, seed=0)
    T = to_torch_state(st)
    models = [T["m1"], T["m2"]]
    adam = dict(m=[[torch.zeros_like(w) for w in m] for m in models], v=[[torch.zeros_like(w) for w in m] for m in models], t=0)
    rng = np.random.default_rng(0)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < a.cpu_seconds:
        ii = rng.integers(0, 4000, (2, a.mb))
        b = [{k: torch.as_tensor(replay[k][ii[m]]) for k in ("s", "a", "sp", "r")} for m in range(2)]
        o = apply_model_grads(cfg, models, adam, b, T, dict(model_lr=1e-3))
        models, adam = o["models"], dict(m=o["m"], v=o["v"], t=o["t"])
        n += 1
    out["cpu_baseline"] = dict(value=n / (time.perf_counter() - t0), unit="agent-fit-steps/s", cores=torch.get_num_threads(),
                               kind="port", sample="%d oracle steps, 1 agent" % n)
print(json.dumps(out))
