#!/usr/bin/env python
"""Population sweep (BASELINE configs[4]): device-timed agent-updates/s of the SAC-EO update for n_agents in
{1, 2, 4, ..., 4096} PER GPU, as absolute rate and as fraction of the HBM roofline, on 1 GPU or under torchrun on N
(every rank holds its own n agents: weak scaling; value = all ranks' agents / max-over-ranks device time).

    python tools/pop_sweep.py [--shape ant] [--max 4096] [--steps 10] > gpurun_out/sweep.json
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pop_sweep.py ...
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench
    from sac_expert_b200 import lib
    from sac_expert_b200.population import Population, PopulationSpec
    from sac_expert_b200.synth import SHAPES, algorithmic_bytes, fill_synthetic
    p = argparse.ArgumentParser()
    p.add_argument("--shape", default="ant")
    p.add_argument("--max", type=int, default=4096)
    p.add_argument("--min", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--replay-rows", type=int, default=20000)
    p.add_argument("--plain-sac", action="store_true")
    p.add_argument("--no-fork", action="store_true")
    a = p.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    S, A, B = SHAPES[a.shape]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rows = []
    n = a.min
    while n <= a.max:
        spec = PopulationSpec(n_agents=n, S=S, A=A, B=B, E=20, num_models=0 if a.plain_sac else 2,
                              replay_capacity=a.replay_rows, gemm_mode=lib.GEMM_TCGEN05_BF16X3, fork_actor=not a.no_fork, device=local)
        pop = Population(spec)
        fill_synthetic(pop, seed=77 + rank)
        steps = max(a.steps, min(200, 2000 // max(n, 1)))            # small populations: more steps per timing
        ms, launches = bench.measure(pop, a, world, barrier, steps, a.warmup, torch, dist)
        per_agent = algorithmic_bytes(spec, pop.L)
        value = world * n * steps / (ms * 1e-3)
        rows.append({"agents_per_gpu": n, "n_gpus": world, "value": value, "ms_per_step": ms / steps, "steps": steps,
                     "us_per_agent_update_per_gpu": ms / steps * 1e3 / n,
                     "hbm_roofline_frac": per_agent * n / (ms / steps * 1e-3) / 1e9 / hbm,
                     "l2_resident": bool(per_agent * n < 126e6)})
        pop.close()
        del pop
        torch.cuda.empty_cache()
        n *= 2
    if rank == 0:
        print(json.dumps({"metric": "agent-updates/sec", "unit": "agent-updates/s", "shape": a.shape, "n_gpus": world,
                          "workload": "SAC-EO" if not a.plain_sac else "plain SAC", "batch": B, "replay_rows": a.replay_rows,
                          "hbm_peak_GBs": hbm, "second_stream_branch": not a.no_fork,
                          "note": "weak scaling per row: every GPU holds agents_per_gpu agents; populations whose state fits the "
                                  "126 MB L2 (l2_resident) are not HBM-bound - their roofline fraction is reported for continuity only",
                          "rows": rows}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
