"""Data-parallel mode check, run under torchrun with >= 2 GPUs:
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py
Every rank holds the SAME agent, B/world rows of ONE global draw; gradients are NCCL-averaged between the
phase-split kernels (parallel.dp_update).  Rank 0 compares parameters after the update with the CPU oracle's
single-process full-batch update."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from oracle.sac_eo_oracle import NetCfg, draw_batch, make_problem, sac_eo_update, to_torch_state
from sac_expert_b200 import parallel as P
from sac_expert_b200.population import Population
from tests.helpers import rel, spec_from_cfg

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
WIDE = os.environ.get("SACEO_DP_WIDE") == "1"     # reference sizes on the tcgen05 engine (tests/test_gpu_tc_parity.py)
if WIDE:
    cfg = NetCfg(S=27, A=8)
    B, E, N = 256, 20, 1200
else:
    cfg = NetCfg(S=11, A=3, actor_hidden=(64, 64), critic_hidden=(64, 64), model_hidden=(64, 64))
    B, E, N = 64, 8, 400
st, replay, expert, hyper = make_problem(cfg, B, E, N, seed=21, perturb=0.05)
hyper["eps"] = 0.25
full = draw_batch(cfg, replay, expert, B, seed=22)
from sac_expert_b200 import lib as _lib
pop = Population(spec_from_cfg(cfg, 1, B // world, E, N, device=local,
                               gemm_mode=_lib.GEMM_TCGEN05_BF16X3 if WIDE else _lib.GEMM_FP32_SIMT))
pop.load_agent(0, st, hyper)
pop.append_rows(0, replay["s"], replay["a"], replay["r"], replay["sp"], replay["d"])
pop.set_expert(0, expert["sE"], expert["spE"])
noise = np.concatenate([full["u1"], full["u2"], full["u3"], full["u4"], full["u5"]]).astype(np.float32)
idx = P.slice_rows(full["idx"], rank, world)
pop.set_draws(idx[None], P.dp_noise_for_rank(noise, B, E, rank, world)[None],
              np.concatenate([full["I1"], full["I2"]]).astype(np.int32)[None])
losses = P.dp_update(pop, 0).cpu().numpy()[0]
torch.cuda.synchronize()
if rank == 0:
    o = sac_eo_update(cfg, to_torch_state(st), full, hyper)
    worst = 0.0
    for name in ("q1", "q2", "t1", "t2", "actor"):
        for got, new, old in zip(pop.get_net(0, name), o["new"][name], st[name]):
            d = new.numpy() - np.asarray(old)
            if np.linalg.norm(d) > 0:
                worst = max(worst, rel(got - np.asarray(old), d))
    la = abs(losses[6] - float(o["new"]["alpha"])) / abs(float(o["new"]["alpha"]))
    lp = abs(losses[4] - float(o["p_loss"])) / abs(float(o["p_loss"]))
    print(f"DP world={world}: worst dtheta rel err {worst:.3e}, alpha rel err {la:.3e}, p_loss rel err {lp:.3e}")
    assert worst < 1e-3 and la < 1e-5 and lp < 1e-4
    print("DP_CHECK_OK")
dist.barrier()
dist.destroy_process_group()
