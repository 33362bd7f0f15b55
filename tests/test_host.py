"""CPU: host-side logic - flat weight layout, replay row packing, normalisers, logger, sharding helpers, and the
data-parallel gradient-averaging identity over a 2-rank gloo group."""
import os
import pickle

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.sac_eo_oracle import NetCfg, draw_batch, flat, make_problem, sac_eo_update, to_torch_state
from sac_expert_b200 import parallel as P
from sac_expert_b200.population import net_shapes, pack_flat, unpack_flat
from sac_expert_b200.sac_eo.common.logger import Logger
from sac_expert_b200.sac_eo.common.normalizer import RunningNormalizer, RunningNormalizers


def test_flat_layout_round_trip():
    shapes = net_shapes(11, (256, 256), 6)
    rng = np.random.default_rng(0)
    ws = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    back = unpack_flat(pack_flat(ws), shapes)
    assert all(np.array_equal(a, b) for a, b in zip(ws, back))
    assert pack_flat(ws).size == 11 * 256 + 256 + 256 * 256 + 256 + 256 * 6 + 6


def test_normalizer_matches_reference_conventions():
    n = RunningNormalizer(3)
    x = np.array([1.0, 2.0, 3.0], np.float32)
    assert np.array_equal(n.normalize(x), x)                         # identity until updated
    n.instantiate(5, np.array([1, 1, 1], np.float32), np.array([4, 0, 1e-20], np.float32))
    assert np.allclose(n.normalize(x), [(1 - 1) / 2, (2 - 1) / 1e-8, (3 - 1) / 1e-8])      # std floored at 1e-8
    assert np.allclose(n.denormalize(n.normalize(x)), x)
    assert np.allclose(n.normalize(x, center=False), x / np.maximum(n.std, 1e-8))
    r = RunningNormalizer(1)
    assert isinstance(r.mean, float) and r.std == 1.0
    rng = np.random.default_rng(0)
    data = rng.standard_normal((200, 3)) * [1, 2, 3] + [0, 5, -1]
    m = RunningNormalizer(3)
    m.update(data[:50]); m.update(data[50:])
    assert np.allclose(m.mean, data.mean(0), atol=1e-5) and np.allclose(m.var, data.var(0, ddof=1), rtol=1e-5)
    rn = RunningNormalizers(3, 2, 0.99)
    rn.update_rms(data[:10].astype(np.float32), np.zeros((10, 2), np.float32), np.ones(10, np.float32),
                  data[1:11].astype(np.float32))
    assert set(rn.get_rms_stats()) == {"s_rms", "a_rms", "r_rms", "delta_rms", "ret_rms"}


def test_logger_pickle_schema_and_append(tmp_path):
    lg = Logger()
    lg.log_train({"p_loss": 1.0, "alpha_loss": 2.0}); lg.log_train({"p_loss": 3.0})
    lg.log_params({"x": 1}); lg.log_final({"actor_weights": [np.zeros(2)]})
    lg.dump_and_save(str(tmp_path), "ck_0")
    lg.reset(); lg.log_train({"p_loss": 5.0})
    lg.dump_and_save(str(tmp_path), "ck_0")
    out = pickle.load(open(os.path.join(tmp_path, "ck_0"), "rb"))
    assert set(out) == {"param", "train", "final"} and out["train"]["p_loss"].tolist() == [1.0, 3.0, 5.0]


def test_agent_sharding_partitions_the_population():
    for n, w in ((256, 8), (10, 4), (3, 8)):
        shards = [P.agent_shard(n, r, w) for r in range(w)]
        assert sorted(sum(shards, [])) == list(range(n))
        assert P.shard_sizes(n, w) == [len(s) for s in shards]
    noise = np.arange((3 * 8 + 4) * 2, dtype=np.float32).reshape(-1, 2)
    parts = [P.dp_noise_for_rank(noise, 8, 4, r, 2) for r in range(2)]
    assert parts[0].shape == (3 * 4 + 4, 2)
    assert np.array_equal(np.concatenate([parts[0][:4], parts[1][:4]]), noise[:8])
    assert np.array_equal(parts[0][8:12], noise[16:20]) and np.array_equal(parts[1][8:12], noise[16:20])


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = NetCfg(S=5, A=2, actor_hidden=(16, 16), critic_hidden=(16, 16), model_hidden=(16, 16))
    st, replay, expert, hyper = make_problem(cfg, B=16, E=4, N=100, seed=3, perturb=0.05)
    hyper["eps"] = 0.3
    full = draw_batch(cfg, replay, expert, 16, seed=4)
    loc = dict(full)
    for k in ("idx", "s", "a", "sp", "r", "d", "u1", "u2", "u5"):
        loc[k] = P.slice_rows(np.asarray(full[k]), rank, world)        # rows split, expert draw replicated
    o = sac_eo_update(cfg, to_torch_state(st, torch.float64), loc, hyper)
    # phase 1: critic grads averaged over ranks == global-batch gradient (y is per-row, no cross-rank term)
    g = flat(o["g_q1"]).clone()
    P.average_(g)
    if rank == 0:
        ref = sac_eo_update(cfg, to_torch_state(st, torch.float64), full, hyper)
        q.put(float((g - flat(ref["g_q1"])).norm() / flat(ref["g_q1"]).norm()))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_gradient_average_equals_global_batch_gloo():
    """world_size-2 gloo: each rank differentiates its half of ONE global draw normalised by its LOCAL row count;
    the all-reduce average equals the single-process global-batch gradient (the identity dp_update relies on)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    err = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert err < 1e-12
