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


def test_population_line_search_equals_the_per_agent_reference_flow():
    """``backtrack_population`` (vectorised accept / shrink decisions of Population.trpo_update) against the restated
    single-agent ``TRPO._backtrack`` control flow (oracle.backtrack, itself pinned on the reference's logged
    trajectory) on synthetic per-agent (kl, improve) tables: accept at any level, never accept, mixed populations."""
    import json
    import os
    from oracle import sac_eo_oracle as O
    from sac_expert_b200.population import backtrack_population
    rng = np.random.default_rng(0)
    delta, klf = 0.02, 1.5
    for rep in range(30):
        n = int(rng.integers(1, 9))
        # level k = number of sqrt(2) shrinks (0..10); level 11 = restored parameters (adj 0)
        kl = rng.uniform(0.0, 0.09, size=(n, 12)) * (0.75 ** np.arange(12))[None]
        imp = rng.normal(0.01, 0.02, size=(n, 12))
        tv = rng.uniform(0, 0.2, size=(n, 12))
        if rep % 3 == 0:
            kl[0] = 1.0                                    # an agent no shrink can save
        kl[:, 11] = 0.0; imp[:, 11] = 0.0; tv[:, 11] = 0.0
        level = lambda a: 11 if a == 0 else int(round(-2 * np.log2(a)))
        calls = []

        def trial(adj):
            calls.append(adj.copy())
            k = [level(a) for a in adj]
            st = np.stack([[0.0, kl[i, k[i]], tv[i, k[i]]] for i in range(n)])
            return st, np.array([imp[i, k[i]] for i in range(n)])

        adj, stats, improve, tv_pre, kl_pre = backtrack_population(trial, n, klf, delta)
        for i in range(n):
            def trial_i(step, i=i):
                k = level(step)
                return None, dict(kl=kl[i, k], tv=tv[i, k]), imp[i, k]
            _, e, im, adj_i, step, tvp, klp = O.backtrack(trial_i, 1.0, klf, delta)
            assert adj[i] == adj_i, (rep, i, adj, adj_i)
            assert (tv_pre[i], kl_pre[i]) == (tvp, klp)
            k = level(adj[i])
            assert stats[i, 1] == kl[i, k] and stats[i, 2] == tv[i, k] and improve[i] == imp[i, k]
            if adj_i != 0:
                assert stats[i, 1] == e["kl"] and improve[i] == im and step == adj_i
        assert len(calls) <= 12
    # the reference's own logged trajectory (tests/golden/templog0_trpo_log.json), both updates as a 2-agent population
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "templog0_trpo_log.json")))

    def trial_ref(adj):
        st = np.zeros((2, 3)); im = np.array(ref["improve"])
        for u in range(2):
            first = adj[u] == 1.0
            st[u, 1] = ref["kl_pre"][u] if first else ref["kl"][u]
            st[u, 2] = ref["tv_pre"][u] if first else ref["tv"][u]
        return st, im

    adj, stats, _, tv_pre, kl_pre = backtrack_population(trial_ref, 2, ref["hyper"]["kl_maxfactor"], ref["hyper"]["delta_trpo"])
    assert list(adj) == ref["adj"] and list(stats[:, 1]) == ref["kl"] and list(stats[:, 2]) == ref["tv"]
    assert list(kl_pre) == ref["kl_pre"] and list(tv_pre) == ref["tv_pre"]


@pytest.mark.parametrize("which", ["trpo", "ppo"])
def test_on_policy_temperature_step_is_keras_adam_with_the_zero_floor(which):
    """``alpha_optimizer.apply_gradients([alpha_grad * -1])`` + ``alpha = max(alpha, 0)`` (trpo.py:169-172,
    ppo.py:221-224): the host scalar step of the mirror classes against the oracle's Keras-Adam restatement."""
    from oracle.sac_eo_oracle import keras_adam
    from sac_expert_b200.sac_eo.algs.model_free.ppo import PPO
    from sac_expert_b200.sac_eo.algs.model_free.trpo import TRPO
    kw = dict(adv_center=True, adv_scale=True, delta_trpo=0.02, cg_it=5, trust_sub=1, trust_damp=0.01, kl_maxfactor=1.5,
              ent_reg=True, ent_targ=-1.0, alpha_lr=0.01, actor_lr=3e-4, actor_update_it=2, actor_nminibatch=3,
              eps_ppo=0.2, max_grad_norm=0.5, adaptlr=True, adapt_factor=0.03, adapt_minthresh=0.0, adapt_maxthresh=1.0)
    alg = (TRPO if which == "trpo" else PPO)(None, kw)          # constructing the update object needs no device
    a, m, v, t = [torch.zeros((), dtype=torch.float64)], [torch.zeros((), dtype=torch.float64)], \
        [torch.zeros((), dtype=torch.float64)], 0
    grads = [-0.8, -1.1, -0.3, 0.5, 2.0, 2.5, -0.1]             # first pushes alpha up, later ones drive it into the floor
    for g in grads:
        alg._alpha_step(g)
        a, m, v, t = keras_adam(a, [torch.tensor(g, dtype=torch.float64)], m, v, t, 0.01)
        a = [torch.clamp(a[0], min=0.0)]
        assert abs(float(alg.alpha) - float(a[0])) < 2e-6, (g, alg.alpha, a)
    assert alg._alpha_t == len(grads) and float(alg.alpha) >= 0.0


@pytest.mark.parametrize("buffer_size", [None, 50, 7])
def test_trajectory_buffer_window_equals_concatenate_semantics(buffer_size):
    """TrajectoryBuffer.add without the reference's O(N) re-concatenation (buffers.py:41-71): the exposed arrays must be
    exactly what np.concatenate + tail truncation (:60-66) produces, after every add, for bounded and unbounded buffers."""
    from sac_expert_b200.sac_eo.common.buffers import TrajectoryBuffer
    rng = np.random.default_rng(0)
    tb = TrajectoryBuffer(3, 2, 0.99, 0.95, buffer_size)
    S, R, D, I = np.empty((0, 3), np.float32), np.empty((0,), np.float32), np.empty((0,)), np.empty((0,))
    for t in range(300):
        k = int(rng.integers(1, 9))
        s, a = rng.standard_normal((k, 3)).astype(np.float32), rng.standard_normal((k, 2)).astype(np.float32)
        r, d = rng.standard_normal(k).astype(np.float32), (rng.random(k) < 0.1).astype(np.float64)
        tb.add(s, a, r, s + 1, d)
        S, R, D, I = np.concatenate((S, s)), np.concatenate((R, r)), np.concatenate((D, d)), np.concatenate((I, np.ones(k) * t))
        if buffer_size and len(R) > buffer_size:
            S, R, D, I = S[-buffer_size:], R[-buffer_size:], D[-buffer_size:], I[-buffer_size:]
        assert tb.current_size == len(R) and tb.steps_total == sum(1 for _ in range(1)) * tb.steps_total
        assert np.array_equal(tb.s_all, S) and np.array_equal(tb.sp_all, S + 1) and np.array_equal(tb.r_all, R)
        assert np.array_equal(tb.d_all, D) and tb.d_all.dtype == np.float64 and np.array_equal(tb.idx_all, I)
    s_, a_, sp_, r_ = tb.get_model_info()
    assert s_.shape == (len(R), 3) and a_.shape == (len(R), 2)


def test_train_parser_groups_and_seed_derivation():
    """The kwarg groups `train()` consumes carry the reference's flag names and defaults
    (/root/reference/sac_eo/common/train_parser.py) and the per-run seeds follow train.py:116-147."""
    from sac_expert_b200.sac_eo import train as T
    from sac_expert_b200.sac_eo.common.train_parser import all_kwargs, create_train_parser
    args = create_train_parser().parse_args(["--env_type", "synthetic", "--env_name", "ant", "--actor_squash", "--runs", "3",
                                             "--no_model_batch_shuffle", "--alg_seed", "5"])
    inputs = T.build_inputs(args)
    assert len(inputs) == 3 and [i["setup_kwargs"]["idx"] for i in inputs] == [0, 1, 2]
    a = inputs[0]["alg_kwargs"]
    assert (a["gamma"], a["soft_tau"], a["sac_batch_size"], a["epsilon"], a["expert_buffer_size"], a["alg_type"]) == \
        (0.995, 5e-3, 256, 1e-3, 20, "sac_imit")
    assert a["model_batch_shuffle"] is False and a["init_rms_stats"] is None and a["save_path"] == "./logs"
    assert inputs[0]["model_kwargs"]["model_layers"] == [512, 512] and inputs[0]["critic_kwargs"]["num_models"] == 2
    seeds = np.random.SeedSequence(0).generate_state(5)
    assert inputs[1]["setup_kwargs"]["sim_seed"] == int(np.random.SeedSequence(seeds[1]).generate_state(3)[1])
    assert all(i["setup_kwargs"]["algorithm_seed"] == 5 for i in inputs)
    assert set(all_kwargs) == {"setup_kwargs", "env_kwargs", "actor_kwargs", "critic_kwargs", "model_kwargs",
                               "model_setup_kwargs", "alg_kwargs", "mf_update_kwargs"}
