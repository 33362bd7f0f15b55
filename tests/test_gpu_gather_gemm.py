"""GPU: bit-exact replay gather (buffers.py:126-144) and the standalone batched GEMM surface."""
import numpy as np
import pytest
import torch

from oracle.sac_eo_oracle import NetCfg, gather as oracle_gather, make_problem
from sac_expert_b200 import lib as L
from sac_expert_b200.population import Population
from tests.helpers import spec_from_cfg

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("S,A,B", [(11, 3, 256), (17, 6, 64), (27, 8, 33), (376, 17, 128), (1, 1, 1)])
def test_gather_bit_exact(S, A, B):
    cfg = NetCfg(S=S, A=A, actor_hidden=(8, 8), critic_hidden=(8, 8), num_models=0)
    n, N = 3, 777
    pop = Population(spec_from_cfg(cfg, n, B, 0, N))
    reps, idxs = [], []
    rng = np.random.default_rng(S)
    for i in range(n):
        _, replay, _, _ = make_problem(cfg, B, 2, N, seed=i)
        # arbitrary f64 payload in d, not just 0/1, and odd bit patterns in the floats
        replay["d"] = rng.standard_normal(N)
        replay["r"][::7] = np.float32(-0.0)
        replay["s"][::5, 0] = np.nextafter(np.float32(1), np.float32(2))
        pop.append_rows(i, replay["s"], replay["a"], replay["r"], replay["sp"], replay["d"])
        reps.append(replay)
        idxs.append(rng.integers(0, N, size=B))
    idxs[0][:] = N - 1            # max index, all duplicates (sampling is with replacement)
    idx = np.stack(idxs).astype(np.int64)
    s, a, sp, r, d = [t.cpu().numpy() for t in pop.gather(torch.from_numpy(idx))]
    for i in range(n):
        es, ea, esp, er, ed = oracle_gather(reps[i], idx[i])
        assert s[i].tobytes() == es.tobytes()
        assert a[i].tobytes() == ea.tobytes()
        assert sp[i].tobytes() == esp.tobytes()
        assert r[i].tobytes() == er.tobytes()
        assert d[i].dtype == np.float64 and d[i].tobytes() == ed.tobytes()


def test_gather_ring_truncation():
    """TrajectoryBuffer.add keeps the LAST buffer_size rows (buffers.py:60-66); logical index 0 is the oldest."""
    cfg = NetCfg(S=4, A=2, actor_hidden=(8, 8), critic_hidden=(8, 8), num_models=0)
    cap, B = 100, 50
    pop = Population(spec_from_cfg(cfg, 1, B, 0, cap))
    rng = np.random.default_rng(0)
    ref = {k: np.empty((0,) + sh, dt) for k, sh, dt in (("s", (4,), np.float32), ("a", (2,), np.float32),
                                                       ("sp", (4,), np.float32), ("r", (), np.float32),
                                                       ("d", (), np.float64))}
    for chunk in (30, 50, 45, 70, 1, 120):
        new = dict(s=rng.standard_normal((chunk, 4)).astype(np.float32), a=rng.standard_normal((chunk, 2)).astype(np.float32),
                   sp=rng.standard_normal((chunk, 4)).astype(np.float32), r=rng.standard_normal(chunk).astype(np.float32),
                   d=(rng.random(chunk) < 0.3).astype(np.float64))
        pop.append_rows(0, new["s"], new["a"], new["r"], new["sp"], new["d"])
        for k in ref:
            ref[k] = np.concatenate([ref[k], new[k]])[-cap:]
        size = len(ref["r"])
        idx = rng.integers(0, size, size=(1, B)).astype(np.int64)
        s, a, sp, r, d = [t.cpu().numpy()[0] for t in pop.gather(torch.from_numpy(idx))]
        es, ea, esp, er, ed = oracle_gather(ref, idx[0])
        assert np.array_equal(s, es) and np.array_equal(a, ea) and np.array_equal(sp, esp)
        assert np.array_equal(r, er) and np.array_equal(d, ed)


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(64, 64, 16), (276, 256, 11), (10, 512, 35), (257, 1, 256), (33, 7, 129),
                                   (20, 256, 256), (256, 16, 64), (1, 300, 70), (300, 8, 256), (32, 32, 32), (40, 33, 100),
                                   (36, 256, 256), (36, 256, 276), (63, 100, 70), (12, 512, 200), (49, 64, 31)])
def test_simt_gemm(ta, tb, M, N, K):
    lib = L.load()
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    batch = 3
    A = torch.randn(batch, *((K, M) if ta else (M, K)), generator=g)
    B = torch.randn(batch, *((N, K) if tb else (K, N)), generator=g)
    ref = torch.matmul((A.transpose(1, 2) if ta else A).double(), (B.transpose(1, 2) if tb else B).double())
    Ad, Bd = A.cuda().contiguous(), B.cuda().contiguous()
    Cd = torch.zeros(batch, M, N, device="cuda")
    L.check(lib.saceo_test_gemm(L.GEMM_FP32_SIMT, batch, M, N, K, ta, tb, Ad.data_ptr(), Bd.data_ptr(), Cd.data_ptr(),
                                torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    err = (Cd.cpu().double() - ref).norm() / ref.norm()
    assert err < 1e-6, float(err)


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 256, 256), (276, 256, 27), (256, 64, 35), (257, 128, 276),
                                   (100, 96, 16), (384, 512, 130)])
@pytest.mark.parametrize("variant", [0, 1])
def test_tcgen05_bf16x3_gemm(ta, tb, M, N, K, variant):
    """tcgen05 engine (bf16 hi/lo split, 3 MMAs, fp32 accumulate in TMEM) vs an fp64 matmul."""
    lib = L.load()
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N + K)
    batch = 3
    A = torch.randn(batch, *((K, M) if ta else (M, K)), generator=g)
    B = torch.randn(batch, *((N, K) if tb else (K, N)), generator=g)
    ref = torch.matmul((A.transpose(1, 2) if ta else A).double(), (B.transpose(1, 2) if tb else B).double())
    Ad, Bd = A.cuda().contiguous(), B.cuda().contiguous()
    Cd = torch.full((batch, M, N), float("nan"), device="cuda")
    L.check(lib.saceo_test_gemm(L.GEMM_TCGEN05_BF16X3 | (variant << 8), batch, M, N, K, ta, tb, Ad.data_ptr(), Bd.data_ptr(),
                                Cd.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    err = (Cd.cpu().double() - ref).norm() / ref.norm()
    assert err < 5e-5, float(err)     # ~2^-16 per product; single-pass bf16 would sit at ~4e-3


def test_replay_append_all_agents_wraparound():
    """saceo_replay_append: one call appends k rows to EVERY agent's ring on the device; after several
    wrap-arounds the gather is still a bit-exact copy of the last `capacity` rows in chronological order
    (buffers.py:41-71 keep-the-tail semantics), for all agents at once, with different fill levels per agent."""
    import numpy as np
    from sac_expert_b200.population import Population, PopulationSpec
    n, S, A, cap, B = 5, 7, 3, 96, 64
    pop = Population(PopulationSpec(n_agents=n, S=S, A=A, B=B, E=2, num_models=0, replay_capacity=cap, gemm_mode=0))
    rng = np.random.default_rng(3)
    hist = [dict(s=[], a=[], r=[], sp=[], d=[]) for _ in range(n)]
    # agents start at different fill levels (per-agent appends), then advance together
    for ag in range(n):
        k0 = 5 + 9 * ag
        s, a = rng.standard_normal((k0, S)).astype(np.float32), rng.standard_normal((k0, A)).astype(np.float32)
        r, sp, d = rng.standard_normal(k0).astype(np.float32), rng.standard_normal((k0, S)).astype(np.float32), (rng.random(k0) < 0.3).astype(np.float64)
        pop.append_rows(ag, s, a, r, sp, d)
        for key, v in zip("s a r sp d".split(), (s, a, r, sp, d)):
            hist[ag][key].append(v)
    for step in range(40):
        k = int(rng.integers(1, 12))
        s, a = rng.standard_normal((n, k, S)).astype(np.float32), rng.standard_normal((n, k, A)).astype(np.float32)
        r, sp = rng.standard_normal((n, k)).astype(np.float32), rng.standard_normal((n, k, S)).astype(np.float32)
        d = (rng.random((n, k)) < 0.3).astype(np.float64)
        s[0, 0, 0] = -0.0                                   # sign of zero survives
        pop.append_all(s, a, r, sp, d)
        for ag in range(n):
            for key, v in zip("s a r sp d".split(), (s, a, r, sp, d)):
                hist[ag][key].append(v[ag])
    torch.cuda.synchronize()
    sizes = pop.t["replay_size"].cpu().numpy()
    full = {key: [np.concatenate(h[key])[-cap:] for h in hist] for key in "s a r sp d".split()}
    assert np.array_equal(sizes, [len(x) for x in full["r"]]) and np.array_equal(sizes, pop._host_size)
    assert np.array_equal(pop.t["replay_start"].cpu().numpy(), pop._host_start)
    idx = np.stack([rng.integers(0, sizes[ag], size=B) for ag in range(n)]).astype(np.int64)
    idx[:, 0] = 0
    idx[:, 1] = sizes - 1                                   # oldest and newest logical rows
    gs, ga, gsp, gr, gd = [t.cpu().numpy() for t in pop.gather(torch.from_numpy(idx))]
    for ag in range(n):
        assert gs[ag].tobytes() == full["s"][ag][idx[ag]].tobytes()
        assert ga[ag].tobytes() == full["a"][ag][idx[ag]].tobytes()
        assert gsp[ag].tobytes() == full["sp"][ag][idx[ag]].tobytes()
        assert gr[ag].tobytes() == full["r"][ag][idx[ag]].tobytes()
        assert gd[ag].tobytes() == full["d"][ag][idx[ag]].tobytes()
    pop.close()
