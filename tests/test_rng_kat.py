"""CPU: the in-kernel Philox4x32-10 (csrc/rng.cuh) against the published Random123 known-answer vectors.

``Philox::gen`` is ``__host__ __device__``: the header the kernels include is compiled into a host executable with
nvcc (no GPU needed) and fed the KAT counters / keys; an independent NumPy restatement is checked against the same
vectors and against the compiled code on random inputs.  The reference draws from NumPy MT19937 / PCG64 streams
(buffers.py:135, continuous_actors.py:297,350, SAC_expert.py:301-303) which cannot be reproduced on the device, so
this pins that the device stream is a correct Philox stream; its distributional use is tested on the GPU
(tests/test_gpu_engines.py::test_device_rng_statistics)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

# Random123 kat_vectors, philox4x32 10 rounds: (counter[4], key[2]) -> output[4]
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox4x32_10(ctr, key):
    """Independent restatement (Salmon et al., 'Parallel random numbers: as easy as 1, 2, 3', SC'11)."""
    c = [int(x) for x in ctr]
    k = [int(x) for x in key]
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xffffffff, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xffffffff]
        k = [(k[0] + 0x9E3779B9) & 0xffffffff, (k[1] + 0xBB67AE85) & 0xffffffff]
    return tuple(c)


def test_numpy_restatement_matches_the_known_answers():
    for ctr, key, out in KAT:
        assert philox4x32_10(ctr, key) == out


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if not (os.path.exists(NVCC) or shutil.which("nvcc")):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("philox") / "philox_host")
    subprocess.run([NVCC if os.path.exists(NVCC) else "nvcc", "-O1", "-Wno-deprecated-gpu-targets",
                    "-I", os.path.join(ROOT, "sac_expert_b200", "csrc"), "-o", exe,
                    os.path.join(ROOT, "tests", "philox_host.cu")], check=True, capture_output=True)
    return exe


def _run(exe, rows):
    """rows of (seed64, i, agent, step, stream) -> list of 4-tuples."""
    text = "".join("%x %x %x %x %x\n" % r for r in rows)
    out = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout.split()
    vals = [int(x, 16) for x in out]
    return [tuple(vals[i:i + 4]) for i in range(0, len(vals), 4)]


def test_device_generator_source_matches_the_known_answers(harness):
    # rng.cuh: counter = (i, agent, step, stream), key = (seed low, seed high)
    rows = [((key[1] << 32) | key[0],) + ctr for ctr, key, _ in KAT]
    assert _run(harness, rows) == [out for _, _, out in KAT]


def test_device_generator_source_matches_the_restatement_on_random_counters(harness):
    rng = np.random.default_rng(0)
    rows = [tuple(int(x) for x in (rng.integers(0, 1 << 63), *rng.integers(0, 1 << 32, size=4))) for _ in range(200)]
    got = _run(harness, rows)
    for r, g in zip(rows, got):
        assert g == philox4x32_10(r[1:], (r[0] & 0xffffffff, r[0] >> 32))
