"""CPU: the driver-facing contract of ``bench.py --impl reference`` (the reference arm runs without a GPU): one JSON
line with the keys the driver reads, the CPU baseline described with both ways the reference can use the host cores,
and a rank != 0 process that exits 0 without output."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, SACEO_REF_SECONDS="2", **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args],
                          capture_output=True, text=True, env=env, timeout=600, check=True)


def test_reference_arm_prints_one_contract_line():
    out = _run({}, "--shape", "hopper", "--steps", "2", "--warmup", "1").stdout.strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    assert line["impl"] == "reference" and line["metric"] == "agent-updates/sec" and line["unit"] == "agent-updates/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
    assert line["value"] > 0 and abs(line["ms_per_step"] - 1e3 / line["value"]) < 1e-6
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert cb["value"] == max(cb["single_process"], cb["process_pool"]) and cb["single_process"] > 0
    assert "train.py:130-152" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert r.stdout.strip() == ""
