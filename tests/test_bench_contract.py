# SPDX-License-Identifier: Apache-2.0
"""bench.py's reference arm runs without a GPU: its JSON line carries the keys the driver reads, rank > 0 of a
multi-rank launch stays silent, at --gpus N > 1 the workload is the sharded 2^30 transform (a bounded sample of it), and
the reference's scalar kernel (CPU baseline B2) is what gets timed when it has been built."""
import json
import os
import subprocess
import sys

from conftest import ROOT

KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def run(args, **env):
    e = dict(os.environ, **dict({"XNTT_BENCH_REF_LOG2": "14"}, **env))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args],
                          capture_output=True, text=True, timeout=300, env=e, cwd=ROOT)


def test_reference_arm_line():
    out = run(["--steps", "2", "--warmup", "1"])
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert KEYS <= set(line) and line["impl"] == "reference" and line["unit"] == "Gelem/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0


def test_reference_arm_multi_rank():
    quiet = run(["--gpus", "2", "--steps", "1", "--warmup", "0"], RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""
    out = run(["--gpus", "2", "--steps", "1", "--warmup", "0"], RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong" and "2^30" in line["config"]["workload"]
    assert line["config"]["same_size_as_workload"] is False  # a bounded sample, and the line says so


def test_reference_arm_times_the_scalar_kernel_when_built():
    import oracle_lib
    if not oracle_lib.have_reference_scalar():
        import pytest
        pytest.skip("oracle/_ref/libnttref_scalar.so not built")
    out = run(["--steps", "2", "--warmup", "1"], XNTT_BENCH_REF_LOG2="12")
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert "RadixEightScalarLayer" in line["cpu_baseline"]["sample"] and line["cpu_baseline"]["kind"] == "reference"
    assert line["config"]["log2_n"] == 12 and line["value"] > 0
    # the batched workload runs B3: OpenMP over the batch on all host cores
    out = run(["--steps", "1", "--warmup", "0", "--workload", "batch20"], XNTT_BENCH_REF_LOG2="12")
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and line["config"]["batch"] >= 8
