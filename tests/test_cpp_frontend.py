# SPDX-License-Identifier: Apache-2.0
"""The reference's benchmark-as-test (tests/bench-ntt.cpp + tests/ntt-tests/*.hpp) compiled against
the drop-in C++20 headers (sve-ntt_b200/host/sventt): on the CPU with the host emulator, on the GPU
with libxntt.so."""
import os
import subprocess

import pytest

from conftest import ROOT

CPP = os.path.join(ROOT, "tests", "cpp")


def _build(target):
    subprocess.run(["make", f"_build/{target}"], cwd=CPP, check=True, stdout=subprocess.DEVNULL)
    return os.path.join(CPP, "_build", target)


def test_reference_compositions_on_emulator():
    exe = _build("ntt_tests_emu")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL OK" in out.stdout and "MISMATCH" not in out.stdout
    assert out.stdout.count(" ok") == 27


@pytest.mark.gpu
def test_reference_compositions_on_gpu():
    exe = _build("ntt_tests_gpu")
    out = subprocess.run([exe, "--big"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL OK" in out.stdout and "MISMATCH" not in out.stdout
    assert out.stdout.count(" ok") == 31


def test_kinnaes_class_on_emulator():
    """MagicSeriesKinnaes<m, PAdic64SVE<Modulus<N, g>>, n> (examples/magic-series-kinnaes) through the C++ drop-in."""
    exe = _build("kinnaes_tests_emu")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ALL OK" in out.stdout and out.stdout.count(" ok") == 2, out.stdout + out.stderr


@pytest.mark.gpu
def test_kinnaes_class_on_gpu():
    exe = _build("kinnaes_tests_gpu")
    out = subprocess.run([exe, "--all"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ALL OK" in out.stdout and out.stdout.count(" ok") == 12, out.stdout + out.stderr


def test_magic_series_example_on_emulator():
    """examples/magic-series (gaussian-polynomial.hpp drop-in): chunked power-series division through 2^15 NTT polynomial
    multiplies; the reference's known answers m = 10 .. 100 over its eight moduli (test-magic-series.cpp:22-39, 315-325)."""
    exe = _build("magic_series_tests_emu")
    out = subprocess.run([exe, "--all"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "ALL OK" in out.stdout and out.stdout.count(" ok") == 41, out.stdout + out.stderr


@pytest.mark.gpu
def test_magic_series_example_on_gpu():
    exe = _build("magic_series_tests_gpu")
    out = subprocess.run([exe, "--all"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "ALL OK" in out.stdout and out.stdout.count(" ok") == 41, out.stdout + out.stderr


@pytest.mark.gpu
def test_bench_harness_on_gpu():
    """tests/bench-ntt.cpp re-hosted for timing (tests/cpp/bench_ntt.cpp): 'Forward, <name>' / 'Inverse, <name>' lines with
    device-resident and host-buffer figures; here only that it runs and reports sane numbers."""
    exe = _build("bench_ntt")
    out = subprocess.run([exe, "--reps", "5", "--devices", "0,0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith(("Forward,", "Inverse,"))]
    assert len(lines) == 6 and sum("multi-GPU" in ln for ln in lines) == 2, out.stdout
