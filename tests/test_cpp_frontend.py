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


# --- the reference's OWN test sources, unmodified --------------------------------------------------------------------
# tests/bench-ntt.cpp (driver: random input, NTT<kernel>::compute_forward / compute_inverse, every word against
# NTTReference, bench-ntt.cpp:60-64) with each tests/ntt-tests/*.hpp as NTT_TEST_CASE_FILE, and tests/test-modulus.cpp,
# compiled from /root/reference where they lie against the drop-in headers by `make -C oracle refbench` (outputs in
# oracle/_ref/refbench/, which travels to the GPU box like the other oracle/_ref files).  Only tests/cpp/shim stands in for
# Google Benchmark / GoogleTest and for the SVE intrinsics of the reference's tests/utility.hpp.
REFBENCH = os.path.join(ROOT, "oracle", "_ref", "refbench")
REFERENCE = "/root/reference"
REF_CASES = 15  # tests/ntt-tests/*.hpp: 10 scalar + 5 SVE compositions


def _refbench(suffix):
    if os.path.isdir(os.path.join(REFERENCE, "tests", "ntt-tests")):
        subprocess.run(["make", "-j", str(os.cpu_count() or 4), "refbench"], cwd=os.path.join(ROOT, "oracle"), check=True,
                       stdout=subprocess.DEVNULL)
    exes = sorted(os.path.join(REFBENCH, f) for f in (os.listdir(REFBENCH) if os.path.isdir(REFBENCH) else [])
                  if f.endswith(suffix))
    if not exes:
        pytest.skip("oracle/_ref/refbench not built (needs the reference checkout at build time)")
    return exes


def _run_reference_driver(exes):
    assert len(exes) == REF_CASES, exes
    for exe in exes:
        out = subprocess.run([exe, "--iterations=2"], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, (exe, out.stdout + out.stderr)
        lines = [ln for ln in out.stdout.splitlines() if ln.endswith(": ok")]
        assert len(lines) == 2 and lines[0].startswith("Forward, ") and lines[1].startswith("Inverse, "), (exe, out.stdout)


def test_unmodified_reference_driver_on_emulator():
    _run_reference_driver(_refbench(".emu"))


@pytest.mark.gpu
def test_unmodified_reference_driver_on_gpu():
    _run_reference_driver(_refbench(".gpu"))


def test_unmodified_reference_test_modulus():
    """tests/test-modulus.cpp (sums of all order-th roots vanish, 7 orders up to 2^28, Goldilocks) against the drop-in
    sventt::Modulus."""
    _refbench(".emu")
    exe = os.path.join(REFBENCH, "test-modulus")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "1 test(s), 0 failure(s)" in out.stdout, out.stdout[-2000:] + out.stderr
