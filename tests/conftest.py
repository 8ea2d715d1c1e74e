# SPDX-License-Identifier: Apache-2.0
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

P0 = 0xFFFFFC6E80000001
G0 = 3
SEED = 0x9E3779B97F4A7C15


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def reference():
    """NTTReference compiled from /root/reference (oracle/_ref); skipped where it was never built."""
    import oracle_lib
    if not oracle_lib.have_reference():
        if os.path.exists("/root/reference/tests/ntt-reference.hpp"):
            subprocess.run(["make", "ref"], cwd=os.path.join(ROOT, "oracle"), check=True)
        else:
            pytest.skip("oracle/_ref not built and no reference checkout")
    return oracle_lib.Reference()


@pytest.fixture(scope="session")
def emu(pkg):
    """The host emulator of the kernel templates (tests/emu) - CPU tests only."""
    d = os.path.join(ROOT, "tests", "emu")
    subprocess.run(["make"], cwd=d, check=True, stdout=subprocess.DEVNULL)
    return pkg.Library(os.path.join(d, "_build", "libxntt_emu.so"))


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "ntt_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def cuda_lib(pkg):
    """libxntt.so on a real device.  No fallback: a missing library or device is an error."""
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    lib = pkg.load()
    assert lib.device_count() >= 1
    torch.cuda.set_device(0)
    return lib
