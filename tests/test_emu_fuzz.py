# SPDX-License-Identifier: Apache-2.0
"""Random plan shapes on the host emulator against the oracle: length, explicit splits, batch, modulus, twiddle-table
form, tile shape and inverse_factor drawn from a fixed seed (a few hundred plans; `python tests/test_emu_fuzz.py SEED SECONDS`
keeps drawing).  Forward, scaled inverse and the fused point-wise product are compared word for word."""
import random
import sys
import time

import numpy as np
import pytest

MODS = [(0xFFFFFC6E80000001, 3), (0xFFFFFFFF00000001, 7), (0x3A00000000000001, 3), (0xA3B25F400C7A8001, 5)]


def draw(rng, max_l=16):
    L = rng.randint(2, max_l)
    parts = rng.choice([1, 2, 2, 3])
    splits = None
    if parts > 1:
        if L <= parts:
            return None
        cuts = sorted(rng.sample(range(1, L), parts - 1))
        splits = [b - a for a, b in zip([0] + cuts, cuts + [L])]
    N, g = rng.choice(MODS)
    if (N - 1) % (1 << L):
        return None
    return dict(L=L, splits=splits, batch=rng.choice([1, 1, 2, 3, 5]), N=N, g=g, compact=rng.random() < 0.4,
                invf=rng.choice([None, 1, 12345]), seed=rng.getrandbits(60), tiles=rng.choice([None, None, "wide", "narrow"]))


def check(emu, pkg, orc, c):
    L, N, g, batch = c["L"], c["N"], c["g"], c["batch"]
    m = 1 << L
    try:
        plan = emu.plan(L, modulus=N, generator=g, splits=c["splits"], batch=batch, compact_tables=c["compact"],
                        inverse_factor=c["invf"], tiles=c["tiles"])
    except pkg.XnttError as e:
        assert e.status in (pkg.ERR_INVALID, pkg.ERR_UNSUPPORTED), c  # shapes without a tile layout are refused
        return False
    a = orc.fill_xorshift(m * batch, c["seed"], N)
    out = np.empty_like(a)
    plan.forward(out.ctypes.data, a.ctypes.data)
    for b in range(batch):
        assert np.array_equal(out[b * m:(b + 1) * m], orc.ntt_forward(a[b * m:(b + 1) * m].copy(), N, g)), c
    back = np.empty_like(a)
    plan.inverse(back.ctypes.data, out.ctypes.data)
    f = m if c["invf"] is None else c["invf"]
    assert np.array_equal(back, orc.pointwise_mul(a, np.full_like(a, (m * pow(f, -1, N)) % N), N)), c
    bm = orc.fill_xorshift(m * batch, 7, N)
    fm = np.empty_like(a)
    plan.forward_multiply(fm.ctypes.data, a.ctypes.data, bm.ctypes.data)
    want = orc.pointwise_mul(orc.pointwise_mul(out, bm, N), np.full_like(a, pow(1 << 64, -1, N)), N)
    assert np.array_equal(fm, want), c
    plan.close()
    return True


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_plans(emu, pkg, oracle, seed):
    rng = random.Random(seed)
    done = 0
    while done < 120:
        c = draw(rng)
        if c is not None and check(emu, pkg, oracle, c):
            done += 1


if __name__ == "__main__":
    import os
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    import oracle_lib
    pkg_ = ge.load_package()
    emu_ = pkg_.Library(os.path.join(ROOT, "tests", "emu", "_build", "libxntt_emu.so"))
    orc_ = oracle_lib.Oracle()
    rng_ = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
    t_end = time.time() + (float(sys.argv[2]) if len(sys.argv) > 2 else 60)
    n = 0
    while time.time() < t_end:
        c_ = draw(rng_)
        if c_ is not None and check(emu_, pkg_, orc_, c_):
            n += 1
    print("ok", n)
