# SPDX-License-Identifier: Apache-2.0
"""Generates tests/golden/ntt_golden.json from the REFERENCE ITSELF: oracle/_ref/libnttref.so is
NTTReference (tests/ntt-reference.hpp) and sventt::Modulus (include/sventt/modulus.hpp) compiled
from /root/reference by oracle/Makefile.  Run in the build container (the reference checkout does
not exist on the GPU box):   python tests/golden/make_golden.py
Inputs are the xorshift64 stream of SURVEY.md section 8(c) (oracle_fill_xorshift)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib  # noqa: E402

P0, G0 = 0xFFFFFC6E80000001, 3
SEED = 0x9E3779B97F4A7C15
MODULI = [  # tests/test-ntt-reference.cpp:17-23 plus the production prime and the ntt-tests prime
    (P0, G0),
    (0x3A00000000000001, 3),
    (0x0C40000000000001, 5),
    (0x0C60000000000001, 7),
    (0x0003F00000000001, 11),
    (0x0002580000000001, 11),
    (0xFFFFFFFF00000001, 7),
]


def hx(v):
    return f"{int(v):016x}"


def summary(v, orc):
    x = 0
    for w in v:
        x ^= int(w)
    return {"first": hx(v[0]), "second": hx(v[1]) if v.size > 1 else None, "mid": hx(v[v.size // 2]),
            "last": hx(v[-1]), "xor": hx(x), "fnv": hx(orc.fnv64(v))}


def main():
    orc = oracle_lib.Oracle()  # only for the input stream and the fingerprint
    ref = oracle_lib.Reference()
    out = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libnttref.so (NTTReference)",
           "seed": hx(SEED), "full": [], "spot": [], "roots": []}
    for N, g in MODULI:
        for L in range(1, 8):
            a = orc.fill_xorshift(1 << L, SEED, N)
            f = ref.ntt_forward(a, N, g)
            i = ref.ntt_inverse(a, N, g)
            out["full"].append({"modulus": hx(N), "g": g, "log2_m": L, "input": [hx(v) for v in a],
                                "forward": [hx(v) for v in f], "inverse": [hx(v) for v in i]})
    for L in [3, 8, 10, 12, 13, 15, 17, 20]:
        a = orc.fill_xorshift(1 << L, SEED, P0)
        f = ref.ntt_forward(a, P0, G0)
        i = ref.ntt_inverse(a, P0, G0)
        out["spot"].append({"modulus": hx(P0), "g": G0, "log2_m": L, "forward": summary(f, orc),
                            "inverse": summary(i, orc)})
    for which, (N, g) in enumerate([(P0, G0), (0xFFFFFFFF00000001, 7)]):
        for order in [2, 4, 8, 1 << 12, 1 << 24, 1 << 28, 1 << 31, 3, 5, 17, 257, 65537, 7]:
            out["roots"].append({"modulus": hx(N), "g": g, "order": order,
                                 "forward": hx(ref.lib.ref_root(which, 0, order)),
                                 "inverse": hx(ref.lib.ref_root(which, 1, order))})
        out.setdefault("montgomery_inverse", {})[hx(N)] = hx(ref.lib.ref_montgomery_inverse(which))
    with open(os.path.join(HERE, "ntt_golden.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print("wrote", os.path.join(HERE, "ntt_golden.json"))


if __name__ == "__main__":
    main()
