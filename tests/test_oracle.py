# SPDX-License-Identifier: Apache-2.0
"""Pins the CPU oracle (oracle/ntt_oracle.c): against the reference's own oracle class compiled
from /root/reference, against the committed golden vectors generated from it, and against the
assertions of the reference's own tests (tests/test-ntt-reference.cpp:16-88,
tests/test-modulus.cpp:12-48)."""
import numpy as np
import pytest

from conftest import G0, P0, SEED

REF_PRIMES = [  # tests/test-ntt-reference.cpp:17-23
    (0x0C40000000000001, 5),
    (0x0C60000000000001, 7),
    (0x0003F00000000001, 11),
    (0x0002580000000001, 11),
    (0xFFFFFFFF00000001, 7),
]


def rand_residues(rng, n, N):
    return (rng.integers(0, 2**64, n, dtype=np.uint64) % np.uint64(N)).astype(np.uint64)


@pytest.mark.parametrize("N,g", REF_PRIMES + [(P0, G0), (0x3A00000000000001, 3)])
def test_reference_forward_inverse_assertions(oracle, N, g):
    """The three spot values and the round trip of NTTReference.ForwardInverse."""
    rng = np.random.default_rng(1)
    for log2m in range(1, 8):
        m = 1 << log2m
        a = rand_residues(rng, m, N)
        b = oracle.ntt_forward(a, N, g)
        ai = [int(v) for v in a]
        assert int(b[0]) == sum(ai) % N
        assert int(b[1]) == sum(v if i % 2 == 0 else -v for i, v in enumerate(ai)) % N
        w = pow(g, (N - 1) >> log2m, N)
        assert int(b[m // 2]) == sum(v * pow(w, i, N) for i, v in enumerate(ai)) % N
        assert np.array_equal(oracle.ntt_inverse(b, N, g), a)


def test_modulus_sum_of_roots(oracle):
    """tests/test-modulus.cpp: the powers of a primitive root of any order dividing p-1 sum to 0."""
    N, g = 0xFFFFFFFF00000001, 7
    for order in [3, 5, 17, 257, 65537, 1 << 14]:
        for root in (oracle.root_forward(N, g, order), oracle.root_inverse(N, g, order)):
            assert root > 1
            s, r = 0, 1
            for _ in range(order):
                s = oracle.addmod(s, r, N)
                r = oracle.mulmod(r, root, N)
            assert s == 0 and r == 1
    assert oracle.root_forward(N, g, 7) == 0  # 7 does not divide p-1: the reference throws


def test_oracle_equals_compiled_reference(oracle, reference):
    for N, g in [(P0, G0), (0x3A00000000000001, 3), (0xFFFFFFFF00000001, 7)]:
        for log2m in [1, 2, 3, 6, 9, 12, 14]:
            a = oracle.fill_xorshift(1 << log2m, SEED + log2m, N)
            f = reference.ntt_forward(a, N, g)
            assert np.array_equal(oracle.ntt_forward(a, N, g), f)
            assert np.array_equal(oracle.ntt_inverse(a, N, g), reference.ntt_inverse(a, N, g))
            assert np.array_equal(reference.ntt_inverse(f, N, g), a)


def test_modulus_constants_equal_compiled_reference(oracle, reference):
    for which, (N, g) in enumerate([(P0, G0), (0xFFFFFFFF00000001, 7)]):
        assert oracle.montgomery_inverse(N) == reference.lib.ref_montgomery_inverse(which)
        for order in [2, 8, 1 << 12, 1 << 24, 1 << 31, 3, 5, 7]:
            assert oracle.root_forward(N, g, order) == reference.lib.ref_root(which, 0, order)
            assert oracle.root_inverse(N, g, order) == reference.lib.ref_root(which, 1, order)
        rng = np.random.default_rng(2)
        for a, b in rng.integers(0, 2**63, (50, 2), dtype=np.uint64):
            assert oracle.mulmod(int(a), int(b), N) == reference.lib.ref_multiply(which, int(a), int(b))


def test_golden_full_vectors(oracle, golden):
    for case in golden["full"]:
        N, g = int(case["modulus"], 16), case["g"]
        a = np.array([int(v, 16) for v in case["input"]], dtype=np.uint64)
        assert np.array_equal(oracle.fill_xorshift(a.size, SEED, N), a)
        assert [f"{int(v):016x}" for v in oracle.ntt_forward(a, N, g)] == case["forward"]
        assert [f"{int(v):016x}" for v in oracle.ntt_inverse(a, N, g)] == case["inverse"]


def test_golden_spot_values(oracle, golden):
    for case in golden["spot"]:
        N, g, L = int(case["modulus"], 16), case["g"], case["log2_m"]
        a = oracle.fill_xorshift(1 << L, SEED, N)
        for name, out in (("forward", oracle.ntt_forward(a, N, g)), ("inverse", oracle.ntt_inverse(a, N, g))):
            want = case[name]
            assert f"{int(out[0]):016x}" == want["first"]
            assert f"{int(out[out.size // 2]):016x}" == want["mid"]
            assert f"{int(out[-1]):016x}" == want["last"]
            assert f"{oracle.fnv64(out):016x}" == want["fnv"]


def test_golden_roots(oracle, golden):
    for r in golden["roots"]:
        N = int(r["modulus"], 16)
        assert f"{oracle.root_forward(N, r['g'], r['order']):016x}" == r["forward"]
        assert f"{oracle.root_inverse(N, r['g'], r['order']):016x}" == r["inverse"]
    for mod, inv in golden["montgomery_inverse"].items():
        assert f"{oracle.montgomery_inverse(int(mod, 16)):016x}" == inv


def test_survey_spot_values(oracle):
    """values obtained at survey time from NTTReference."""
    a = oracle.fill_xorshift(1 << 17, SEED, P0)
    assert int(a[0]) == 0xDC1B77AE0BF34DAD and int(a[1]) == 0x64F0EEB9026E6076
    f = oracle.ntt_forward(a, P0, G0)
    assert (int(f[0]), int(f[1]), int(f[1 << 16]), int(f[-1])) == (
        0x10FABBFF4831888D, 0x42F944859A7AC339, 0xD7FA863F1FE232CC, 0x837F665F2EA89A8E)
    assert oracle.fnv64(f) == 0xB88190C5AB40DB29


def test_padic64_scalar_arithmetic(oracle):
    """PAdic64 identities (modmul/sve/p-adic-64.hpp:19-38,64-115) for 64-, 63- and 62-bit moduli."""
    rng = np.random.default_rng(3)
    for N in [P0, 0xA3B25F400C7A8001, 0x41D33D0D1FBF8001, 0x3A00000000000001]:
        r = (1 << 64) % N
        for a, b in rng.integers(0, 2**64, (200, 2), dtype=np.uint64):
            a, b = int(a) % N, int(b) % N
            bm = oracle.to_montgomery(b, N)
            assert bm == b * r % N
            assert oracle.from_montgomery(bm, N) == b
            bp = oracle.precompute(bm, N)
            assert (bp * N) % (1 << 64) == bm
            assert oracle.multiply_normalize(a, bm, bp, N) == a * b % N


def test_dft_point_matches_forward(oracle):
    a = oracle.fill_xorshift(1 << 10, SEED, P0)
    f = oracle.ntt_forward(a, P0, G0)
    for pos in [0, 1, 2, 511, 512, 1023]:
        assert oracle.dft_point(a, P0, G0, pos) == int(f[pos])


def test_reference_scalar_kernel_matches_oracle(oracle):
    """CPU baseline B2/B3: the reference's scalar IterativeNTT<RadixEightScalarLayer...>
    compiled from /root/reference (oracle/refscalar.cpp) computes what NTTReference computes at the 62-bit test
    prime, the check of tests/bench-ntt.cpp:60-64 (`dst[i] % N == dst_ref[i]`); threads only split the batch."""
    import oracle_lib
    if not oracle_lib.have_reference_scalar():
        pytest.skip("oracle/_ref/libnttref_scalar.so not built (needs /root/reference)")
    ref = oracle_lib.ReferenceScalar()
    N, g = ref.N, ref.g
    assert (N, g) == (0x3A00000000000001, 3)
    for L, batch, threads in [(12, 1, 1), (12, 5, 3), (20, 1, 1), (20, 3, 2)]:
        m = 1 << L
        a = oracle.fill_xorshift(m * batch, SEED + L, N)
        f = ref.run(L, False, a, batch, threads)
        back = ref.run(L, True, f, batch, threads)
        for b in range(batch):
            want = oracle.ntt_forward(a[b * m:(b + 1) * m].copy(), N, g)
            assert np.array_equal(f[b * m:(b + 1) * m] % np.uint64(N), want), (L, b)
        assert np.array_equal(back % np.uint64(N), a), L
    with pytest.raises(ValueError):
        ref.run(13, False, np.zeros(1 << 13, np.uint64))
