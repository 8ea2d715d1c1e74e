# SPDX-License-Identifier: Apache-2.0
"""Kinnaes' formula for the number of magic series: the reference's second example,
examples/magic-series-kinnaes (kinnaes.hpp + test-magic-series-kinnaes.cpp).  Known answers: the reference's own
twelve (m, modulus, generator, n) cases (test-magic-series-kinnaes.cpp:18-67) against the counts it stores as
decimal strings (:75-92).  The oracle is pinned on them first, then the emulator (CPU) and libxntt (GPU) must
agree with both."""
import pytest

EXPECTED = {  # test-magic-series-kinnaes.cpp:75-92 (m = 10 .. 42 are the values of test-magic-series.cpp:315-325)
    10: 78132541528,
    25: 140170526450793924490478768121814869629364,
    100: int("904300736808894426574793302240693911261234942398748154528052171724"
             "305279045583459861011357813556260746366850646669062169890178280824"
             "885995375485156399921958991796250954308603011799192842071430359668"
             "946052264146938445899732873114858199920"),
    101: int("651742868521150599423217738842736563193389672725617304609189541060"
             "948075348430211017087941851686538398290713576362337481621156854784"
             "148283104866179994202618028615736621185423913319338987817995082551"
             "755913561634157004344784632798600635226832"),
}
CASES = [  # (m, modulus, generator, n): test-magic-series-kinnaes.cpp:18-67, 64- to 61-bit moduli
    (100, 0xFFFFFFFFFECA467F, 5, 495017), (100, 0xFFFFFFFFFE05E355, 6, 495017),
    (100, 0x7FFFFFFFFED59FB5, 2, 495017), (100, 0x7FFFFFFFFCC4E37F, 13, 495017),
    (100, 0x3FFFFFFFFF4C9937, 5, 495017), (100, 0x3FFFFFFFFF102BEF, 3, 495017),
    (100, 0x1FFFFFFFFFF962DF, 7, 495017), (100, 0x1FFFFFFFFFDB2C3B, 2, 495017),
    (101, 0xFFFFFFFFFE023EC1, 11, 510053), (101, 0x7FFFFFFFFD0F0621, 3, 510053),
    (101, 0x3FFFFFFFFEC5C639, 21, 510053), (101, 0x1FFFFFFFFDCE2E99, 3, 510053),
]


def small_case(m):
    """(modulus, generator, n) for a small order m, found the way generate-parameters.py does: n = the first
    prime above r = m^2 (m - 1) / 2, modulus = the largest 64-bit prime = 1 mod n, generator = a primitive root."""
    from sympy import isprime, nextprime, primitive_root
    n = int(nextprime(m * m * (m - 1) // 2))
    k = (2**64 - 2) // n
    while not isprime(k * n + 1):
        k -= 1
    N = k * n + 1
    return N, int(primitive_root(N)), n


def test_oracle_against_reference_known_answers(oracle):
    for m, N, g, n in CASES:
        assert oracle.kinnaes_compute(N, g, m, n) == EXPECTED[m] % N, (m, hex(N))


@pytest.mark.parametrize("m", [10, 25])
def test_small_orders_on_emulator(emu, oracle, m):
    N, g, n = small_case(m)
    assert oracle.kinnaes_compute(N, g, m, n) == EXPECTED[m] % N
    assert emu.kinnaes_compute(N, g, m, n) == EXPECTED[m] % N
    # compute_sum over sub-ranges (kinnaes.hpp:51): additive, equal to the oracle, empty range = 0
    h = n // 2
    a, b = h // 3, h // 2 + 1
    parts = [emu.kinnaes_sum(N, g, m, n, lo, hi) for lo, hi in ((0, a), (a, b), (b, h))]
    assert parts == [oracle.kinnaes_sum(N, g, m, n, lo, hi) for lo, hi in ((0, a), (a, b), (b, h))]
    assert sum(parts) % N == emu.kinnaes_sum(N, g, m, n, 0, h)
    assert emu.kinnaes_sum(N, g, m, n, a, a) == 0


def test_reference_case_on_emulator(emu):
    for m, N, g, n in (CASES[0], CASES[7], CASES[9]):  # 64-, 61- and 63-bit moduli, both orders
        assert emu.kinnaes_compute(N, g, m, n) == EXPECTED[m] % N, (m, hex(N))


def test_error_paths(emu, pkg):
    m, N, g, n = CASES[0]
    for args in [(N, g, m, n + 2), (N - 2, g, m, n), (N, g, 1, n), (N, 0, m, n)]:
        with pytest.raises(pkg.XnttError) as e:
            emu.kinnaes_compute(*args)
        assert e.value.status == pkg.ERR_INVALID
    with pytest.raises(pkg.XnttError):
        emu.kinnaes_sum(N, g, m, n, 5, 4)
    with pytest.raises(pkg.XnttError):
        emu.kinnaes_sum(N, g, m, n, 0, n // 2 + 1)


@pytest.mark.gpu
def test_reference_known_answers_on_gpu(cuda_lib, oracle):
    for m, N, g, n in CASES:
        assert cuda_lib.kinnaes_compute(N, g, m, n) == EXPECTED[m] % N, (m, hex(N))
    m, N, g, n = CASES[1]
    h = n // 2
    cuts = [0, 1, 257, h // 2, h - 1, h]
    parts = [cuda_lib.kinnaes_sum(N, g, m, n, lo, hi) for lo, hi in zip(cuts, cuts[1:])]
    assert sum(parts) % N == cuda_lib.kinnaes_sum(N, g, m, n, 0, h)
    assert parts[1] == oracle.kinnaes_sum(N, g, m, n, 1, 257)
