# SPDX-License-Identifier: Apache-2.0
"""Application-level known answers of the reference (BASELINE configs[4]): the number of magic series
of order m is the coefficient of q^(m^2 (m-1)/2) of the Gaussian polynomial [m^2 choose m]_q
(examples/magic-series/gaussian-polynomial.hpp:246-251), which the reference obtains with NTT
polynomial multiplication (forward, point-wise multiply_normalize against a to_montgomery'd spectrum,
inverse; gaussian-polynomial.hpp:196-214) and checks against the decimal strings of
examples/magic-series/test-magic-series.cpp:315-325 reduced mod p, over several moduli
(test-magic-series.cpp:22-39).  Here the same product runs through libxntt (fused forward+multiply)."""
import numpy as np
import pytest

from conftest import G0, P0

EXPECTED = {  # test-magic-series.cpp:315-325
    10: 78132541528,
    25: 140170526450793924490478768121814869629364,
    35: 13872534241478210358349096341203128450357241660871429860873721318,
    42: 1195452957914568544628242649935060977711193839443701120065551521757686130217168310,
}
MODULI = [(P0, G0), (0xFFFFFFFF00000001, 7), (0xA3B25F400C7A8001, 5), (0x3164C5D59B090001, 13)]


def gaussian_factors(m, N):
    """numerator prod_{i<=k} (1 - q^(n-k+i)) and 1 / prod_{i<=k} (1 - q^i), both truncated at degree d."""
    n, k, d = m * m, m, m * m * (m - 1) // 2
    num = [0] * (d + 1)
    num[0] = 1
    for i in range(1, k + 1):
        e = n - k + i
        for j in range(d, e - 1, -1):
            num[j] = (num[j] - num[j - e]) % N
    part = [0] * (d + 1)  # partitions into parts of size at most k (restricted partitions)
    part[0] = 1
    for i in range(1, k + 1):
        for j in range(i, d + 1):
            part[j] = (part[j] + part[j - i]) % N
    return num, part, d


def magic_series_via_ntt(lib, m, N, g, to_dev, to_host, stream=0):
    d = m * m * (m - 1) // 2
    L = max(2, (2 * (d + 1) - 1).bit_length())
    size = 1 << L
    if (N - 1) % size:
        return None  # this prime has no root of unity of that order (the reference chunks at 2^15 instead)
    num, part, d = gaussian_factors(m, N)
    a = np.zeros(size, np.uint64)
    b = np.zeros(size, np.uint64)
    a[:d + 1] = num
    b[:d + 1] = part
    plan = lib.plan(L, modulus=N, generator=g)
    da, db = to_dev(a), to_dev(b)
    pa, pb = (da.ctypes.data, db.ctypes.data) if isinstance(da, np.ndarray) else (da.data_ptr(), db.data_ptr())
    plan.forward(pb, pb, stream)
    plan.to_montgomery(pb, pb, size, stream)          # gaussian-polynomial.hpp:176-179
    plan.forward_multiply(pa, pa, pb, stream)         # :199-212, fused
    plan.inverse(pa, pa, stream)                      # :214
    prod = to_host(da)
    plan.close()
    # spot-check a few coefficients of the product against the direct convolution
    for idx in (0, 1, d // 3, d):
        want = sum(num[i] * part[idx - i] for i in range(idx + 1)) % N
        assert int(prod[idx]) == want, (m, hex(N), idx)
    return int(prod[d])


@pytest.mark.parametrize("N,g", MODULI)
def test_magic_series_small_orders_on_emulator(emu, N, g):
    for m in (10, 25):
        got = magic_series_via_ntt(emu, m, N, g, lambda x: x, lambda x: x)
        assert got is None or got == EXPECTED[m] % N, (m, hex(N))


@pytest.mark.gpu
@pytest.mark.parametrize("N,g", MODULI)
def test_magic_series_on_gpu(cuda_lib, N, g):
    import torch
    st = torch.cuda.current_stream().cuda_stream
    to_dev = lambda x: torch.from_numpy(x.view(np.int64)).cuda()  # noqa: E731
    to_host = lambda t: t.cpu().numpy().view(np.uint64)  # noqa: E731
    for m in (10, 25, 35, 42):
        got = magic_series_via_ntt(cuda_lib, m, N, g, to_dev, to_host, st)
        assert got is None or got == EXPECTED[m] % N, (m, hex(N))
