# SPDX-License-Identifier: Apache-2.0
"""Parity tests proper: the CUDA path, called through the C ABI (include/xntt.h) with device
pointers, against the CPU oracle on the same seeded inputs - bit exact.  Shapes follow the
reference's ntt-tests (tests/ntt-tests/*.hpp), its README example and BASELINE.json's configs; at
full size, size-independent properties (round trip, linearity, directly evaluated output words)
complement the word-for-word comparison."""
import numpy as np
import pytest

from conftest import G0, P0, SEED

pytestmark = pytest.mark.gpu


def dev(a):
    import torch
    return torch.from_numpy(a.view(np.int64)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint64)


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def check_against_oracle(lib, oracle, L, splits=None, batch=1, inverse_factor=None, check_batches=None):
    import torch
    m = 1 << L
    a = oracle.fill_xorshift(m * batch, SEED + L, P0)
    plan = lib.plan(L, splits=splits, batch=batch, inverse_factor=inverse_factor)
    src = dev(a)
    dst = torch.full_like(src, 0x5555555555555555)  # bench-ntt.cpp:34 poisons dst
    plan.forward(dst.data_ptr(), src.data_ptr(), stream())
    got = host(dst)
    assert np.array_equal(host(src), a), "out-of-place transform modified its source"
    for b in (check_batches if check_batches is not None else range(batch)):
        want = oracle.ntt_forward(a[b * m:(b + 1) * m].copy(), P0, G0)
        assert np.array_equal(got[b * m:(b + 1) * m], want), (L, plan.splits, b)
    assert (got < np.uint64(P0)).all(), "non-canonical output word"
    back = torch.empty_like(src)
    plan.inverse(back.data_ptr(), dst.data_ptr(), stream())
    f = m if inverse_factor is None else inverse_factor
    scale = np.full_like(a, (m * pow(f, -1, P0)) % P0)
    assert np.array_equal(host(back), oracle.pointwise_mul(a, scale, P0)), (L, plan.splits)
    # in place
    buf = src.clone()
    plan.forward(buf.data_ptr(), buf.data_ptr(), stream())
    assert np.array_equal(host(buf), got)
    plan.close()
    return got


@pytest.mark.parametrize("L", range(1, 21))
def test_forward_inverse_all_sizes(cuda_lib, oracle, L):
    check_against_oracle(cuda_lib, oracle, L)


@pytest.mark.parametrize("L,splits", [
    (17, [8, 9]),   # README.md:28-68 (BASELINE configs[0])
    (13, [9, 4]), (15, [9, 6]),  # four-step shapes of tests/ntt-tests
    (13, [1, 12]), (14, [12, 2]), (10, [5, 5]), (16, [5, 5, 6]), (18, [6, 6, 6]), (20, [10, 10]), (21, [7, 7, 7]),
])
def test_explicit_decompositions(cuda_lib, oracle, L, splits):
    check_against_oracle(cuda_lib, oracle, L, splits=splits)


@pytest.mark.parametrize("L,batch", [(3, 1000), (6, 37), (10, 7), (12, 5), (13, 3), (15, 3), (17, 2)])
def test_batches_and_ragged_tiles(cuda_lib, oracle, L, batch):
    check_against_oracle(cuda_lib, oracle, L, batch=batch)


@pytest.mark.parametrize("L,splits", [(9, None), (12, None), (14, None), (16, [5, 5, 6])])
def test_inverse_factor(cuda_lib, oracle, L, splits):
    check_against_oracle(cuda_lib, oracle, L, splits=splits, inverse_factor=1)
    check_against_oracle(cuda_lib, oracle, L, splits=splits, inverse_factor=0xDEADBEEFCAFE)


def test_golden_fingerprints(cuda_lib, oracle, golden):
    """Fingerprints generated from the reference's NTTReference (tests/golden/make_golden.py)."""
    import torch
    for case in golden["spot"]:
        L = case["log2_m"]
        a = oracle.fill_xorshift(1 << L, SEED, P0)
        plan = cuda_lib.plan(L)
        src = dev(a)
        out = torch.empty_like(src)
        for name, fn in (("forward", plan.forward), ("inverse", plan.inverse)):
            fn(out.data_ptr(), src.data_ptr(), stream())
            v = host(out)
            want = case[name]
            assert f"{int(v[0]):016x}" == want["first"]
            assert f"{int(v[v.size // 2]):016x}" == want["mid"]
            assert f"{int(v[-1]):016x}" == want["last"]
            assert f"{oracle.fnv64(v):016x}" == want["fnv"], (L, name)
        plan.close()


def test_golden_full_vectors(cuda_lib, golden):
    """All seven moduli of the golden file (test-ntt-reference.cpp's five, the ntt-tests prime, p0)."""
    import torch
    for case in golden["full"]:
        N, g = int(case["modulus"], 16), case["g"]
        a = np.array([int(v, 16) for v in case["input"]], dtype=np.uint64)
        plan = cuda_lib.plan(case["log2_m"], modulus=N, generator=g)
        src = dev(a)
        out = torch.empty_like(src)
        plan.forward(out.data_ptr(), src.data_ptr(), stream())
        assert [f"{int(v):016x}" for v in host(out)] == case["forward"]
        plan.inverse(out.data_ptr(), src.data_ptr(), stream())
        assert [f"{int(v):016x}" for v in host(out)] == case["inverse"]
        plan.close()


def test_edge_inputs(cuda_lib, oracle):
    import torch
    for L in (12, 16):
        m = 1 << L
        plan = cuda_lib.plan(L)
        delta = np.zeros(m, np.uint64)
        delta[1] = P0 - 1
        for a in (np.zeros(m, np.uint64), np.full(m, P0 - 1, np.uint64), delta, np.full(m, 1, np.uint64)):
            src = dev(a)
            out = torch.empty_like(src)
            plan.forward(out.data_ptr(), src.data_ptr(), stream())
            assert np.array_equal(host(out), oracle.ntt_forward(a, P0, G0))
            plan.inverse(out.data_ptr(), out.data_ptr(), stream())
            assert np.array_equal(host(out), a)
        plan.close()


@pytest.mark.parametrize("L,splits", [(14, None), (16, [5, 5, 6]), (20, None), (18, [6, 6, 6])])
def test_lazy_residues_between_passes(cuda_lib, oracle, L, splits):
    """Column passes whose consumer begins with a Montgomery product store residues uncanonicalised (forward: every
    twist-free / handover column pass; inverse: the inner column pass of a three-pass plan).  Inputs built from p - 1,
    p - 2, 0 and 1 push the intermediate sums to both ends of [0, 2^64); a complete transform still returns canonical
    words equal to the reference's (tests/ntt-reference.hpp:43-83).  Twin of tests/test_emu_plan.py."""
    import torch
    m = 1 << L
    rnd = oracle.fill_xorshift(m, SEED + 99, P0) | np.uint64(0xFFFFFC0000000000)
    pats = [np.full(m, P0 - 1, dtype=np.uint64),
            np.where(np.arange(m) % 2 == 0, np.uint64(P0 - 1), np.uint64(0)).astype(np.uint64),
            np.where(np.arange(m) % 3 == 0, np.uint64(P0 - 2), np.uint64(1)).astype(np.uint64),
            np.where(rnd >= np.uint64(P0), np.uint64(P0 - 1), rnd).astype(np.uint64)]
    for compact in (False, True):
        plan = cuda_lib.plan(L, splits=splits, compact_tables=compact)
        for a in pats:
            src = dev(a)
            out = torch.empty_like(src)
            plan.forward(out.data_ptr(), src.data_ptr(), stream())
            got = host(out)
            assert (got < np.uint64(P0)).all(), "non-canonical output word"
            assert np.array_equal(got, oracle.ntt_forward(a.copy(), P0, G0)), (L, splits, compact)
            plan.inverse(out.data_ptr(), out.data_ptr(), stream())
            assert np.array_equal(host(out), a), (L, splits, compact)
        plan.close()


def test_padic64_elementwise(cuda_lib, oracle):
    """L0 of SURVEY.md section 4: the device modmul against 128-bit integer arithmetic on random and
    edge operands (0, 1, p-1, 2^32 boundaries, lazy values >= p)."""
    import torch
    rng = np.random.default_rng(5)
    edge = np.array([0, 1, 2, P0 - 1, P0 - 2, 2**63, 2**32, 2**32 - 1, 0x3917FFFFFFF, 0x80000001, 0xFFFFFC6E],
                    dtype=np.uint64)
    a = np.concatenate([np.repeat(edge, edge.size), rng.integers(0, P0, 1 << 16, dtype=np.uint64)])
    b = np.concatenate([np.tile(edge, edge.size), rng.integers(0, P0, 1 << 16, dtype=np.uint64)])
    plan = cuda_lib.plan(4)
    da, db = dev(a), dev(b)
    bm, back, prod = torch.empty_like(db), torch.empty_like(db), torch.empty_like(da)
    plan.to_montgomery(bm.data_ptr(), db.data_ptr(), b.size, stream())
    r = (1 << 64) % P0
    assert [int(v) for v in host(bm)[:300]] == [int(v) * r % P0 for v in b[:300]]
    plan.from_montgomery(back.data_ptr(), bm.data_ptr(), b.size, stream())
    assert np.array_equal(host(back), b)
    plan.multiply_normalize(prod.data_ptr(), da.data_ptr(), bm.data_ptr(), a.size, stream())
    assert np.array_equal(host(prod), oracle.pointwise_mul(a, b, P0))
    # lazy (non-canonical) left operands: a + p still stands for a
    lazy = a[a < np.uint64(2**64 - P0)] + np.uint64(P0)
    small = a[a < np.uint64(2**64 - P0)]
    dl = dev(lazy)
    out = torch.empty_like(dl)
    plan.multiply_normalize(out.data_ptr(), dl.data_ptr(), bm[:lazy.size].data_ptr(), lazy.size, stream())
    assert np.array_equal(host(out), oracle.pointwise_mul(small, b[:lazy.size].copy(), P0))
    plan.close()


def test_polynomial_multiply(cuda_lib, oracle):
    """BASELINE configs[4]: forward, point-wise multiply_normalize against a to_montgomery'd
    spectrum, inverse (examples/magic-series/gaussian-polynomial.hpp:176-214)."""
    import torch
    L = 15  # the reference's GaussianPolynomialCoefficient test length (test-magic-series.cpp:304)
    m = 1 << L
    rng = np.random.default_rng(7)
    a = np.zeros(m, np.uint64)
    b = np.zeros(m, np.uint64)
    a[:m // 2] = rng.integers(0, P0, m // 2, dtype=np.uint64)
    b[:m // 2] = rng.integers(0, P0, m // 2, dtype=np.uint64)
    plan = cuda_lib.plan(L)
    da, db = dev(a), dev(b)
    plan.forward(da.data_ptr(), da.data_ptr(), stream())
    plan.forward(db.data_ptr(), db.data_ptr(), stream())
    plan.to_montgomery(db.data_ptr(), db.data_ptr(), m, stream())
    plan.multiply_normalize(da.data_ptr(), da.data_ptr(), db.data_ptr(), m, stream())
    plan.inverse(da.data_ptr(), da.data_ptr(), stream())
    want = oracle.ntt_inverse(oracle.pointwise_mul(oracle.ntt_forward(a, P0, G0), oracle.ntt_forward(b, P0, G0), P0),
                              P0, G0)
    assert np.array_equal(host(da), want)
    # the fused form (point-wise product inside the last forward pass) gives the same words
    da2 = dev(a)
    plan.forward_multiply(da2.data_ptr(), da2.data_ptr(), db.data_ptr(), stream())
    plan.inverse(da2.data_ptr(), da2.data_ptr(), stream())
    assert np.array_equal(host(da2), want)
    # a few coefficients by schoolbook convolution
    ai, bi = [int(v) for v in a[:64]], [int(v) for v in b[:64]]
    for k in (0, 1, 17, 63):
        assert int(want[k]) == sum(ai[i] * bi[k - i] for i in range(k + 1)) % P0
    plan.close()


def test_host_entry_points(cuda_lib, oracle):
    a = oracle.fill_xorshift(1 << 16, SEED, P0)
    plan = cuda_lib.plan(16)
    out, back = np.empty_like(a), np.empty_like(a)
    plan.forward_host(out.ctypes.data, a.ctypes.data)
    assert np.array_equal(out, oracle.ntt_forward(a, P0, G0))
    plan.inverse_host(back.ctypes.data, out.ctypes.data)
    assert np.array_equal(back, a)
    plan.close()


@pytest.mark.parametrize("L,batch", [(16, 1), (12, 5), (20, 2), (9, 3), (18, 37), (22, 1), (10, 1), (13, 1), (11, 4), (5, 7)])
def test_host_entry_points_page_locked(cuda_lib, oracle, L, batch):
    """Page-locked host buffers (what bench.py's e2e leg and sventt::PageMemory hand in): same words as the
    pageable route, also in place, batched (large batches flow through the copy / transform / copy chunk pipeline,
    here 37 x 2^18 = 8 ragged chunks; one 2^22 transform has its row pass cut into chunks that overlap the copies) and
    for single-pass plans.  Single-pass plans of up to 64 KiB work on the page-locked buffers themselves (2^10, 4 x 2^11,
    7 x 2^5: straight from host to host, no staging)."""
    import torch
    m = 1 << L
    a = oracle.fill_xorshift(m * batch, SEED + L, P0)
    want = np.concatenate([oracle.ntt_forward(a[b * m:(b + 1) * m].copy(), P0, G0) for b in range(batch)])
    plan = cuda_lib.plan(L, batch=batch)
    src = torch.from_numpy(a.view(np.int64)).pin_memory()
    dst = torch.empty_like(src).pin_memory()
    plan.forward_host(dst.data_ptr(), src.data_ptr())
    assert np.array_equal(dst.numpy().view(np.uint64), want)
    assert np.array_equal(src.numpy().view(np.uint64), a)  # out of place keeps the source
    back = torch.empty_like(src).pin_memory()
    plan.inverse_host(back.data_ptr(), dst.data_ptr())
    assert np.array_equal(back.numpy().view(np.uint64), a)
    # in place, page-locked
    plan.forward_host(src.data_ptr(), src.data_ptr())
    assert np.array_equal(src.numpy().view(np.uint64), want)
    plan.inverse_host(src.data_ptr(), src.data_ptr())
    assert np.array_equal(src.numpy().view(np.uint64), a)
    # mixed: page-locked source, pageable destination and the other way round
    out = np.empty_like(a)
    plan.forward_host(out.ctypes.data, src.data_ptr())
    assert np.array_equal(out, want)
    plan.inverse_host(back.data_ptr(), out.ctypes.data)
    assert np.array_equal(back.numpy().view(np.uint64), a)
    plan.close()


def test_full_size_2p24(cuda_lib, oracle):
    """BASELINE configs[1]: n = 2^24 on one GPU, word for word against the oracle, plus round trip,
    linearity and directly evaluated output words."""
    import torch
    L = 24
    m = 1 << L
    a = oracle.fill_xorshift(m, SEED, P0)
    b = oracle.fill_xorshift(m, SEED ^ 0xABCDEF, P0)
    plan = cuda_lib.plan(L)
    da, db = dev(a), dev(b)
    fa, fb = torch.empty_like(da), torch.empty_like(db)
    plan.forward(fa.data_ptr(), da.data_ptr(), stream())
    plan.forward(fb.data_ptr(), db.data_ptr(), stream())
    ha = host(fa)
    assert (int(ha[0]), int(ha[1]), int(ha[m // 2])) == (0x94D8DBE8AB43AB69, 0x21E752B9803B0FD6, 0x0D2A95309AC7D97B)
    assert np.array_equal(ha, oracle.ntt_forward(a, P0, G0))
    for pos in (0, 1, 12345, m // 2 + 7, m - 1):
        assert int(ha[pos]) == oracle.dft_point(a, P0, G0, pos)
    # linearity: F(a + b) == F(a) + F(b)
    s = (a.astype(object) + b.astype(object)) % P0
    s = np.array(s, dtype=np.uint64)
    fs = torch.empty_like(da)
    plan.forward(fs.data_ptr(), dev(s).data_ptr(), stream())
    hb = host(fb)
    sum_f = np.array((ha.astype(object) + hb.astype(object)) % P0, dtype=np.uint64)
    assert np.array_equal(host(fs), sum_f)
    back = torch.empty_like(da)
    plan.inverse(back.data_ptr(), fa.data_ptr(), stream())
    assert np.array_equal(host(back), a)
    plan.close()


def test_batched_256x2p20(cuda_lib, oracle):
    """BASELINE configs[2] on one GPU: 256 transforms of 2^20 in one call."""
    import torch
    L, batch = 20, 256
    m = 1 << L
    plan = cuda_lib.plan(L, batch=batch)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(9)
    src = torch.randint(0, 2**62, (m * batch,), dtype=torch.int64, device="cuda", generator=gen)
    dst = torch.empty_like(src)
    plan.forward(dst.data_ptr(), src.data_ptr(), stream())
    for b in (0, 131, 255):
        a = host(src[b * m:(b + 1) * m])
        assert np.array_equal(host(dst[b * m:(b + 1) * m]), oracle.ntt_forward(a.copy(), P0, G0))
    plan.inverse(dst.data_ptr(), dst.data_ptr(), stream())
    assert torch.equal(dst, src)
    plan.close()


def test_three_pass_full_compare_2p26(cuda_lib, oracle):
    """Three-pass plan (the shape every 2^30 shard takes), every output word against the oracle at 2^26
    (the oracle needs ~25 s for it), forward and inverse - the contract of tests/bench-ntt.cpp:60-64."""
    import torch
    L = 26
    m = 1 << L
    plan = cuda_lib.plan(L)
    assert len(plan.splits) == 3
    a = oracle.fill_xorshift(m, SEED + L, P0)
    src = dev(a)
    dst = torch.empty_like(src)
    plan.forward(dst.data_ptr(), src.data_ptr(), stream())
    want = oracle.ntt_forward(a, P0, G0)
    got = host(dst)
    assert np.array_equal(got, want)
    plan.inverse(dst.data_ptr(), dst.data_ptr(), stream())
    assert torch.equal(dst, src)
    plan.close()
    # 2^6 x 2^8 x 2^12: here the outer pass hands its forward twiddle matrix to the 2^8 pass behind it (kColPre)
    plan = cuda_lib.plan(L, splits=[6, 8, 12])
    assert plan.twiddle_forms(False) == [3, 3, 0] and plan.twiddle_forms(True) == [2, 3, 0]
    plan.forward(dst.data_ptr(), src.data_ptr(), stream())
    assert np.array_equal(host(dst), want)
    plan.inverse(dst.data_ptr(), dst.data_ptr(), stream())
    assert torch.equal(dst, src)
    plan.close()


@pytest.mark.parametrize("L", [28, 30])
def test_three_pass_sizes(cuda_lib, oracle, L):
    """BASELINE configs[3] sizes on one GPU: directly evaluated output words (oracle.dft_point on a sparse input - a
    dense 2^30-term sum per word is minutes of CPU), linearity against a dense input, and the round trip."""
    import torch
    m = 1 << L
    plan = cuda_lib.plan(L)
    assert len(plan.splits) == 3
    # sparse input: 64 non-zero words at scattered positions -> every output word is a 64-term sum the oracle
    # evaluates exactly; checks 4096 output positions spread over all tiles of all three passes
    rng = np.random.default_rng(L)
    pos = np.unique(rng.integers(0, m, 64, dtype=np.int64))
    val = rng.integers(1, P0, pos.size, dtype=np.uint64)
    sparse = torch.zeros(m, dtype=torch.int64, device="cuda")
    sparse[torch.from_numpy(pos).cuda()] = torch.from_numpy(val.view(np.int64)).cuda()
    out = torch.empty_like(sparse)
    plan.forward(out.data_ptr(), sparse.data_ptr(), stream())
    omega = oracle.root_forward(P0, G0, m)
    idx = np.unique(np.concatenate([rng.integers(0, m, 4000, dtype=np.int64), [0, 1, m // 2, m - 1, m - 2]]))
    got = host(out[torch.from_numpy(idx).cuda()])
    for i, g in zip(idx, got):
        k = int(f"{int(i):0{L}b}"[::-1], 2)  # output index i holds frequency bitrev(i)
        want = 0
        for p_, v_ in zip(pos, val):
            want = (want + int(v_) * pow(omega, (k * int(p_)) % m, P0)) % P0
        assert int(g) == want, (L, int(i))
    # linearity: F(dense + sparse) == F(dense) + F(sparse), all words
    gen = torch.Generator(device="cuda")
    gen.manual_seed(L)
    dense = torch.randint(0, 2**62, (m,), dtype=torch.int64, device="cuda", generator=gen)
    fd = torch.empty_like(dense)
    plan.forward(fd.data_ptr(), dense.data_ptr(), stream())
    # (dense + sparse) mod p: only the 64 sparse positions change
    summed = dense.clone()
    sel = torch.from_numpy(pos).cuda()
    sv = host(sparse[sel]).astype(object) + host(dense[sel]).astype(object)
    summed[sel] = torch.from_numpy(np.array([int(x) % P0 for x in sv], dtype=np.uint64).view(np.int64)).cuda()
    del sparse
    fs = torch.empty_like(dense)
    plan.forward(fs.data_ptr(), summed.data_ptr(), stream())
    del summed
    # fs - fd must equal out (mod p): a random sample of positions plus the directly evaluated ones
    samp = torch.from_numpy(np.unique(np.concatenate([rng.integers(0, m, 20000, dtype=np.int64), idx]))).cuda()
    a_, b_, c_ = (host(x[samp]).astype(object) for x in (fs, fd, out))
    assert all((int(x) - int(y)) % P0 == int(z) for x, y, z in zip(a_, b_, c_))
    del fs, out
    plan.inverse(fd.data_ptr(), fd.data_ptr(), stream())
    assert torch.equal(fd, dense)
    plan.close()


@pytest.mark.parametrize("L,G,splits", [(20, 2, None), (20, 4, [10, 10]), (20, 8, None), (24, 2, None), (24, 4, None),
                                        (24, 8, None), (21, 4, [7, 7, 7]), (15, 8, [3, 12]), (26, 8, None)])
def test_sharded_plan_on_one_gpu(cuda_lib, oracle, L, G, splits):
    """The N > 1 path on a single-GPU box: the C++-hosted multi-GPU plan (xntt_mgpu_*) with every rank on device 0.
    The peer-store kernels only need addresses, so 'peer memory' is simply the other ranks' buffers; every rank's
    slice is compared with the oracle word for word, forward and inverse, twice (both exchange buffers)."""
    import torch
    m = 1 << L
    mg = cuda_lib.mgpu(L, [0] * G, splits=splits)
    n0, n1 = mg.n0, mg.n1
    for rep in range(2):
        a = oracle.fill_xorshift(m, SEED + 100 * rep + L, P0)
        want = oracle.ntt_forward(a, P0, G0) if L <= 24 else None
        full = dev(a)
        blocks = [full.view(n0, n1)[:, r * n1 // G:(r + 1) * n1 // G].contiguous().view(-1) for r in range(G)]
        outs = [torch.full((m // G,), 0x5555555555555555, dtype=torch.int64, device="cuda") for _ in range(G)]
        torch.cuda.synchronize()
        mg.forward([o.data_ptr() for o in outs], [b.data_ptr() for b in blocks])
        mg.synchronize()
        got = host(torch.cat(outs))
        if want is not None:
            assert np.array_equal(got, want), (L, G, rep)
        else:
            # 2^26: against the single-GPU plan (itself compared with the oracle in test_three_pass_full_compare_2p26)
            ref = cuda_lib.plan(L)
            w = torch.empty_like(full)
            ref.forward(w.data_ptr(), full.data_ptr(), stream())
            torch.cuda.synchronize()
            assert np.array_equal(got, host(w)), (L, G, rep)
            ref.close()
        backs = [torch.empty_like(b) for b in blocks]
        mg.inverse([b.data_ptr() for b in backs], [o.data_ptr() for o in outs])
        mg.synchronize()
        for r in range(G):
            assert torch.equal(backs[r], blocks[r]), (L, G, rep, r)
    # whole transform through the host entry points (strided scatter / gather of the column blocks)
    if L <= 24:
        out = np.empty_like(a)
        mg.forward_host(out.ctypes.data, a.ctypes.data)
        assert np.array_equal(out, want)
        back = np.empty_like(a)
        mg.inverse_host(back.ctypes.data, out.ctypes.data)
        assert np.array_equal(back, a)
    mg.close()


@pytest.mark.parametrize("N,g,fixed", [(0xFFFFFFFF00000001, 7, False), (0x3A00000000000001, 3, False),
                                       (0x3A00000000000001, 3, True)])
def test_sharded_plan_other_moduli_on_one_gpu(cuda_lib, oracle, N, g, fixed):
    """The sharded path with runtime moduli (Montgomery and Shoup address-mapped kernels), all ranks on device 0."""
    for L, G, splits in [(20, 4, None), (21, 2, [7, 7, 7])]:
        m = 1 << L
        a = oracle.fill_xorshift(m, SEED + L, N)
        want = oracle.ntt_forward(a, N, g)
        mg = cuda_lib.mgpu(L, [0] * G, splits=splits, modulus=N, generator=g, fixed_point=fixed)
        out = np.empty_like(a)
        mg.forward_host(out.ctypes.data, a.ctypes.data)
        assert np.array_equal(out, want), (hex(N), L, G)
        back = np.empty_like(a)
        mg.inverse_host(back.ctypes.data, out.ctypes.data)
        assert np.array_equal(back, a)
        mg.close()


@pytest.mark.parametrize("L,G,K,splits", [(20, 2, 4, None), (22, 4, 2, None), (21, 8, 1, [7, 7, 7]), (24, 8, 4, None)])
def test_sharded_chunked_exchange_on_one_gpu(cuda_lib, oracle, L, G, K, splits):
    """The NCCL-pipelined variant's kernels (column pass per chunk into message layout, tiled row half) with all ranks
    on device 0 and the all-to-all played by tensor copies: same check as above."""
    import torch
    m = 1 << L
    plans = [cuda_lib.plan(L, splits=splits, shard_count=G, shard_rank=r) for r in range(G)]
    n0 = 1 << plans[0].splits[0]
    n1 = m // n0
    w = n1 // (G * K)
    msg = (n0 // G) * w
    a = oracle.fill_xorshift(m, SEED + L + G, P0)
    want = oracle.ntt_forward(a, P0, G0)
    full = dev(a)
    st = stream()
    blocks = [full.view(n0, n1)[:, r * n1 // G:(r + 1) * n1 // G].contiguous().view(-1) for r in range(G)]
    send = [torch.empty(m // G, dtype=torch.int64, device="cuda") for _ in range(G)]
    for r in range(G):
        for c in range(K):
            plans[r].shard_forward_cols_chunk(send[r].data_ptr(), blocks[r].data_ptr(), c, K, st)
    recv = [torch.empty(m // G, dtype=torch.int64, device="cuda") for _ in range(G)]
    for r in range(G):
        for s_ in range(G):
            recv[s_].view(K, G, msg)[:, r, :] = send[r].view(K, G, msg)[:, s_, :]
    outs = []
    for r in range(G):
        o = torch.empty(m // G, dtype=torch.int64, device="cuda")
        plans[r].shard_forward_rows_tiled(o.data_ptr(), recv[r].data_ptr(), K, st)
        outs.append(o)
    assert np.array_equal(host(torch.cat(outs)), want), (L, G, K)
    tiles = [torch.empty(m // G, dtype=torch.int64, device="cuda") for _ in range(G)]
    work = torch.empty(m // G, dtype=torch.int64, device="cuda")
    for r in range(G):
        plans[r].shard_inverse_rows_tiled(tiles[r].data_ptr(), outs[r].data_ptr(), work.data_ptr(), K, st)
    back = [torch.empty(m // G, dtype=torch.int64, device="cuda") for _ in range(G)]
    for r in range(G):
        for s_ in range(G):
            back[s_].view(K, G, msg)[:, r, :] = tiles[r].view(K, G, msg)[:, s_, :]
    for r in range(G):
        got = torch.empty(m // G, dtype=torch.int64, device="cuda")
        for c in range(K):
            plans[r].shard_inverse_cols_chunk(got.data_ptr(), back[r].data_ptr(), c, K, st)
        assert torch.equal(got, blocks[r]), (L, G, K, r)
    for p_ in plans:
        p_.close()


def test_error_paths(cuda_lib, pkg):
    with pytest.raises(pkg.XnttError) as e:
        cuda_lib.plan(10, splits=[4, 4])
    assert e.value.status == pkg.ERR_INVALID
    with pytest.raises(pkg.XnttError) as e:
        cuda_lib.plan(20, modulus=0x10001)
    assert e.value.status == pkg.ERR_INVALID
    plan = cuda_lib.plan(6, forward=False)
    import torch
    buf = torch.zeros(64, dtype=torch.int64, device="cuda")
    with pytest.raises(pkg.XnttError) as e:
        plan.forward(buf.data_ptr(), buf.data_ptr(), stream())
    assert e.value.status == pkg.ERR_STATE
    plan.close()


OTHER_MODULI = [
    (0x3A00000000000001, 3), (0xFFFFFFFF00000001, 7), (0xFFFFFFFF00000001, 0xF44872F5EC1C4CC0),
    (0xA3B25F400C7A8001, 5), (0x41D33D0D1FBF8001, 6), (0x3164C5D59B090001, 13), (0x1E4A0E19E4548001, 3),
    (0x08AA90297F870001, 3), (0x0000000000010001, 3), (0x0C40000000000001, 5), (0x0002580000000001, 11),
]


@pytest.mark.parametrize("N,g", OTHER_MODULI)
def test_other_moduli(cuda_lib, oracle, N, g):
    """The moduli of tests/ntt-tests, tests/test-ntt-reference.cpp:17-23 and
    examples/magic-series/test-magic-series.cpp:22-39 through the runtime-modulus kernels."""
    import torch
    for L, splits, batch in [(1, None, 1), (3, None, 5), (7, None, 1), (10, None, 1), (12, None, 2), (13, None, 1),
                             (15, None, 1), (13, [9, 4], 1), (16, [5, 5, 6], 1), (20, None, 1)]:
        if (N - 1) % (1 << L):
            continue
        m = 1 << L
        a = oracle.fill_xorshift(m * batch, SEED + L, N)
        plan = cuda_lib.plan(L, modulus=N, generator=g, splits=splits, batch=batch)
        src = dev(a)
        dst = torch.empty_like(src)
        plan.forward(dst.data_ptr(), src.data_ptr(), stream())
        got = host(dst)
        for b in range(batch):
            assert np.array_equal(got[b * m:(b + 1) * m], oracle.ntt_forward(a[b * m:(b + 1) * m].copy(), N, g)), (L, b)
        assert (got < np.uint64(N)).all()
        plan.inverse(dst.data_ptr(), dst.data_ptr(), stream())
        assert np.array_equal(host(dst), a)
        bm, prod = torch.empty_like(src), torch.empty_like(src)
        plan.to_montgomery(bm.data_ptr(), src.data_ptr(), a.size, stream())
        plan.multiply_normalize(prod.data_ptr(), dev(got).data_ptr(), bm.data_ptr(), a.size, stream())
        assert np.array_equal(host(prod), oracle.pointwise_mul(got, a, N))
        plan.close()


@pytest.mark.parametrize("N,g", [(N, g) for N, g in OTHER_MODULI if N < (1 << 62)])
def test_fixed_point_modmul(cuda_lib, oracle, N, g):
    """XNTT_MODMUL_FIXED_POINT: the Shoup kernels (FixedPoint64SVE, modmul/sve/fixed-point-64.hpp:13-69; csrc/field.cuh
    FieldShoup) for moduli below 2^62 against the oracle - every pass kind, scaled inverse, batches, 2^24."""
    import torch
    for L, splits, batch, kw in [(1, None, 1, {}), (3, None, 5, {}), (7, None, 1, {}), (10, None, 3, {}), (13, None, 2, {}),
                                 (15, None, 1, {}), (13, [9, 4], 1, {}), (16, [5, 5, 6], 1, {}),
                                 (14, None, 1, {"compact_tables": True}), (12, None, 2, {"inverse_factor": 12345}),
                                 (20, None, 1, {}), (24, None, 1, {})]:
        if (N - 1) % (1 << L):
            continue
        m = 1 << L
        a = oracle.fill_xorshift(m * batch, SEED + L, N)
        plan = cuda_lib.plan(L, modulus=N, generator=g, splits=splits, batch=batch, fixed_point=True, **kw)
        assert plan.modmul == 1
        src = dev(a)
        dst = torch.full_like(src, 0x5555555555555555)
        plan.forward(dst.data_ptr(), src.data_ptr(), stream())
        got = host(dst)
        for b in range(batch):
            assert np.array_equal(got[b * m:(b + 1) * m], oracle.ntt_forward(a[b * m:(b + 1) * m].copy(), N, g)), (hex(N), L, b)
        back = torch.empty_like(src)
        plan.inverse(back.data_ptr(), dst.data_ptr(), stream())
        f = kw.get("inverse_factor", m)
        scale = np.full_like(a, (m * pow(f, -1, N)) % N)
        assert np.array_equal(host(back), oracle.pointwise_mul(a, scale, N)), (hex(N), L)
        if batch == 1 and L >= 3:
            bm, fm = torch.empty_like(src), torch.empty_like(src)
            plan.to_montgomery(bm.data_ptr(), src.data_ptr(), m, stream())
            plan.forward_multiply(fm.data_ptr(), src.data_ptr(), bm.data_ptr(), stream())
            assert np.array_equal(host(fm), oracle.pointwise_mul(got, a, N)), (hex(N), L)
        plan.close()


@pytest.mark.parametrize("L,splits,N,g", [(14, None, P0, G0), (13, [9, 4], 0x3A00000000000001, 3), (20, None, P0, G0),
                                         (24, None, P0, G0)])
def test_fused_forward_multiply(cuda_lib, oracle, L, splits, N, g):
    import torch
    m = 1 << L
    rng = np.random.default_rng(L)
    a = rng.integers(0, N, m, dtype=np.uint64)
    b = rng.integers(0, N, m, dtype=np.uint64)
    plan = cuda_lib.plan(L, modulus=N, generator=g, splits=splits)
    da, db = dev(a), dev(b)
    fb, unfused, fused = torch.empty_like(db), torch.empty_like(da), torch.empty_like(da)
    plan.forward(fb.data_ptr(), db.data_ptr(), stream())
    plan.to_montgomery(fb.data_ptr(), fb.data_ptr(), m, stream())
    plan.forward(unfused.data_ptr(), da.data_ptr(), stream())
    plan.multiply_normalize(unfused.data_ptr(), unfused.data_ptr(), fb.data_ptr(), m, stream())
    plan.forward_multiply(fused.data_ptr(), da.data_ptr(), fb.data_ptr(), stream())
    assert torch.equal(fused, unfused)
    if L <= 20:
        want = oracle.pointwise_mul(oracle.ntt_forward(a, N, g), oracle.ntt_forward(b, N, g), N)
        assert np.array_equal(host(fused), want)
    plan.close()


@pytest.mark.parametrize("L,splits,batch", [(5, None, 3), (9, None, 1), (11, None, 37), (14, None, 1), (17, [8, 9], 1), (17, None, 3),
                                            (18, None, 1), (19, None, 1), (20, None, 1), (20, None, 2), (21, None, 1),
                                            (16, [5, 5, 6], 2), (15, [9, 6], 1), (13, [9, 4], 3), (19, [8, 11], 1)])
def test_tile_shapes(cuda_lib, oracle, L, splits, batch):
    """Whole tiles (2^13 residues per CTA) against narrow tiles (a quarter: plans of at most 2^20 residues take them by
    default, XNTT_TILES_WIDE / XNTT_TILES_NARROW force either): both against the oracle, forward and inverse, with both
    twiddle forms."""
    import torch
    m = 1 << L
    a = oracle.fill_xorshift(m * batch, SEED + 3 * L, P0)
    want = np.concatenate([oracle.ntt_forward(a[b * m:(b + 1) * m].copy(), P0, G0) for b in range(batch)])
    src = dev(a)
    shapes = {}
    for tiles, compact in (("wide", False), ("narrow", False), ("narrow", True), (None, False)):
        plan = cuda_lib.plan(L, splits=splits, batch=batch, tiles=tiles, compact_tables=compact)
        shapes[(tiles, compact)] = plan.tile_log2
        dst = torch.full_like(src, 0x5555555555555555)
        plan.forward(dst.data_ptr(), src.data_ptr(), stream())
        assert np.array_equal(host(dst), want), (L, splits, tiles, compact, plan.splits, plan.tile_log2)
        plan.inverse(dst.data_ptr(), dst.data_ptr(), stream())
        assert np.array_equal(host(dst), a), (L, splits, tiles, compact)
        plan.close()
    if splits is not None:  # (the planner's own decomposition differs between small and large plans)
        small = (m * batch) <= (1 << 20)
        assert shapes[(None, False)] == (shapes[("narrow", False)] if small else shapes[("wide", False)])


def test_tile_shapes_runtime_modulus(cuda_lib, oracle):
    """narrow tiles of the runtime-modulus (Montgomery) kernels; Goldilocks and Shoup plans keep whole tiles"""
    import torch
    for N, g, fixed, has in ((0x3A00000000000001, 3, False, True), (0x0C40000000000001, None, False, True),
                             (0xFFFFFFFF00000001, 7, False, False), (0x3A00000000000001, 3, True, False)):
        if g is None:
            g = dict(OTHER_MODULI)[N]
        for L in (12, 17, 20):
            m = 1 << L
            a = oracle.fill_xorshift(m, SEED + L, N)
            want = oracle.ntt_forward(a.copy(), N, g)
            plan = cuda_lib.plan(L, modulus=N, generator=g, fixed_point=fixed)
            if has and L >= 14:
                assert max(plan.tile_log2) <= 11, (hex(N), L, plan.tile_log2)
            if not has:
                assert plan.tile_log2 == cuda_lib.plan(L, modulus=N, generator=g, fixed_point=fixed, tiles="wide").tile_log2
            src = dev(a)
            dst = torch.empty_like(src)
            plan.forward(dst.data_ptr(), src.data_ptr(), stream())
            assert np.array_equal(host(dst), want), (hex(N), L)
            plan.inverse(dst.data_ptr(), dst.data_ptr(), stream())
            assert np.array_equal(host(dst), a), (hex(N), L)
            plan.close()


@pytest.mark.parametrize("L,batch,tiles", [(10, 1, None), (13, 1, None), (17, 1, None), (19, 1, "wide"), (20, 1, None), (21, 1, None),
                                           (24, 1, None)])
def test_dependent_launch_chains(cuda_lib, L, batch, tiles):
    """Every pass is launched as a programmatic dependent of the kernel before it: a chain of in-place transforms enqueued
    back to back must equal the same chain with a synchronisation after every call, and forward / inverse ping-pongs (in
    place and between two buffers) must return their input (tools/stress_chain.py is the long version)."""
    import torch
    m = 1 << L
    reps = 60 if L <= 21 else 10
    plan = cuda_lib.plan(L, batch=batch, tiles=tiles)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(L)
    x0 = torch.randint(0, 2**62, (m * batch,), dtype=torch.int64, device="cuda", generator=gen)
    x, y = x0.clone(), x0.clone()
    for _ in range(reps):
        plan.forward(x.data_ptr(), x.data_ptr(), stream())
    for _ in range(reps):
        plan.forward(y.data_ptr(), y.data_ptr(), stream())
        torch.cuda.synchronize()
    assert torch.equal(x, y)
    p, q = x0.clone(), torch.empty_like(x0)
    for _ in range(reps):
        plan.forward(q.data_ptr(), p.data_ptr(), stream())
        plan.inverse(p.data_ptr(), q.data_ptr(), stream())
        plan.forward(p.data_ptr(), p.data_ptr(), stream())
        plan.inverse(p.data_ptr(), p.data_ptr(), stream())
    torch.cuda.synchronize()
    assert torch.equal(p, x0)
    plan.close()


@pytest.mark.parametrize("seed", [21, 22])
def test_random_plans_on_gpu(cuda_lib, pkg, oracle, seed):
    """The emulator's fuzz (tests/test_emu_fuzz.py: random length up to 2^19, explicit splits, batch, modulus, twiddle form,
    tile shape, inverse_factor from a fixed seed) on the GPU: forward, scaled inverse and the fused point-wise product,
    word for word against the oracle."""
    import random
    import torch
    from test_emu_fuzz import draw
    rng = random.Random(seed)
    done = 0
    while done < 100:
        c = draw(rng, max_l=19)
        if c is None:
            continue
        L, N, g, batch = c["L"], c["N"], c["g"], c["batch"]
        m = 1 << L
        try:
            plan = cuda_lib.plan(L, modulus=N, generator=g, splits=c["splits"], batch=batch, compact_tables=c["compact"],
                                 inverse_factor=c["invf"], tiles=c["tiles"])
        except pkg.XnttError as e:
            assert e.status in (pkg.ERR_INVALID, pkg.ERR_UNSUPPORTED), c
            continue
        a = oracle.fill_xorshift(m * batch, c["seed"], N)
        src = dev(a)
        out = torch.full_like(src, 0x5555555555555555)
        plan.forward(out.data_ptr(), src.data_ptr(), stream())
        got = host(out)
        for b in range(batch):
            assert np.array_equal(got[b * m:(b + 1) * m], oracle.ntt_forward(a[b * m:(b + 1) * m].copy(), N, g)), (c, plan.splits, plan.tile_log2)
        back = torch.empty_like(src)
        plan.inverse(back.data_ptr(), out.data_ptr(), stream())
        f = m if c["invf"] is None else c["invf"]
        assert np.array_equal(host(back), oracle.pointwise_mul(a, np.full_like(a, (m * pow(f, -1, N)) % N), N)), c
        bm = oracle.fill_xorshift(m * batch, 7, N)
        fm = torch.empty_like(src)
        plan.forward_multiply(fm.data_ptr(), src.data_ptr(), dev(bm).data_ptr(), stream())
        want = oracle.pointwise_mul(oracle.pointwise_mul(got, bm, N), np.full_like(a, pow(1 << 64, -1, N)), N)
        assert np.array_equal(host(fm), want), c
        plan.close()
        done += 1


@pytest.mark.parametrize("L,batch", [(17, 1), (20, 3), (26, 1)])
def test_device_calls_capture_into_a_cuda_graph(cuda_lib, oracle, L, batch):
    """The device entry points only enqueue kernels on the caller's stream (no allocation, no synchronisation), so a
    forward + inverse pair can be captured into a CUDA graph and replayed: two- and three-pass plans, replayed on fresh
    data, bit-exact against the oracle."""
    import torch
    m = 1 << L
    plan = cuda_lib.plan(L, batch=batch)
    a = oracle.fill_xorshift(m * batch, SEED + 7 * L, P0)
    src, spec, back = dev(a), torch.empty(m * batch, dtype=torch.int64, device="cuda"), \
        torch.empty(m * batch, dtype=torch.int64, device="cuda")
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        # first launches set kernel attributes: keep them out of the capture
        plan.forward(spec.data_ptr(), src.data_ptr(), side.cuda_stream)
        plan.inverse(back.data_ptr(), spec.data_ptr(), side.cuda_stream)
    side.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        st = torch.cuda.current_stream().cuda_stream
        plan.forward(spec.data_ptr(), src.data_ptr(), st)
        plan.inverse(back.data_ptr(), spec.data_ptr(), st)
    for rep in range(2):
        b = oracle.fill_xorshift(m * batch, SEED + 11 * L + rep, P0)
        src.copy_(dev(b))
        spec.fill_(0x5555555555555555)
        back.fill_(0x2AAAAAAAAAAAAAAA)
        graph.replay()
        torch.cuda.synchronize()
        got = host(spec)
        if L <= 20:
            for i in range(batch):
                assert np.array_equal(got[i * m:(i + 1) * m], oracle.ntt_forward(b[i * m:(i + 1) * m].copy(), P0, G0)), (L, i, rep)
        else:
            # 2^26: directly evaluated output words
            for pos in (0, 1, 12345, m - 1):
                assert int(got[pos]) == oracle.dft_point(b, P0, G0, pos), (L, pos, rep)
        assert np.array_equal(host(back), b), (L, rep)
    plan.close()


def test_transpose_contract(cuda_lib):
    """tests/bench-transpose.cpp: every shape transposed and transposed back must return the input;
    sizes 2^8..2^13, pads {0, 32}, plus ragged shapes and the in-place square form."""
    import torch
    gen = torch.Generator(device="cuda")
    gen.manual_seed(3)
    shapes = [(1 << a, 1 << b, ps, pd) for a in (8, 10, 13) for b in (8, 11, 13) for ps, pd in ((0, 0), (32, 32))]
    shapes += [(100, 37, 5, 3), (63, 65, 0, 0), (1, 1, 0, 0), (4096, 4096, 0, 0)]
    for rows, cols, ps, pd in shapes:
        src = torch.randint(0, 2**62, (rows, cols + ps), dtype=torch.int64, device="cuda", generator=gen)
        dst = torch.full((cols, rows + pd), 0x55, dtype=torch.int64, device="cuda")
        cuda_lib.transpose(dst.data_ptr(), src.data_ptr(), rows, cols, rows + pd, cols + ps, stream())
        assert torch.equal(dst[:, :rows], src[:, :cols].t())
        assert bool((dst[:, rows:] == 0x55).all())
    for dim, pad in [(64, 0), (200, 8), (2048, 0), (513, 0)]:
        a = torch.randint(0, 2**62, (dim, dim + pad), dtype=torch.int64, device="cuda", generator=gen)
        b = a.clone()
        cuda_lib.transpose(b.data_ptr(), b.data_ptr(), dim, dim, dim + pad, dim + pad, stream())
        assert torch.equal(b[:, :dim], a[:, :dim].t())


def test_decompositions_agree_at_2p30(cuda_lib):
    """BASELINE configs[3] size on one GPU: two different three-pass decompositions of n = 2^30 must
    produce the same 8 GiB of output, and the inverse must return the input."""
    import torch
    L = 30
    m = 1 << L
    free, _ = torch.cuda.mem_get_info()
    if free < 30 * 2**30:
        pytest.skip("needs ~26 GiB of device memory")
    gen = torch.Generator(device="cuda")
    gen.manual_seed(30)
    src = torch.randint(0, 2**62, (m,), dtype=torch.int64, device="cuda", generator=gen)
    a, b = torch.empty_like(src), torch.empty_like(src)
    p1 = cuda_lib.plan(L)
    p2 = cuda_lib.plan(L, splits=[10, 10, 10])
    assert p1.splits != p2.splits
    p1.forward(a.data_ptr(), src.data_ptr(), stream())
    p2.forward(b.data_ptr(), src.data_ptr(), stream())
    assert torch.equal(a, b)
    p2.inverse(b.data_ptr(), b.data_ptr(), stream())
    assert torch.equal(b, src)
    p1.close()
    p2.close()


def _dist_worker(rank, world, port, log2_m, q):
    import os
    import sys
    import torch
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "sve-ntt_b200"))
    import __graft_entry__ as ge
    pkg = ge.load_package()
    import dist_ntt
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        lib = pkg.load()
        dev = torch.device("cuda", rank)
        st = torch.cuda.current_stream().cuda_stream
        m = 1 << log2_m
        res = []
        for mode in ("peer", "pipelined", "simple"):
            sh = dist_ntt.ShardedNTT(lib, log2_m, world, rank, device=rank, mode=mode)
            gen = torch.Generator(device=dev)
            gen.manual_seed(77)
            full = torch.randint(0, 2**62, (m,), dtype=torch.int64, device=dev, generator=gen)
            n0, n1 = sh.n0, sh.n1
            src = full.view(n0, n1)[:, rank * n1 // world:(rank + 1) * n1 // world].contiguous().view(-1)
            dst = torch.empty_like(src)
            sh.forward(dst, src, st)
            ref = lib.plan(log2_m, device=rank)
            want = torch.empty_like(full)
            ref.forward(want.data_ptr(), full.data_ptr(), st)
            ok_f = bool(torch.equal(dst, want[rank * m // world:(rank + 1) * m // world]))
            back = torch.empty_like(src)
            sh.inverse(back, dst, st)
            ok_i = bool(torch.equal(back, src))
            res.append((mode, sh.mode, ok_f, ok_i))
            ref.close()
            sh.close()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_sharded_transform_on_two_gpus():
    """N > 1 on real devices (skipped on a single-GPU box): the sharded transform, in all three exchange
    modes, equals the single-GPU plan word for word."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, 22, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    for _ in range(2):
        rank, res = q.get(timeout=5)
        for want_mode, got_mode, ok_f, ok_i in res:
            assert ok_f and ok_i, (rank, want_mode, got_mode)


def test_two_devices_in_one_process(cuda_lib, oracle):
    """xntt_desc::device: plans on different GPUs of one process (kernel attributes are per device)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    L = 16
    a = oracle.fill_xorshift(1 << L, SEED + 77, P0)
    want = oracle.ntt_forward(a, P0, G0)
    for dev_id in (1, 0):
        plan = cuda_lib.plan(L, device=dev_id)
        with torch.cuda.device(dev_id):
            d = torch.from_numpy(a.view(np.int64)).cuda(dev_id)
            o = torch.empty_like(d)
            plan.forward(o.data_ptr(), d.data_ptr(), torch.cuda.current_stream(dev_id).cuda_stream)
            torch.cuda.synchronize(dev_id)
            assert np.array_equal(o.cpu().numpy().view(np.uint64), want), dev_id
            t = torch.empty((256, 128), dtype=torch.int64, device=f"cuda:{dev_id}")
            src = torch.arange(128 * 256, dtype=torch.int64, device=f"cuda:{dev_id}").view(128, 256)
            cuda_lib.transpose(t.data_ptr(), src.data_ptr(), 128, 256, 128, 256,
                               torch.cuda.current_stream(dev_id).cuda_stream)
            torch.cuda.synchronize(dev_id)
            assert torch.equal(t, src.t())
        plan.close()
