# SPDX-License-Identifier: Apache-2.0
"""Planner + kernel index algebra on the CPU: the host emulator (tests/emu) executes the same
kernel templates as the GPU, one emulated thread at a time, and must match the oracle word for
word.  Mirrors the shapes of the reference's ntt-tests (tests/ntt-tests/*.hpp: 2^5 .. 2^15,
iterative / recursive / four-step) and the README example (2^17 = 2^8 x 2^9)."""
import numpy as np
import pytest

from conftest import G0, P0, SEED


def roundtrip(emu, oracle, L, splits=None, batch=1, inverse_factor=None, **kw):
    m = 1 << L
    a = oracle.fill_xorshift(m * batch, SEED + L, P0)
    plan = emu.plan(L, splits=splits, batch=batch, inverse_factor=inverse_factor, **kw)
    out = np.empty_like(a)
    plan.forward(out.ctypes.data, a.ctypes.data)
    for b in range(batch):
        want = oracle.ntt_forward(a[b * m:(b + 1) * m].copy(), P0, G0)
        assert np.array_equal(out[b * m:(b + 1) * m], want), (L, splits, b)
    back = np.empty_like(a)
    plan.inverse(back.ctypes.data, out.ctypes.data)
    f = m if inverse_factor is None else inverse_factor
    scale = np.full_like(a, (m * pow(f, -1, P0)) % P0)
    assert np.array_equal(back, oracle.pointwise_mul(a, scale, P0)), (L, splits)
    inplace = a.copy()
    plan.forward(inplace.ctypes.data, inplace.ctypes.data)
    assert np.array_equal(inplace, out)
    plan.inverse(inplace.ctypes.data, inplace.ctypes.data)
    assert np.array_equal(inplace, back)
    return plan.splits


@pytest.mark.parametrize("L", range(1, 14))
def test_single_pass_sizes(emu, oracle, L):
    # one row pass; a lone 2^12 / 2^13 transform is cut in two so that narrow tiles spread it over several CTAs
    assert roundtrip(emu, oracle, L) == ([L] if L < 12 else [(L + 1) // 2, L // 2])
    assert roundtrip(emu, oracle, L, tiles="wide") == [L]
    assert roundtrip(emu, oracle, L, splits=[L]) == [L]


@pytest.mark.parametrize("L", [14, 15, 16, 17, 18, 20])
def test_default_two_pass(emu, oracle, L):
    assert len(roundtrip(emu, oracle, L)) == 2


@pytest.mark.parametrize("L,splits", [
    (17, [8, 9]),    # README.md:28-68  blocked six-step 2^8 x 2^9
    (13, [9, 4]),    # ntt-tests/recursive-scalar-fourstep-two13.hpp
    (15, [9, 6]),    # ntt-tests/recursive-sve-fourstep-two13.hpp (m = 2^15)
    (13, [1, 12]), (14, [12, 2]), (10, [5, 5]), (16, [5, 5, 6]), (18, [6, 6, 6]), (13, [4, 4, 5]),
])
def test_explicit_splits(emu, oracle, L, splits):
    assert roundtrip(emu, oracle, L, splits=splits) == splits


@pytest.mark.parametrize("L,splits,batch", [(5, None, 3), (9, None, 1), (11, None, 5), (14, None, 1), (17, [8, 9], 1), (18, None, 1),
                                            (20, None, 1), (16, [5, 5, 6], 2), (15, [9, 6], 1), (13, [9, 4], 3), (19, [8, 11], 1)])
def test_tile_shapes(emu, oracle, L, splits, batch):
    """Every pass exists on whole tiles (2^13 residues per CTA) and, up to 2^11 rows / 2^9 columns, on narrow tiles (a
    quarter of that): XNTT_TILES_WIDE / XNTT_TILES_NARROW force either, the default picks narrow for plans of at most
    2^20 residues.  Same words from all three."""
    roundtrip(emu, oracle, L, splits=splits, batch=batch, tiles="wide")
    roundtrip(emu, oracle, L, splits=splits, batch=batch, tiles="narrow")
    roundtrip(emu, oracle, L, splits=splits, batch=batch, tiles="narrow", compact_tables=True, inverse_factor=5)
    wide = emu.plan(L, splits=splits, batch=batch, tiles="wide")
    narrow = emu.plan(L, splits=splits, batch=batch, tiles="narrow")
    auto = emu.plan(L, splits=splits, batch=batch)
    assert all(t <= 13 for t in wide.tile_log2) and all(t <= 11 or n > 11 for t, n in zip(narrow.tile_log2, narrow.splits))
    assert any(a < b for a, b in zip(narrow.tile_log2, wide.tile_log2)) or all(n >= 12 for n in narrow.splits)
    assert auto.tile_log2 == narrow.tile_log2  # all of these are small plans
    if splits is None and 14 <= L <= 20:
        assert auto.splits[0] <= 9 and auto.splits[1] <= 11  # both passes have a narrow tile
    for pl in (wide, narrow, auto):
        pl.close()


def test_tile_shape_rule_for_large_plans(emu):
    """Above 2^20 residues the planner keeps whole tiles and its large-plan decomposition."""
    big = emu.plan(22)
    assert big.splits == [10, 12] and big.tile_log2 == [13, 13]
    batched = emu.plan(12, batch=1024)
    assert batched.tile_log2 == [13]
    small = emu.plan(10, batch=4)
    assert small.tile_log2 == [11]
    no_variant = emu.plan(12, batch=64)  # a 2^12 row pass has no narrow tile; 64 of them are enough CTAs
    assert no_variant.splits == [12] and no_variant.tile_log2 == [13]
    no_variant.close()
    few = emu.plan(13, batch=16)
    assert few.splits == [7, 6] and max(few.tile_log2) <= 11
    few.close()
    for pl in (big, batched, small):
        pl.close()


@pytest.mark.parametrize("L,splits", [(14, None), (17, [8, 9]), (16, [5, 5, 6]), (13, [1, 12])])
def test_compact_twiddle_tables(emu, oracle, L, splits):
    """The six-step twiddles exist in two forms (whole matrix / two sqrt(M) tables, XNTT_COMPACT_TABLES): the
    default plans above run the first, these the second - by flag and by a table budget (xntt_desc::twist_table_max_mb)
    that holds one direction's matrix but not both (L = 17: 2 MiB each) or neither."""
    assert len(roundtrip(emu, oracle, L, splits=splits, compact_tables=True)) >= 2
    roundtrip(emu, oracle, L, splits=splits, compact_tables=True, inverse_factor=777)
    roundtrip(emu, oracle, L, splits=splits, twist_table_max_mb=1)
    roundtrip(emu, oracle, L, splits=splits, twist_table_max_mb=3)


@pytest.mark.parametrize("L,batch", [(3, 1), (3, 33), (6, 5), (10, 7), (12, 3), (13, 2), (15, 3)])
def test_batches_including_ragged_tiles(emu, oracle, L, batch):
    roundtrip(emu, oracle, L, batch=batch)


@pytest.mark.parametrize("L,splits", [(9, None), (12, None), (14, None), (16, [5, 5, 6])])
def test_inverse_factor_variants(emu, oracle, L, splits):
    roundtrip(emu, oracle, L, splits=splits, inverse_factor=1)       # README example: unscaled inverse
    roundtrip(emu, oracle, L, splits=splits, inverse_factor=12345)   # arbitrary inverse_factor


def test_out_of_place_keeps_source(emu, oracle):
    a = oracle.fill_xorshift(1 << 14, SEED, P0)
    keep = a.copy()
    plan = emu.plan(14)
    out = np.empty_like(a)
    plan.forward(out.ctypes.data, a.ctypes.data)
    assert np.array_equal(a, keep)


def test_edge_inputs(emu, oracle):
    """All-zero, all p-1, a delta and a constant vector (values the lazy arithmetic must survive)."""
    for L, splits in [(12, None), (14, None)]:
        m = 1 << L
        plan = emu.plan(L, splits=splits)
        for a in (np.zeros(m, np.uint64), np.full(m, P0 - 1, np.uint64),
                  np.eye(1, m, 0, dtype=np.uint64)[0] * np.uint64(P0 - 1), np.full(m, 1, np.uint64)):
            a = np.ascontiguousarray(a)
            out = np.empty_like(a)
            plan.forward(out.ctypes.data, a.ctypes.data)
            assert np.array_equal(out, oracle.ntt_forward(a, P0, G0))
            assert (out < np.uint64(P0)).all()
            back = np.empty_like(a)
            plan.inverse(back.ctypes.data, out.ctypes.data)
            assert np.array_equal(back, a)


def test_padic64_elementwise(emu, oracle):
    rng = np.random.default_rng(5)
    edge = np.array([0, 1, 2, P0 - 1, P0 - 2, 2**63, 2**32, 2**32 - 1, 0x3917FFFFFFF], dtype=np.uint64)
    a = np.concatenate([np.repeat(edge, edge.size), rng.integers(0, P0, 4000, dtype=np.uint64)])
    b = np.concatenate([np.tile(edge, edge.size), rng.integers(0, P0, 4000, dtype=np.uint64)])
    plan = emu.plan(4)
    bm, back, prod = np.empty_like(b), np.empty_like(b), np.empty_like(a)
    plan.to_montgomery(bm.ctypes.data, b.ctypes.data, b.size)
    assert all(int(x) == oracle.to_montgomery(int(y), P0) for x, y in zip(bm[:200], b[:200]))
    plan.from_montgomery(back.ctypes.data, bm.ctypes.data, b.size)
    assert np.array_equal(back, b)
    plan.multiply_normalize(prod.ctypes.data, a.ctypes.data, bm.ctypes.data, a.size)
    assert np.array_equal(prod, oracle.pointwise_mul(a, b, P0))


def test_polynomial_multiply(emu, oracle):
    """forward, point-wise multiply_normalize against a to_montgomery'd spectrum, inverse
    (examples/magic-series/gaussian-polynomial.hpp:176-214) == schoolbook cyclic convolution."""
    L = 8
    m = 1 << L
    rng = np.random.default_rng(7)
    a = np.zeros(m, np.uint64)
    b = np.zeros(m, np.uint64)
    a[:m // 2] = rng.integers(0, P0, m // 2, dtype=np.uint64)
    b[:m // 2] = rng.integers(0, P0, m // 2, dtype=np.uint64)
    plan = emu.plan(L)
    fa, fb = np.empty_like(a), np.empty_like(b)
    plan.forward(fa.ctypes.data, a.ctypes.data)
    plan.forward(fb.ctypes.data, b.ctypes.data)
    plan.to_montgomery(fb.ctypes.data, fb.ctypes.data, m)
    plan.multiply_normalize(fa.ctypes.data, fa.ctypes.data, fb.ctypes.data, m)
    plan.inverse(fa.ctypes.data, fa.ctypes.data)
    want = [0] * m
    ai, bi = [int(v) for v in a[:m // 2]], [int(v) for v in b[:m // 2]]
    for i, x in enumerate(ai):
        for j, y in enumerate(bi):
            want[i + j] = (want[i + j] + x * y) % P0
    assert [int(v) for v in fa] == want


def test_error_paths(emu, pkg):
    """Same failure classes as the reference: invalid_argument for impossible shapes/roots,
    logic_error for a direction that was not prepared."""
    with pytest.raises(pkg.XnttError) as e:
        emu.plan(40)
    assert e.value.status == pkg.ERR_INVALID
    with pytest.raises(pkg.XnttError) as e:
        emu.plan(10, splits=[4, 4])  # product of radices != m (iterative.hpp:24-27)
    assert e.value.status == pkg.ERR_INVALID
    with pytest.raises(pkg.XnttError) as e:
        emu.plan(10, modulus=0xFFFFFC6E80000001 - 2)  # even / not prime
    assert e.value.status == pkg.ERR_INVALID
    with pytest.raises(pkg.XnttError) as e:
        emu.plan(20, modulus=0x10001)  # 2^20 does not divide p - 1: the field has no such root
    assert e.value.status == pkg.ERR_INVALID
    with pytest.raises(pkg.XnttError) as e:
        emu.plan(10, generator=1)  # not a generator of the order-m subgroup
    assert e.value.status == pkg.ERR_INVALID
    plan = emu.plan(6, inverse=False)
    buf = np.zeros(64, np.uint64)
    plan.forward(buf.ctypes.data, buf.ctypes.data)
    with pytest.raises(pkg.XnttError) as e:
        plan.inverse(buf.ctypes.data, buf.ctypes.data)
    assert e.value.status == pkg.ERR_STATE


def test_sharded_plan_matches_single(emu, oracle):
    """Distributed six-step with the exchange done by numpy: every rank runs
    the column half on its column block, tiles are exchanged all-to-all, every rank runs the row
    half; the concatenated result equals the oracle's transform.  Inverse mirrors it."""
    for L, splits, G in [(14, [7, 7], 2), (16, [6, 5, 5], 4), (16, [8, 8], 8)]:
        m = 1 << L
        n0 = 1 << splits[0]
        n1 = m // n0
        a = oracle.fill_xorshift(m, SEED + 3, P0)
        want = oracle.ntt_forward(a, P0, G0)
        A = a.reshape(n0, n1)
        plans = [emu.plan(L, splits=splits, shard_count=G, shard_rank=r) for r in range(G)]
        cols = []
        for r in range(G):
            blk = np.ascontiguousarray(A[:, r * n1 // G:(r + 1) * n1 // G])
            plans[r].shard_forward_cols(blk.ctypes.data, blk.ctypes.data)
            cols.append(blk)
        outs = []
        for r in range(G):
            rows = np.ascontiguousarray(
                np.concatenate([cols[s][r * n0 // G:(r + 1) * n0 // G, :] for s in range(G)], axis=1))
            plans[r].shard_forward_rows(rows.ctypes.data, rows.ctypes.data)
            outs.append(rows)
        got = np.concatenate([o.reshape(-1) for o in outs])
        assert np.array_equal(got, want), (L, splits, G)
        # inverse: rows first, exchange back, columns
        for r in range(G):
            plans[r].shard_inverse_rows(outs[r].ctypes.data, outs[r].ctypes.data)
        back = np.empty_like(A)
        for r in range(G):
            blk = np.ascontiguousarray(
                np.concatenate([outs[s][:, r * n1 // G:(r + 1) * n1 // G] for s in range(G)], axis=0))
            plans[r].shard_inverse_cols(blk.ctypes.data, blk.ctypes.data)
            back[:, r * n1 // G:(r + 1) * n1 // G] = blk
        assert np.array_equal(back.reshape(-1), a), (L, splits, G)


# moduli of the reference's own tests: ntt-tests (62 bit), test-ntt-reference.cpp:17-23 and
# examples/magic-series/test-magic-series.cpp:22-39 (64 .. 60 bit, Goldilocks, the Fermat prime)
OTHER_MODULI = [
    (0x3A00000000000001, 3), (0xFFFFFFFF00000001, 7), (0xFFFFFFFF00000001, 0xF44872F5EC1C4CC0),
    (0xA3B25F400C7A8001, 5), (0x41D33D0D1FBF8001, 6), (0x3164C5D59B090001, 13), (0x1E4A0E19E4548001, 3),
    (0x08AA90297F870001, 3), (0x0000000000010001, 3), (0x0C40000000000001, 5), (0x0C60000000000001, 7),
    (0x0003F00000000001, 11), (0x0002580000000001, 11),
]


@pytest.mark.parametrize("N,g", OTHER_MODULI)
def test_other_moduli(emu, oracle, N, g):
    """Modulus<p, g> is a template: every prime the reference's tests use runs through the
    runtime-modulus kernels."""
    for L, splits, batch in [(1, None, 1), (3, None, 5), (7, None, 1), (10, None, 1), (12, None, 2), (13, None, 1),
                             (15, None, 1), (13, [9, 4], 1), (16, [5, 5, 6], 1)]:
        if (N - 1) % (1 << L):
            continue
        m = 1 << L
        a = oracle.fill_xorshift(m * batch, SEED + L, N)
        plan = emu.plan(L, modulus=N, generator=g, splits=splits, batch=batch)
        out = np.empty_like(a)
        plan.forward(out.ctypes.data, a.ctypes.data)
        for b in range(batch):
            assert np.array_equal(out[b * m:(b + 1) * m], oracle.ntt_forward(a[b * m:(b + 1) * m].copy(), N, g))
        assert (out < np.uint64(N)).all()
        back = np.empty_like(a)
        plan.inverse(back.ctypes.data, out.ctypes.data)
        assert np.array_equal(back, a)
        # PAdic64 helpers in this field
        bm = np.empty_like(a)
        plan.to_montgomery(bm.ctypes.data, a.ctypes.data, a.size)
        prod = np.empty_like(a)
        plan.multiply_normalize(prod.ctypes.data, out.ctypes.data, bm.ctypes.data, a.size)
        assert np.array_equal(prod, oracle.pointwise_mul(out, a, N))


def test_golden_full_vectors_all_moduli(emu, golden):
    for case in golden["full"]:
        N, g = int(case["modulus"], 16), case["g"]
        a = np.array([int(v, 16) for v in case["input"]], dtype=np.uint64)
        plan = emu.plan(case["log2_m"], modulus=N, generator=g)
        out = np.empty_like(a)
        plan.forward(out.ctypes.data, a.ctypes.data)
        assert [f"{int(v):016x}" for v in out] == case["forward"]
        plan.inverse(out.ctypes.data, a.ctypes.data)
        assert [f"{int(v):016x}" for v in out] == case["inverse"]


@pytest.mark.parametrize("L,splits,N,g", [(8, None, P0, G0), (14, None, P0, G0), (13, [9, 4], 0x3A00000000000001, 3),
                                         (12, None, 0xFFFFFFFF00000001, 7), (16, [5, 5, 6], P0, G0)])
def test_fused_forward_multiply(emu, oracle, L, splits, N, g):
    """xntt_forward_multiply == compute_forward followed by the multiply_normalize loop
    (examples/magic-series/gaussian-polynomial.hpp:196-212), for every plan shape."""
    m = 1 << L
    rng = np.random.default_rng(L)
    a = rng.integers(0, N, m, dtype=np.uint64)
    b = rng.integers(0, N, m, dtype=np.uint64)
    plan = emu.plan(L, modulus=N, generator=g, splits=splits)
    fb = np.empty_like(b)
    plan.forward(fb.ctypes.data, b.ctypes.data)
    plan.to_montgomery(fb.ctypes.data, fb.ctypes.data, m)
    fused = np.empty_like(a)
    plan.forward_multiply(fused.ctypes.data, a.ctypes.data, fb.ctypes.data)
    want = oracle.pointwise_mul(oracle.ntt_forward(a, N, g), oracle.ntt_forward(b, N, g), N)
    assert np.array_equal(fused, want)
    plan.inverse(fused.ctypes.data, fused.ctypes.data)
    assert np.array_equal(fused, oracle.ntt_inverse(want, N, g))


@pytest.mark.parametrize("L,splits,G,K", [(14, [7, 7], 2, 1), (14, [7, 7], 2, 2), (16, [6, 5, 5], 4, 1), (18, [6, 6, 6], 2, 2),
                                         (18, [7, 11], 4, 4), (17, [8, 4, 5], 2, 2)])
def test_sharded_tiled_variants(emu, oracle, L, splits, G, K):
    """Chunked column passes + tiled row half (no separate pack/unpack pass): numpy plays the
    all-to-all, chunk by chunk; result == oracle, and the inverse returns the column blocks."""
    m = 1 << L
    n0 = 1 << splits[0]
    n1 = m // n0
    w = n1 // (G * K)
    a = oracle.fill_xorshift(m, SEED + 11, P0)
    want = oracle.ntt_forward(a, P0, G0)
    A = a.reshape(n0, n1)
    plans = [emu.plan(L, splits=splits, shard_count=G, shard_rank=r) for r in range(G)]
    blocks = [np.ascontiguousarray(A[:, r * n1 // G:(r + 1) * n1 // G]).reshape(-1) for r in range(G)]
    send = [np.full(m // G, 0xDEAD, np.uint64) for _ in range(G)]
    for r in range(G):
        for c in range(K):
            plans[r].shard_forward_cols_chunk(send[r].ctypes.data, blocks[r].ctypes.data, c, K)
    # all-to-all per chunk: message (chunk c, dest s) of rank r = send[r][c][s]
    recv = [np.empty(m // G, np.uint64) for _ in range(G)]
    msg = (n0 // G) * w
    for r in range(G):
        sv = send[r].reshape(K, G, msg)
        for s in range(G):
            recv[s].reshape(K, G, msg)[:, r, :] = sv[:, s, :]
    outs = []
    for r in range(G):
        dst = np.empty(m // G, np.uint64)
        plans[r].shard_forward_rows_tiled(dst.ctypes.data, recv[r].ctypes.data, K)
        outs.append(dst)
    assert np.array_equal(np.concatenate(outs), want), (L, splits, G, K)
    # inverse
    tiles = [np.empty(m // G, np.uint64) for _ in range(G)]
    work = [np.empty(m // G, np.uint64) for _ in range(G)]
    for r in range(G):
        plans[r].shard_inverse_rows_tiled(tiles[r].ctypes.data, outs[r].ctypes.data, work[r].ctypes.data, K)
    back_tiles = [np.empty(m // G, np.uint64) for _ in range(G)]
    for r in range(G):
        tv = tiles[r].reshape(K, G, msg)
        for s in range(G):
            back_tiles[s].reshape(K, G, msg)[:, r, :] = tv[:, s, :]
    for r in range(G):
        got = np.empty(m // G, np.uint64)
        for c in range(K):
            plans[r].shard_inverse_cols_chunk(got.ctypes.data, back_tiles[r].ctypes.data, c, K)
        assert np.array_equal(got, blocks[r]), (L, splits, G, K, r)


@pytest.mark.parametrize("L,splits,G", [
    (14, [7, 7], 2), (16, [6, 5, 5], 4), (18, [7, 11], 8), (17, [8, 4, 5], 2),
    # n0 == G: every output row of the column pass belongs to another rank (peer_bits == 0)
    (15, [3, 12], 8), (13, [1, 12], 2), (14, [2, 6, 6], 4),
    # inverse: one inner run of pass 1 per peer (block == n1b)
    (14, [4, 2, 8], 4), (16, [4, 4, 8], 4),
    # a 2^8 pass behind the exchange: the sharded first pass hands it its forward twiddle matrix (kColPre, row block)
    (18, [5, 8, 5], 2), (17, [3, 8, 6], 2),
])
def test_sharded_peer_store_variants(emu, oracle, L, splits, G):
    """Fused exchange: the pass next to the all-to-all stores straight into every rank's buffer
    (peer pointers).  All ranks live in this process, so 'peer memory' is just the other arrays."""
    m = 1 << L
    n0 = 1 << splits[0]
    n1 = m // n0
    a = oracle.fill_xorshift(m, SEED + 12, P0)
    want = oracle.ntt_forward(a, P0, G0)
    A = a.reshape(n0, n1)
    # default: the inverse of the sharded first pass applies its rank's column block of the twiddle matrix;
    # with compact tables (every other case here) the two-table form
    compact = (L + G) % 2 == 1 and splits[1:2] != [8]
    plans = [emu.plan(L, splits=splits, shard_count=G, shard_rank=r, compact_tables=compact) for r in range(G)]
    if len(splits) == 3 and splits[1] == 8:
        assert plans[0].twiddle_forms(False) == [3, 3, 0]
        assert plans[0].twiddle_forms(True) == [2, 3 if G < 4 else 2, 0]  # G >= 4: the link-bound pass keeps its own
    blocks = [np.ascontiguousarray(A[:, r * n1 // G:(r + 1) * n1 // G]).reshape(-1) for r in range(G)]
    bufs = [np.full(m // G, 0xDEAD, np.uint64) for _ in range(G)]
    peers = [b.ctypes.data for b in bufs]
    for r in range(G):
        plans[r].shard_forward_cols_peer(peers, blocks[r].ctypes.data)
    outs = []
    for r in range(G):
        dst = np.empty(m // G, np.uint64)
        plans[r].shard_forward_rows_tiled(dst.ctypes.data, bufs[r].ctypes.data, 1)
        outs.append(dst)
    assert np.array_equal(np.concatenate(outs), want), (L, splits, G)
    bufs2 = [np.full(m // G, 0xBEEF, np.uint64) for _ in range(G)]
    peers2 = [b.ctypes.data for b in bufs2]
    work = np.empty(m // G, np.uint64)
    for r in range(G):
        plans[r].shard_inverse_rows_peer(peers2, outs[r].ctypes.data, work.ctypes.data)
    for r in range(G):
        got = np.empty(m // G, np.uint64)
        plans[r].shard_inverse_cols_chunk(got.ctypes.data, bufs2[r].ctypes.data, 0, 1)
        assert np.array_equal(got, blocks[r]), (L, splits, G, r)


def test_transpose_contract(emu):
    """Transpose*::transpose: dst[ld_dst*c + r] = src[ld_src*r + c] with padded leading dimensions
    (tests/bench-transpose.cpp sweeps pads {0, 32}), ragged shapes, and the in-place square form."""
    rng = np.random.default_rng(0)
    for rows, cols, ps, pd in [(64, 64, 0, 0), (256, 512, 32, 0), (512, 256, 0, 32), (100, 37, 5, 3), (1, 1, 0, 0),
                               (128, 4096, 32, 32), (63, 65, 0, 0)]:
        src = rng.integers(0, 2**63, (rows, cols + ps), dtype=np.uint64)
        dst = np.full((cols, rows + pd), 0x5555555555555555, dtype=np.uint64)
        emu.transpose(dst.ctypes.data, src.ctypes.data, rows, cols, rows + pd, cols + ps)
        assert np.array_equal(dst[:, :rows], src[:, :cols].T)
        assert (dst[:, rows:] == 0x5555555555555555).all()  # padding untouched
        back = np.empty_like(src)
        emu.transpose(back.ctypes.data, dst.ctypes.data, cols, rows, cols + ps, rows + pd)
        assert np.array_equal(back[:, :cols], src[:, :cols])  # the reference's iota round-trip check
    for dim, pad in [(64, 0), (200, 8), (1, 0), (513, 0)]:
        a = rng.integers(0, 2**63, (dim, dim + pad), dtype=np.uint64)
        b = a.copy()
        emu.transpose(b.ctypes.data, b.ctypes.data, dim, dim, dim + pad, dim + pad)
        assert np.array_equal(b[:, :dim], a[:, :dim].T)


@pytest.mark.parametrize("L,batch,splits", [(10, 1, None), (13, 1, None), (13, 1, [13]), (11, 4, None), (5, 7, None), (12, 2, [5, 7]),
                                            (14, 1, None)])
def test_host_entry_points_small(emu, oracle, L, batch, splits):
    """Single-pass plans of up to 64 KiB run straight on the caller's mapped host buffers (the emulator counts every host
    buffer as mapped); everything else is staged.  Out of place and in place."""
    m = 1 << L
    a = oracle.fill_xorshift(m * batch, SEED + L, P0)
    want = np.concatenate([oracle.ntt_forward(a[b * m:(b + 1) * m].copy(), P0, G0) for b in range(batch)])
    plan = emu.plan(L, batch=batch, splits=splits)
    out, back = np.empty_like(a), np.empty_like(a)
    plan.forward_host(out.ctypes.data, a.ctypes.data)
    assert np.array_equal(out, want)
    plan.inverse_host(back.ctypes.data, out.ctypes.data)
    assert np.array_equal(back, a)
    buf = a.copy()
    plan.forward_host(buf.ctypes.data, buf.ctypes.data)
    assert np.array_equal(buf, want)
    plan.inverse_host(buf.ctypes.data, buf.ctypes.data)
    assert np.array_equal(buf, a)
    plan.close()


@pytest.mark.parametrize("L,batch,splits", [(20, 5, None), (16, 200, None), (22, 1, None), (22, 1, [7, 7, 8])])
def test_host_entry_point_pipelines(emu, oracle, L, batch, splits):
    """xntt_forward_host / xntt_inverse_host on buffers large enough for the chunk pipelines: a batch cut into
    (ragged) chunks of transforms, and one transform whose row pass runs in row chunks (twiddle-matrix rows offset
    per chunk).  The emulator executes streams in program order, so this checks the chunk index algebra."""
    m = 1 << L
    a = oracle.fill_xorshift(m * batch, SEED + 3 * L, P0)
    plan = emu.plan(L, batch=batch, splits=splits, inverse_factor=5)
    out, back = np.empty_like(a), np.empty_like(a)
    plan.forward_host(out.ctypes.data, a.ctypes.data)
    for b in sorted({0, batch // 2, batch - 1}):
        assert np.array_equal(out[b * m:(b + 1) * m], oracle.ntt_forward(a[b * m:(b + 1) * m].copy(), P0, G0)), b
    plan.inverse_host(back.ctypes.data, out.ctypes.data)
    scale = np.full_like(a, (m * pow(5, -1, P0)) % P0)
    assert np.array_equal(back, oracle.pointwise_mul(a, scale, P0))
    # in place
    plan.forward_host(a.ctypes.data, a.ctypes.data)
    assert np.array_equal(a, out)


@pytest.mark.parametrize("L,splits,G", [(14, None, 2), (16, [6, 5, 5], 4), (18, [7, 11], 8), (15, [3, 12], 8),
                                        (17, [8, 4, 5], 2)])
def test_mgpu_plan_in_one_process(emu, oracle, L, splits, G):
    """xntt_mgpu_*: the C++-hosted multi-GPU plan (csrc/mgpu.cpp) - G sharded sub-plans, peer-store exchange, event
    ordering - against the oracle, through the device-shard entry points and the whole-transform host entry points;
    three transforms in a row so that both exchange buffers get reused.  (The emulator has one 'device'; on a GPU
    box the same test runs with every rank on device 0, tests/test_gpu_parity.py.)"""
    m = 1 << L
    mg = emu.mgpu(L, list(range(G)), splits=splits, inverse_factor=m)
    n0, n1 = mg.n0, mg.n1
    assert n0 * n1 == m
    for rep in range(3):
        a = oracle.fill_xorshift(m, SEED + 40 + rep, P0)
        want = oracle.ntt_forward(a, P0, G0)
        A = a.reshape(n0, n1)
        blocks = [np.ascontiguousarray(A[:, r * n1 // G:(r + 1) * n1 // G]).reshape(-1) for r in range(G)]
        outs = [np.full(m // G, 0xDEAD, np.uint64) for _ in range(G)]
        mg.forward([o.ctypes.data for o in outs], [b.ctypes.data for b in blocks])
        mg.synchronize()
        assert np.array_equal(np.concatenate(outs), want), (L, splits, G, rep)
        backs = [np.full(m // G, 0xBEEF, np.uint64) for _ in range(G)]
        mg.inverse([b.ctypes.data for b in backs], [o.ctypes.data for o in outs])
        mg.synchronize()
        for r in range(G):
            assert np.array_equal(backs[r], blocks[r]), (L, splits, G, rep, r)
        # whole transform on host buffers
        got = np.empty_like(a)
        mg.forward_host(got.ctypes.data, a.ctypes.data)
        assert np.array_equal(got, want)
        back = np.empty_like(a)
        mg.inverse_host(back.ctypes.data, got.ctypes.data)
        assert np.array_equal(back, a)
    mg.close()


def test_mgpu_argument_checks(emu):
    import pytest as _pt
    for devs in ([0], [0, 1, 2], list(range(16))):
        with _pt.raises(Exception):
            emu.mgpu(16, devs)
    with _pt.raises(Exception):
        emu.mgpu(12, [0, 1])  # single-pass size: nothing to exchange


@pytest.mark.parametrize("N,g,fixed", [(0xFFFFFFFF00000001, 7, False), (0x3A00000000000001, 3, False),
                                       (0x3A00000000000001, 3, True), (0x0003F00000000001, 11, True)])
def test_mgpu_other_moduli(emu, oracle, N, g, fixed):
    """Sharded plans are not tied to the production prime: the address-mapped (exchange-side) kernels exist for the
    runtime-modulus Montgomery and Shoup flavours too."""
    for L, splits, G in [(14, None, 2), (16, [6, 5, 5], 4)]:
        m = 1 << L
        a = oracle.fill_xorshift(m, SEED + L, N)
        want = oracle.ntt_forward(a, N, g)
        mg = emu.mgpu(L, list(range(G)), splits=splits, modulus=N, generator=g, fixed_point=fixed)
        got = np.empty_like(a)
        mg.forward_host(got.ctypes.data, a.ctypes.data)
        assert np.array_equal(got, want), (hex(N), L, G)
        back = np.empty_like(a)
        mg.inverse_host(back.ctypes.data, got.ctypes.data)
        assert np.array_equal(back, a)
        mg.close()


SHOUP_MODULI = [(N, g) for N, g in OTHER_MODULI if N < (1 << 62)]


@pytest.mark.parametrize("N,g", SHOUP_MODULI)
def test_fixed_point_modmul(emu, oracle, N, g):
    """XNTT_MODMUL_FIXED_POINT: Shoup arithmetic (FixedPoint64SVE, modmul/sve/fixed-point-64.hpp:13-69) for moduli below
    2^62 - same words as the oracle, every pass kind (compact / whole-matrix twiddles, scaled inverse, batches)."""
    for L, splits, batch, kw in [(1, None, 1, {}), (3, None, 5, {}), (7, None, 1, {}), (10, None, 3, {}),
                                 (13, None, 1, {}), (15, None, 1, {}), (13, [9, 4], 1, {}),
                                 (16, [5, 5, 6], 1, {}), (14, None, 1, {"compact_tables": True}),
                                 (12, None, 2, {"inverse_factor": 12345})]:
        if (N - 1) % (1 << L):
            continue
        m = 1 << L
        a = oracle.fill_xorshift(m * batch, SEED + L, N)
        plan = emu.plan(L, modulus=N, generator=g, splits=splits, batch=batch, fixed_point=True, **kw)
        assert plan.modmul == 1
        out = np.empty_like(a)
        plan.forward(out.ctypes.data, a.ctypes.data)
        for b in range(batch):
            assert np.array_equal(out[b * m:(b + 1) * m], oracle.ntt_forward(a[b * m:(b + 1) * m].copy(), N, g)), (hex(N), L)
        back = np.empty_like(a)
        plan.inverse(back.ctypes.data, out.ctypes.data)
        f = kw.get("inverse_factor", m)
        scale = np.full_like(a, (m * pow(f, -1, N)) % N)
        assert np.array_equal(back, oracle.pointwise_mul(a, scale, N)), (hex(N), L)
        # fused point-wise product keeps its Montgomery (PAdic64) meaning on such a plan
        if L >= 3 and batch == 1:
            bm = np.empty_like(a)
            plan.to_montgomery(bm.ctypes.data, a.ctypes.data, m)
            fm = np.empty_like(a)
            plan.forward_multiply(fm.ctypes.data, a.ctypes.data, bm.ctypes.data)
            assert np.array_equal(fm, oracle.pointwise_mul(out, a, N)), (hex(N), L)


def test_fixed_point_flag_is_ignored_where_illegal(emu, oracle):
    """4p must fit 64 bits: for 62-bit-plus moduli and the production modulus the flag leaves the Montgomery kernels."""
    for N, g in [(P0, G0), (0xFFFFFFFF00000001, 7), (0xA3B25F400C7A8001, 5), (0x41D33D0D1FBF8001, 6)]:
        plan = emu.plan(10, modulus=N, generator=g, fixed_point=True)
        assert plan.modmul == 0
        a = oracle.fill_xorshift(1 << 10, SEED, N)
        out = np.empty_like(a)
        plan.forward(out.ctypes.data, a.ctypes.data)
        assert np.array_equal(out, oracle.ntt_forward(a, N, g))


@pytest.mark.parametrize("L,splits", [(14, None), (16, [5, 5, 6]), (13, [4, 4, 5]), (15, [9, 6])])
def test_lazy_residues_between_passes(emu, oracle, L, splits):
    """Column passes whose consumer begins with a Montgomery product store residues uncanonicalised (forward: every
    twist-free / handover column pass; inverse: the inner column pass of a three-pass plan, PassParams::lazy_out).
    Inputs built from p - 1, p - 2, 0 and 1 push intermediate sums to both ends of [0, 2^64); whatever crosses a pass
    boundary, a complete transform returns canonical words equal to the reference's (tests/ntt-reference.hpp:43-83)."""
    m = 1 << L
    pats = [np.full(m, P0 - 1, dtype=np.uint64),
            np.where(np.arange(m) % 2 == 0, np.uint64(P0 - 1), np.uint64(0)).astype(np.uint64),
            np.where(np.arange(m) % 3 == 0, np.uint64(P0 - 2), np.uint64(1)).astype(np.uint64),
            oracle.fill_xorshift(m, SEED + 99, P0) | np.uint64(0xFFFFFC0000000000)]
    pats[3] = np.where(pats[3] >= np.uint64(P0), np.uint64(P0 - 1), pats[3])
    for mx in ({}, {"compact_tables": True}):
        plan = emu.plan(L, splits=splits, **mx)
        for a in pats:
            out, back = np.empty_like(a), np.empty_like(a)
            plan.forward(out.ctypes.data, a.ctypes.data)
            assert int(out.max()) < P0
            assert np.array_equal(out, oracle.ntt_forward(a.copy(), P0, G0)), (L, splits, mx)
            plan.inverse(back.ctypes.data, out.ctypes.data)
            assert np.array_equal(back, a), (L, splits, mx)
        plan.close()
