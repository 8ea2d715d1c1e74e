// SPDX-License-Identifier: Apache-2.0
//
// Planner and C ABI (include/xntt.h).  Pure C++: everything device-side goes through backend.h, so
// this file is compiled unchanged into libxntt.so (CUDA back-end) and into the host emulator used
// by the CPU test-suite.  It plays the role of the reference's wrapper + kernel composition
// (include/sventt/wrapper.hpp:13-83, kernel/recursive.hpp:15-145, kernel/iterative.hpp:17-107):
// a transform of length m = 2^L is lowered to one, two or three "passes" (pass_kernel.cuh).
//
// Decomposition (forward order), m = n0 * n1 [* n2], data viewed row-major [n0][n1][n2]:
//   pass i < last : column pass of length n_i over stride n_{i+1}*..., fused with the six-step
//                   twiddle omega_M^(bitrev(k) * col), M = n_i * inner
//                   (GenericSVELayer::twiddle_rows_forward, layer/sve/generic.hpp:169-268)
//   last pass     : row pass of length n_last on contiguous rows
// The output order is the global bit reversal, identical to NTTReference.
// The inverse runs the same passes backwards with inverse tables; 1/inverse_factor is folded into
// the outermost pass's twiddle table (or applied at the end of a single-pass inverse).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/xntt.h"
#include "backend.h"
#include "field_host.h"

using namespace xntt;

struct PassDesc {
  int logn = 0;
  bool col = false;
  int log_inner = 0;  // column mode: log2 of the distance between consecutive k (unsharded)
  int log_outer = 0;  // log2 of the number of independent [N][inner] blocks per transform
  int twist_shift = 0;
  bool narrow = false;  // narrow tiles (params.h: pass_logw): plans too small to fill the GPU with whole tiles
  const Tw* fwd_tw = nullptr;
  const Tw* inv_tw = nullptr;
  const Tw* fwd_lo = nullptr;
  const Tw* fwd_hi = nullptr;
  const Tw* inv_lo = nullptr;
  const Tw* inv_hi = nullptr;
  // whole twiddle matrix [N][inner] (one Montgomery product per residue instead of two), when it fits the budget
  const Tw* fwd_full = nullptr;
  const Tw* inv_full = nullptr;
  // forward matrix of an outer pass that the NEXT column pass applies while loading (kColPre); this pass then runs
  // twist-free.  For the column-sharded first pass of a sharded plan the table is the rank's ROW block.
  bool fwd_by_next = false;
};

struct xntt_plan {
  xntt_desc desc{};
  int device = 0;
  int log2_m = 0;
  u32 batch = 1;
  bool fwd = true, inv = true;
  std::vector<PassDesc> passes;  // forward execution order; the last one is the row pass
  void* arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<void*> matrices;  // whole-matrix twiddles, one allocation each (a failed one falls back to compact)
  Tw scale{};  // Montgomery pair of inverse_factor^-1
  bool scale_on = false;
  u64 r2 = 0;  // 2^128 mod p
  u32 shard_count = 1, shard_rank = 0;
  FieldConsts field{};
  // The *_host entry points own lazily created state (staging buffer, pipeline streams and events): they are
  // serialised on this mutex, so a plan may be shared by host threads; the device entry points touch no plan state.
  mutable std::mutex host_mu;
  mutable void* staging = nullptr;  // device buffer behind the *_host entry points (lazy)
  // chunk pipeline of the host entry points of batched plans (lazy): copy-in, compute, copy-out streams + events
  mutable void* pipe_streams[3] = {nullptr, nullptr, nullptr};
  mutable std::vector<void*> pipe_events;
};

namespace {

thread_local std::string g_err = "no error";

int be_fail(int rc) {
  g_err = be::last_error();
  return rc == 2 ? XNTT_ERR_ALLOC : XNTT_ERR_CUDA;
}
#define BE(call)                       \
  do {                                 \
    int rc_ = (call);                  \
    if (rc_ != 0) return be_fail(rc_); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (be::get_device(&prev) != 0) {
      ok = false;
      prev = -1;
      return;
    }
    if (prev != dev && be::set_device(dev) != 0) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) be::set_device(prev);
  }
};

// Plans of at most this many residues (transform length x batch) are "small": whole tiles (2^13 residues) would leave
// most of the 148 SMs idle - 2^20 residues are 128 whole tiles for 296 resident CTAs - so their passes run on narrow
// tiles where a pass length has one, and the planner's own decomposition keeps both passes short enough to have one.
// Measured on B200 (tools/small_sizes.py, one forward transform, whole -> narrow tiles): 2^16 28.5 -> 13.3 us,
// 2^17 28.7 -> 14.4, 2^18 28.9 -> 16.4, 2^19 31.9 -> 21.5, 2^20 33.8 -> 30.8; at 2^21 whole tiles win (56 vs 60 us).
constexpr int kSmallPlanLog = 20;

int choose_splits(int L, std::vector<int>& out, bool small = false, u64 residues = 0) {
  out.clear();
  if (small && (L == 12 || L == 13) && residues <= (1ull << 17)) {
    // a few transforms of 2^12 / 2^13: one row pass would be one CTA of a whole tile each (16.4 / 18.5 us); two passes
    // on narrow tiles spread them over 16 CTAs per pass (7.2 / 8.6 us, still 11.2 us for 16 x 2^13; from 64 x 2^13 on the
    // single pass wins)
    out.push_back((L + 1) / 2);
    out.push_back(L / 2);
  } else if (L <= kMaxRowLog) {
    out.push_back(L);
  } else if (small && L <= 20) {
    // column pass of at most 2^9 (narrow column tiles stay four columns wide), row pass of at most 2^11
    const int l0 = L / 2 < 9 ? L / 2 : 9;
    out.push_back(l0);
    out.push_back(L - l0);
  } else if (L <= 24) {
    int l1 = L - 8 > 12 ? 12 : L - 8;  // prefer a 2^12 row pass (four clean radix-8 stages)
    if (l1 < L / 2) l1 = L / 2;
    int l0 = L - l1;
    if (l0 > 11) {  // keep column tiles at least 4 columns (32 bytes) wide
      l0 = 11;
      l1 = L - l0;
    }
    out.push_back(l0);
    out.push_back(l1);
  } else if (L <= 31) {
    const int l2 = 12, rest = L - l2;
    out.push_back((rest + 1) / 2);
    out.push_back(rest / 2);
    out.push_back(l2);
  } else {
    return XNTT_ERR_INVALID;
  }
  return XNTT_OK;
}

int gen_table(const FieldConsts& fc, Tw* out, u32 count, int kind, int logn, int shift, u64 root, u64 scale_plain,
              u32 col0 = 0, u32 row0 = 0) {
  const u64 p = fc.p;
  PowTable t{};
  t.col0 = col0;
  t.row0 = row0;
  u64 r = root;
  for (int i = 0; i < 32; ++i) {
    t.sq[i] = h_to_mont(r, p);
    r = h_mul(r, r, p);
  }
  t.scale = h_to_mont(scale_plain % p, p);
  BE(be::launch_gen_table(fc, out, count, kind, logn, shift, t, nullptr));
  return XNTT_OK;
}

// The twiddle matrix of column pass i is applied by the row pass next to it - forward while that pass loads its
// rows, inverse before it stores them (contiguous rows of the matrix next to contiguous rows of data) - and the
// column pass itself runs twist-free.
// Holds for the last column pass whenever the planner stored its forward matrix (never for the column-sharded
// first pass of a sharded plan).
int log2u(u64 v);

bool row_applies_twist(const xntt_plan* pl, size_t i, bool inverse = false) {
  // the column-sharded first pass keeps its twiddle: its (per-rank) matrix is a column block, the row half of a
  // sharded plan would need whole rows
  if (pl->shard_count > 1 && i == 0) return false;
  // Inverse of a plan sharded over 4 or more GPUs: column pass i is the one that stores its output into the peers'
  // buffers and is bound by the links, not by arithmetic - the twiddle product is free there and would only lengthen
  // the row pass (2^30 over 8 GPUs: inverse 4.38 vs 4.45 ms; over 2 GPUs the row pass wins, 2^28 3.91 vs 3.97 ms).
  if (inverse && pl->shard_count >= 4) return false;
  return i + 2 == pl->passes.size() && (inverse ? pl->passes[i].inv_full : pl->passes[i].fwd_full) != nullptr;
}

// Forward column pass i of a plan whose outer passes hand their twiddle matrix to the next pass (PassDesc::fwd_by_next):
// the outer pass runs twist-free, the pass behind it multiplies by that matrix while loading (kColPre).
void set_forward_handover(const xntt_plan* pl, size_t i, PassParams& prm) {
  const PassDesc& ps = pl->passes[i];
  if (ps.fwd_by_next) prm.twist_lo = prm.twist_hi = prm.twist_full = nullptr;
  if (i > 0 && pl->passes[i - 1].fwd_by_next) {
    const PassDesc& pp = pl->passes[i - 1];
    prm.pre_twist = pp.fwd_full;
    // rows of the stored table: all of them, or the rank's row block behind the exchange of a sharded plan
    const u32 rows = (1u << pp.logn) / ((i - 1 == 0 && pl->shard_count > 1) ? pl->shard_count : 1u);
    prm.pre_rows_mask = rows - 1u;
    prm.pre_shift = (u32)pp.log_inner;
    prm.pre_kshift = (u32)ps.log_inner;
  }
}

// Inverse column pass i > 0: the pass that consumes its output (column pass i - 1, executed next) applies its own
// six-step twiddle while loading, i.e. begins with a Montgomery product - which takes any 64-bit value - so pass i may
// skip canonicalising what it stores (PassParams::lazy_out).
bool inverse_output_may_stay_lazy(const xntt_plan* pl, size_t i) {
  if (i == 0 || !pl->passes[i].col) return false;
  const PassDesc& c = pl->passes[i - 1];
  if (!c.col || row_applies_twist(pl, i - 1, true)) return false;
  return c.inv_full != nullptr || (c.inv_lo != nullptr && c.inv_hi != nullptr);
}

// One pass.  `count_override` (when non-zero) replaces the number of outer blocks / rows: the row
// half of a sharded plan only holds 1/shard_count of them.
int run_pass(const xntt_plan* pl, size_t i, bool inverse, u64* dst, const u64* src, void* st,
             u64 count_override, const u64* pointwise = nullptr, u64 chunk_row0 = 0) {
  const PassDesc& ps = pl->passes[i];
  PassParams prm{};
  prm.src = src;
  prm.dst = dst;
  prm.tw = inverse ? ps.inv_tw : ps.fwd_tw;
  prm.scale = pl->scale;
  prm.field = pl->field;
  const int logw = pass_logw(ps.logn, ps.narrow);
  prm.narrow = ps.narrow ? 1u : 0u;
  unsigned grid;
  if (ps.col) {
    const bool sharded_first = (i == 0 && pl->shard_count > 1);
    const u64 inner_full = 1ull << ps.log_inner;
    const u64 inner = sharded_first ? inner_full / pl->shard_count : inner_full;
    const u64 outer = count_override ? count_override : ((u64)pl->batch << ps.log_outer);
    prm.inner = inner;
    prm.outer_stride = inner << ps.logn;
    prm.tiles_per_outer = (u32)(inner >> logw);
    prm.twist_lo = inverse ? ps.inv_lo : ps.fwd_lo;
    prm.twist_hi = inverse ? ps.inv_hi : ps.fwd_hi;
    prm.twist_shift = (u32)ps.twist_shift;
    prm.twist_full = inverse ? ps.inv_full : ps.fwd_full;
    prm.twist_full_shift = (u32)log2u(inner);  // columns per row of the stored matrix (a rank's block if sharded)
    if (row_applies_twist(pl, i, inverse)) prm.twist_lo = prm.twist_hi = prm.twist_full = nullptr;
    if (!inverse) set_forward_handover(pl, i, prm);
    prm.lazy_out = (inverse && inverse_output_may_stay_lazy(pl, i)) ? 1u : 0u;
    prm.twist_col0 = sharded_first ? (u32)(inner * pl->shard_rank) : 0u;
    const u64 tiles = outer * prm.tiles_per_outer;
    if (tiles == 0 || tiles > 0x7fffffffull) return XNTT_ERR_INVALID;
    grid = (unsigned)tiles;
  } else {
    const u64 rows = count_override ? count_override : ((u64)pl->batch << (pl->log2_m - ps.logn));
    if (rows == 0 || rows > 0xffffffffull) return XNTT_ERR_INVALID;
    prm.rows = (u32)rows;
    prm.scale_on = (inverse && pl->scale_on && pl->passes.size() == 1) ? 1u : 0u;
    prm.pointwise = inverse ? nullptr : pointwise;
    if (i > 0 && row_applies_twist(pl, i - 1, inverse)) {
      prm.pre_twist = inverse ? pl->passes[i - 1].inv_full : pl->passes[i - 1].fwd_full;
      prm.pre_rows_mask = (1u << pl->passes[i - 1].logn) - 1u;
      // a row chunk (host pipeline): power-of-two chunks are either whole multiples of the matrix height or lie
      // inside one copy of it, so the rows before the chunk turn into a pointer offset
      prm.pre_twist += (chunk_row0 & prm.pre_rows_mask) << ps.logn;
    }
    grid = (unsigned)((rows + (1u << logw) - 1) >> logw);
  }
  BE(be::launch_pass(ps.logn, ps.col, inverse, false, prm, grid, st));
  return XNTT_OK;
}

// passes [first, last) in forward order (reverse order for the inverse); the first executed pass
// reads src, every later one works in place on dst.
int run_range(const xntt_plan* pl, bool inverse, size_t first, size_t last, u64* dst, const u64* src, void* st,
              bool shard_rows, const u64* pointwise = nullptr, u32 batch_override = 0) {
  if (!dst || !src) return XNTT_ERR_INVALID;
  if (inverse ? !pl->inv : !pl->fwd) return XNTT_ERR_STATE;
  DeviceGuard g(pl->device);
  if (!g.ok) return be_fail(1);
  const u64* cur = src;
  for (size_t s = 0; s < last - first; ++s) {
    const size_t i = inverse ? last - 1 - s : first + s;
    u64 count = 0;
    if (batch_override) {  // a chunk of the plan's batch (host pipeline)
      const PassDesc& ps = pl->passes[i];
      count = ps.col ? ((u64)batch_override << ps.log_outer) : ((u64)batch_override << (pl->log2_m - ps.logn));
    } else if (shard_rows) {
      const PassDesc& ps = pl->passes[i];
      const u64 full = ps.col ? ((u64)pl->batch << ps.log_outer) : ((u64)pl->batch << (pl->log2_m - ps.logn));
      count = full / pl->shard_count;
    }
    const int rc = run_pass(pl, i, inverse, dst, cur, st, count, pointwise);
    if (rc != XNTT_OK) return rc;
    cur = dst;
  }
  return XNTT_OK;
}

// Batched plans: the transforms are independent, so the batch is cut into chunks that flow through three streams -
// copy in, transform, copy out - and the two DMA directions of the link run at the same time.  (One transform
// cannot do that: its last pass needs every word of its first.)
int host_pipeline(const xntt_plan* pl, u64* d, uint64_t* dst, const uint64_t* src, bool inverse, u32 chunks) {
  const size_t words = (size_t)1 << pl->log2_m;
  const u32 per = (pl->batch + chunks - 1) / chunks;
  if (!pl->pipe_streams[0])
    for (int i = 0; i < 3; ++i) BE(be::stream_create(&pl->pipe_streams[i]));
  while (pl->pipe_events.size() < 2 * (size_t)chunks) {
    void* ev = nullptr;
    BE(be::event_create(&ev));
    pl->pipe_events.push_back(ev);
  }
  void *s_in = pl->pipe_streams[0], *s_run = pl->pipe_streams[1], *s_out = pl->pipe_streams[2];
  const size_t q = pl->passes.size();
  for (u32 i = 0, b0 = 0; b0 < pl->batch; ++i, b0 += per) {
    const u32 nb = pl->batch - b0 < per ? pl->batch - b0 : per;
    const size_t off = (size_t)b0 * words, bytes = (size_t)nb * words * sizeof(u64);
    BE(be::memcpy_h2d(d + off, src + off, bytes, s_in));
    BE(be::event_record(pl->pipe_events[2 * i], s_in));
    BE(be::stream_wait_event(s_run, pl->pipe_events[2 * i]));
    const int rc = run_range(pl, inverse, 0, q, d + off, d + off, s_run, false, nullptr, nb);
    if (rc != XNTT_OK) return rc;
    BE(be::event_record(pl->pipe_events[2 * i + 1], s_run));
    BE(be::stream_wait_event(s_out, pl->pipe_events[2 * i + 1]));
    BE(be::memcpy_d2h(dst + off, d + off, bytes, s_out));
  }
  BE(be::stream_sync(s_out));
  BE(be::stream_sync(s_run));
  return XNTT_OK;
}

// One large transform (or a batch too small to cut): only the row pass - contiguous rows - can overlap a copy.
// (Measured and dropped: running the outermost column pass column block by column block under strided copies in /
// out hides its 0.16 ms at 2^24, but the strided copies cost more than that - forward 5.05 -> 5.13 ms, inverse 5.08 -> 5.43 ms.)
// Forward: copy in, column passes, then the row pass in row chunks, each followed by its copy out; inverse: the
// row pass of a chunk as soon as its rows have arrived, then the column passes and one copy out.
int host_row_pipeline(const xntt_plan* pl, u64* d, uint64_t* dst, const uint64_t* src, bool inverse, u32 chunks) {
  const size_t q = pl->passes.size(), last = q - 1;
  const PassDesc& rp = pl->passes[last];
  const u64 rows = (u64)pl->batch << (pl->log2_m - rp.logn);
  while (chunks > 1 && (rows % chunks != 0 || (rows / chunks) % (1u << pass_logw(rp.logn, rp.narrow)) != 0)) chunks /= 2;
  const u64 per = rows / chunks;
  const size_t chunk_words = (size_t)per << rp.logn, chunk_bytes = chunk_words * sizeof(u64);
  const size_t bytes = (sizeof(u64) << pl->log2_m) * pl->batch;
  if (!pl->pipe_streams[0])
    for (int i = 0; i < 3; ++i) BE(be::stream_create(&pl->pipe_streams[i]));
  while (pl->pipe_events.size() < 2 * (size_t)chunks) {
    void* ev = nullptr;
    BE(be::event_create(&ev));
    pl->pipe_events.push_back(ev);
  }
  void *s_in = pl->pipe_streams[0], *s_run = pl->pipe_streams[1], *s_out = pl->pipe_streams[2];
  int rc;
  if (!inverse) {
    BE(be::memcpy_h2d(d, src, bytes, s_run));
    if (q > 1 && (rc = run_range(pl, false, 0, last, d, d, s_run, false)) != XNTT_OK) return rc;
    for (u32 i = 0; i < chunks; ++i) {
      const size_t off = (size_t)i * chunk_words;
      if ((rc = run_pass(pl, last, false, d + off, d + off, s_run, per, nullptr, (u64)i * per)) != XNTT_OK) return rc;
      BE(be::event_record(pl->pipe_events[i], s_run));
      BE(be::stream_wait_event(s_out, pl->pipe_events[i]));
      BE(be::memcpy_d2h(dst + off, d + off, chunk_bytes, s_out));
    }
    BE(be::stream_sync(s_out));
  } else {
    for (u32 i = 0; i < chunks; ++i) {
      const size_t off = (size_t)i * chunk_words;
      BE(be::memcpy_h2d(d + off, src + off, chunk_bytes, s_in));
      BE(be::event_record(pl->pipe_events[i], s_in));
      BE(be::stream_wait_event(s_run, pl->pipe_events[i]));
      if ((rc = run_pass(pl, last, true, d + off, d + off, s_run, per, nullptr, (u64)i * per)) != XNTT_OK) return rc;
    }
    if (q > 1 && (rc = run_range(pl, true, 0, last, d, d, s_run, false)) != XNTT_OK) return rc;
    BE(be::memcpy_d2h(dst, d, bytes, s_run));
  }
  BE(be::stream_sync(s_run));
  return XNTT_OK;
}

// Largest single-pass plan (bytes per buffer) that works on the caller's page-locked host buffers directly.  Measured on
// B200 (tools/host_small.py, blocking call on pinned buffers, staged -> direct): 2^10 36.7 -> 19.5 us.  Letting only the
// row pass of a longer plan touch the host buffer gains nothing at 2^13 (33.8 -> 33.4 us) and from 2^15 on SM-issued
// PCIe traffic loses against the copy engines (2^17 forward 117 -> 132 us, 2^20 381 -> 844 us): those stay staged.
#ifndef XNTT_ZERO_COPY_MAX_KB
#define XNTT_ZERO_COPY_MAX_KB 64
#endif
constexpr size_t kZeroCopyMaxBytes = (size_t)XNTT_ZERO_COPY_MAX_KB << 10;

int host_roundtrip(const xntt_plan* pl, uint64_t* dst, const uint64_t* src, bool inverse) {
  if (!pl || !dst || !src) return XNTT_ERR_INVALID;
  if (pl->shard_count > 1) return XNTT_ERR_STATE;
  if (inverse ? !pl->inv : !pl->fwd) return XNTT_ERR_STATE;
  DeviceGuard g(pl->device);
  if (!g.ok) return be_fail(1);
  std::lock_guard<std::mutex> host_lock(pl->host_mu);
  const size_t bytes = (sizeof(u64) << pl->log2_m) * pl->batch;
  if (!pl->staging) BE(be::dev_malloc(&pl->staging, bytes));
  void* d = pl->staging;
  // chunks of at least 8 MiB (below that the copies are latency-bound), at most 8 of them
  u32 chunks = 1;
  while (chunks < 8 && chunks * 2 <= pl->batch && bytes / (chunks * 2) >= ((size_t)8 << 20)) chunks *= 2;
  if (chunks > 1) {
    const int prc = host_pipeline(pl, (u64*)d, dst, src, inverse, chunks);
    if (prc != XNTT_OK) be::stream_sync(nullptr);
    return prc;
  }
  if (bytes >= ((size_t)32 << 20)) {
    const int prc = host_row_pipeline(pl, (u64*)d, dst, src, inverse, 8);
    if (prc != XNTT_OK) be::stream_sync(nullptr);
    return prc;
  }
  int rc = XNTT_OK, brc;
  // Tiny single-pass plans on page-locked (mapped) host buffers: the one kernel reads and writes the host buffers
  // themselves over PCIe - no staging, no copy launches.
  void *src_dev = nullptr, *dst_dev = nullptr;
  if (bytes <= kZeroCopyMaxBytes && pl->passes.size() == 1) {
    be::host_device_pointer(src, &src_dev);
    be::host_device_pointer(dst, &dst_dev);
  }
  if (src_dev && dst_dev) {
    rc = run_pass(pl, 0, inverse, (u64*)dst_dev, (const u64*)src_dev, nullptr, 0);
  } else {
    if ((brc = be::memcpy_h2d(d, src, bytes, nullptr)) != 0) rc = be_fail(brc);
    if (rc == XNTT_OK) rc = run_range(pl, inverse, 0, pl->passes.size(), (u64*)d, (const u64*)d, nullptr, false);
    if (rc == XNTT_OK && (brc = be::memcpy_d2h(dst, d, bytes, nullptr)) != 0) rc = be_fail(brc);
  }
  brc = be::stream_sync(nullptr);
  if (rc == XNTT_OK && brc != 0) rc = be_fail(brc);
  return rc;
}

// ---- sharded plans: exchange-friendly ("tiled") variants -----------------------------------------
// Notation: G ranks, K chunks, m = n0 * n1, w = n1 / (G * K) columns per chunk.
//   chunk layout  : chunk c of a rank's column block, compact, [n0][w] row-major = G contiguous
//                   messages of (n0/G) * w words (message s = rows of rank s)
//   tiled layout  : what a rank holds after the all-to-all of every chunk: [K][G][n0/G][w]
// The column pass writes (reads) chunk layout; the first (last) local pass of the row half reads
// (writes) the tiled layout directly through a StrideMap, so no separate (un)packing pass exists.
StrideMap natural_map(const xntt_plan* pl, const PassDesc& ps) {
  StrideMap m{};
  m.b1 = m.b2 = 31;
  if (ps.col) {
    m.s0 = 1ull << ps.log_inner;
    m.outer = 1ull << (ps.log_inner + ps.logn);
  } else {
    m.s0 = 1;
    m.outer = 1ull << ps.logn;
  }
  (void)pl;
  return m;
}

int log2u(u64 v) {
  int l = 0;
  while ((1ull << l) < v) ++l;
  return l;
}

// map of pass 1's transform index onto the tiled layout
int tiled_map(const xntt_plan* pl, u32 nchunks, StrideMap& out) {
  const PassDesc& p1 = pl->passes[1];
  const u64 G = pl->shard_count, K = nchunks;
  const u64 n0 = 1ull << pl->passes[0].logn, n1 = (1ull << pl->log2_m) / n0;
  if (K == 0 || (K & (K - 1)) || n1 % (G * K)) return XNTT_ERR_INVALID;
  const u64 w = n1 / (G * K);
  StrideMap m{};
  if (p1.col) {
    const u64 n1b = 1ull << p1.log_inner;  // contiguous inner run of pass 1
    if (w % n1b) return XNTT_ERR_UNSUPPORTED;
    m.s0 = n1b;
    m.b1 = (u32)log2u(w / n1b);
  } else {
    m.s0 = 1;
    m.b1 = (u32)log2u(w);
  }
  m.b2 = m.b1 + (u32)log2u(K);
  m.s1 = n0 * w;          // next chunk
  m.s2 = (n0 / G) * w;    // next source rank inside a chunk
  m.outer = w;            // next local row
  out = m;
  return XNTT_OK;
}

int run_pass_mapped(const xntt_plan* pl, size_t i, bool inverse, u64* dst, const u64* src, const StrideMap& smap,
                    const StrideMap& dmap, u64 units, u32 tiles_per_outer, u32 twist_col0, void* st,
                    u64* const* peers = nullptr, u32 peer_bits = 0, u64 peer_offset = 0) {
  const PassDesc& ps = pl->passes[i];
  PassParams prm{};
  prm.src = src;
  prm.dst = dst;
  prm.tw = inverse ? ps.inv_tw : ps.fwd_tw;
  prm.scale = pl->scale;
  prm.field = pl->field;
  prm.smap = smap;
  prm.dmap = dmap;
  if (peers) {
    for (u32 s = 0; s < pl->shard_count; ++s) prm.peer[s] = peers[s] + peer_offset;
    prm.peer_bits = peer_bits;
    prm.peer_on = 1;
    prm.dst = peers[pl->shard_rank] + peer_offset;  // tile offsets are taken relative to prm.dst
  }
  const int logw = tile_logw(ps.logn);
  unsigned grid;
  if (ps.col) {
    prm.tiles_per_outer = tiles_per_outer;
    prm.twist_lo = inverse ? ps.inv_lo : ps.fwd_lo;
    prm.twist_hi = inverse ? ps.inv_hi : ps.fwd_hi;
    prm.twist_shift = (u32)ps.twist_shift;
    prm.twist_full = inverse ? ps.inv_full : nullptr;  // forward: compact, or none when the row pass applies it
    prm.twist_full_shift = (u32)ps.log_inner - ((i == 0 && pl->shard_count > 1) ? (u32)log2u(pl->shard_count) : 0u);
    if (row_applies_twist(pl, i, inverse)) prm.twist_lo = prm.twist_hi = prm.twist_full = nullptr;
    if (!inverse) set_forward_handover(pl, i, prm);
    prm.lazy_out = (inverse && inverse_output_may_stay_lazy(pl, i)) ? 1u : 0u;
    prm.twist_col0 = twist_col0;
    const u64 tiles = units * tiles_per_outer;
    if (tiles == 0 || tiles > 0x7fffffffull) return XNTT_ERR_INVALID;
    grid = (unsigned)tiles;
  } else {
    prm.rows = (u32)units;
    prm.scale_on = 0;
    grid = (unsigned)((units + (1u << logw) - 1) >> logw);
  }
  BE(be::launch_pass(ps.logn, ps.col, inverse, true, prm, grid, st));
  return XNTT_OK;
}

int shard_cols_chunk(const xntt_plan* pl, bool inverse, u64* dst, const u64* src, u32 chunk, u32 nchunks, void* st) {
  if (!pl || !dst || !src || pl->shard_count < 2) return XNTT_ERR_STATE;
  if (inverse ? !pl->inv : !pl->fwd) return XNTT_ERR_STATE;
  if (pl->batch != 1 || nchunks == 0 || (nchunks & (nchunks - 1)) || chunk >= nchunks) return XNTT_ERR_INVALID;
  const PassDesc& p0 = pl->passes[0];
  const u64 G = pl->shard_count, n0 = 1ull << p0.logn, n1 = (1ull << pl->log2_m) / n0;
  if (n1 % (G * nchunks)) return XNTT_ERR_INVALID;
  const u64 w = n1 / (G * nchunks), block = n1 / G;
  if (w < (1ull << tile_logw(p0.logn))) return XNTT_ERR_UNSUPPORTED;
  DeviceGuard g(pl->device);
  if (!g.ok) return be_fail(1);
  StrideMap full{}, compact{};
  full.b1 = full.b2 = compact.b1 = compact.b2 = 31;
  full.s0 = block;
  full.outer = n0 * block;
  compact.s0 = w;
  compact.outer = n0 * w;
  const u32 tpo = (u32)(w >> tile_logw(p0.logn));
  const u32 col0 = (u32)(block * pl->shard_rank + w * chunk);
  if (!inverse)  // block -> chunk layout
    return run_pass_mapped(pl, 0, false, dst + n0 * w * chunk, src + w * chunk, full, compact, 1, tpo, col0, st);
  return run_pass_mapped(pl, 0, true, dst + w * chunk, src + n0 * w * chunk, compact, full, 1, tpo, col0, st);
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int xntt_plan_create(xntt_plan** out, const xntt_desc* d) {
  if (!out || !d) return XNTT_ERR_INVALID;
  *out = nullptr;
  if (d->log2_m < 1 || d->log2_m > 31) return XNTT_ERR_INVALID;
  const u64 p = d->modulus, gen = d->generator;
  if ((p & 1) == 0 || p < 3 || gen == 0) return XNTT_ERR_INVALID;
  if (!h_is_prime(p)) return XNTT_ERR_INVALID;  // Modulus<> assumes a prime (modulus.hpp:13)
  const u64 m = 1ull << d->log2_m;
  if ((p - 1) % m != 0) return XNTT_ERR_INVALID;  // Modulus::get_root_forward: invalid_argument
  const u64 root_m = h_pow(gen % p, (p - 1) / m, p);
  // g must generate the order-m subgroup: omega^(m/2) == -1
  if (h_pow(root_m, m / 2, p) != p - 1) return XNTT_ERR_INVALID;

  // narrow tiles: small unsharded plans of a field that has the kernels, unless a flag says otherwise
  bool narrow_tiles = false;
  {
    FieldConsts fc{};
    fc.p = p;
    fc.kind = ((d->flags & XNTT_MODMUL_FIXED_POINT) && p < (1ull << 62) && p != kP0) ? kFieldShoup : kFieldMontgomery;
    const u64 residues = (u64)(d->batch ? d->batch : 1) << d->log2_m;
    narrow_tiles = d->shard_count <= 1 && field_has_narrow(fc) && !(d->flags & XNTT_TILES_WIDE) &&
                   ((d->flags & XNTT_TILES_NARROW) || residues <= (1ull << kSmallPlanLog));
  }
  std::vector<int> splits;
  if (d->n_splits) {
    if (d->n_splits > XNTT_MAX_SPLITS) return XNTT_ERR_INVALID;
    u32 sum = 0;
    for (u32 i = 0; i < d->n_splits; ++i) {
      splits.push_back((int)d->split_log2[i]);
      sum += d->split_log2[i];
    }
    // IterativeNTT / RecursiveNTT static_assert: the product of the radices equals m
    if (sum != d->log2_m) return XNTT_ERR_INVALID;
  } else {
    const int rc = choose_splits((int)d->log2_m, splits, narrow_tiles, (u64)(d->batch ? d->batch : 1) << d->log2_m);
    if (rc != XNTT_OK) return rc;
  }
  const u32 shard_count = d->shard_count ? d->shard_count : 1;
  if (shard_count & (shard_count - 1)) return XNTT_ERR_INVALID;
  int shard_log = 0;
  while ((1u << shard_log) < shard_count) ++shard_log;
  if (shard_count > 1) {
    if (splits.size() < 2 || d->shard_rank >= shard_count || splits[0] < shard_log) return XNTT_ERR_INVALID;
  }
  {
    int rem = (int)d->log2_m;
    for (size_t i = 0; i < splits.size(); ++i) {
      const int l = splits[i];
      rem -= l;
      const bool last = i + 1 == splits.size();
      if (l < 1) return XNTT_ERR_INVALID;
      if (l > (last ? kMaxRowLog : kMaxColLog)) return XNTT_ERR_UNSUPPORTED;
      if (!last) {
        const int log_inner = rem - (i == 0 ? shard_log : 0);
        if (log_inner < tile_logw(l)) return XNTT_ERR_UNSUPPORTED;  // tile wider than the matrix
      }
    }
  }

  // inverse scaling factor
  u64 f = d->inverse_factor == 0 ? 1 : d->inverse_factor % p;
  if (f == 0) return XNTT_ERR_INVALID;

  xntt_plan* pl = new (std::nothrow) xntt_plan;
  if (!pl) return XNTT_ERR_ALLOC;
  pl->desc = *d;
  pl->log2_m = (int)d->log2_m;
  pl->batch = d->batch ? d->batch : 1;
  const u32 dirs = d->flags & (XNTT_ENABLE_FORWARD | XNTT_ENABLE_INVERSE);
  const u32 flags = dirs ? dirs : (XNTT_ENABLE_FORWARD | XNTT_ENABLE_INVERSE);
  pl->fwd = (flags & XNTT_ENABLE_FORWARD) != 0;
  pl->inv = (flags & XNTT_ENABLE_INVERSE) != 0;
  pl->shard_count = shard_count;
  pl->shard_rank = d->shard_rank;
  const u64 finv = h_inv(f, p);
  pl->scale_on = (f != 1);
  pl->r2 = h_to_mont(h_to_mont(1, p), p);
  pl->field.p = p;
  pl->field.pinv = h_montgomery_inverse(p);
  pl->field.one = h_to_mont(1, p);
  // FixedPoint64 (Shoup) arithmetic where it is legal: 4p < 2^64; the production modulus has its own kernels
  pl->field.kind = ((d->flags & XNTT_MODMUL_FIXED_POINT) && p != kP0 && (p >> 62) == 0) ? kFieldShoup : kFieldMontgomery;
  if (pl->field.kind == kFieldShoup) {
    pl->scale.w = finv;  // (omega, floor(omega * 2^64 / p))
    pl->scale.wp = (u64)(((u128)finv << 64) / p);
  } else {
    pl->scale.w = h_to_mont(finv, p);
    pl->scale.wp = pl->scale.w * h_montgomery_inverse(p);
  }

  int dev = d->device;
  if (dev < 0) {
    const int rc = be::get_device(&dev);
    if (rc != 0) {
      delete pl;
      return be_fail(rc);
    }
  }
  pl->device = dev;
  DeviceGuard g(dev);
  if (!g.ok) {
    delete pl;
    return be_fail(1);
  }

  // arena layout (Tw units)
  const size_t q = splits.size();
  pl->passes.resize(q);
  size_t words = 0;
  std::vector<size_t> off_fwd(q), off_inv(q), off_flo(q), off_fhi(q), off_ilo(q), off_ihi(q);
  std::vector<char> use_ffull(q, 0), use_ifull(q, 0);
  // Whole-matrix twiddles (N * inner entries of 16 bytes per pass and direction: one modular product per residue
  // instead of two, no random table look-ups) while they fit the budget, outermost pass first, inverse before
  // forward.  The matrices of the last column pass are applied by the row pass next to it (row_applies_twist); an
  // outer column pass of a three-pass plan consumes its own, on load in the inverse and at the end of its tiles in
  // the forward direction - the latter only pays off while the matrix stays L2-resident (<= 64 MiB).
  // Default budget 512 MiB per plan (both directions up to 2^24 cells), plus the outermost matrix of a three-pass plan
  // while memory is plentiful; xntt_desc::twist_table_max_mb overrides;
  // XNTT_COMPACT_TABLES and the column-sharded first pass of a sharded plan keep the compact two-table form.
  size_t full_budget = (size_t)(d->twist_table_max_mb ? d->twist_table_max_mb : 512u) << 20;
  if (!d->twist_table_max_mb && shard_count > 1) {
    // sharded plans: the rank's column block of the first pass's matrix (16 m / G bytes) for the inverse, which applies
    // it on load - one modular product per residue instead of two in the pass that feeds the exchange
    const size_t big = (sizeof(Tw) << pl->log2_m) / shard_count * (q >= 3 && p == kP0 ? 2 : 1);  // + forward (kColPre)
    size_t free_b = 0, total_b = 0;
    if (be::mem_info(&free_b, &total_b) == 0 && big <= free_b / 4) full_budget += big;
  }
  if (!d->twist_table_max_mb && q >= 3 && shard_count == 1) {
    // three-pass plans (2^25 and above): the outermost pass has an m-entry matrix of its own (16 m bytes, twice the
    // data of one transform).  The inverse applies it on load and gains 7 % (2^30: 32.0 -> 29.8 ms); by default it is
    // stored when it takes no more than a quarter of the memory that is free right now.
    // (the forward hands its copy to the next pass, kColPre: production modulus only)
    const size_t big = (sizeof(Tw) << pl->log2_m) * (p == kP0 ? 2 : 1);
    size_t free_b = 0, total_b = 0;
    if (be::mem_info(&free_b, &total_b) == 0 && big <= free_b / 4) full_budget += big;
  }
  if (d->flags & XNTT_COMPACT_TABLES) full_budget = 0;
  {
    int rem = pl->log2_m, before = 0;
    for (size_t i = 0; i < q; ++i) {
      PassDesc& ps = pl->passes[i];
      ps.logn = splits[i];
      rem -= ps.logn;
      ps.col = i + 1 < q;
      ps.log_inner = rem;
      ps.log_outer = before;
      ps.narrow = narrow_tiles && has_narrow_tile(ps.logn, ps.col);
      before += ps.logn;
      const size_t n = 1ull << ps.logn;
      off_fwd[i] = words;
      words += n / 2 ? n / 2 : 1;
      off_inv[i] = words;
      words += n;
      if (ps.col) {
        const int lm = ps.logn + ps.log_inner;
        ps.twist_shift = (lm + 1) / 2;
        const size_t nlo = 1ull << ps.twist_shift, nhi = 1ull << (lm - ps.twist_shift);
        off_flo[i] = words;
        words += nlo;
        off_fhi[i] = words;
        words += nhi;
        off_ilo[i] = words;
        words += nlo;
        off_ihi[i] = words;
        words += nhi;
      }
    }
    // Which matrices get stored.  First the last column pass (both directions: the row pass next to it applies them),
    // then the outer passes from the outermost in: inverse = applied by the pass itself on load (a column-sharded
    // first pass stores its rank's column block); forward = applied by the pass itself at the end of its tiles while
    // the matrix stays L2-resident (<= 64 MiB), else handed to the NEXT column pass (production modulus; that pass has
    // to be twist-free itself, i.e. the last column pass with its own matrix stored; a sharded first pass stores its
    // rank's row block, which is what the pass behind the exchange sees).
    auto cells_of = [&](size_t i, bool block) {
      const int lm = pl->passes[i].logn + pl->passes[i].log_inner;
      return ((size_t)1 << lm) >> (block ? shard_log : 0);
    };
    if (q >= 2) {
      const size_t i = q - 2;
      const bool sharded_first = shard_count > 1 && i == 0;
      const int lm = pl->passes[i].logn + pl->passes[i].log_inner;
      if (lm <= 31) {
        const size_t bytes = cells_of(i, sharded_first) * sizeof(Tw);
        if (pl->inv && bytes <= full_budget) full_budget -= bytes, use_ifull[i] = 1;
        if (!sharded_first && pl->fwd && bytes <= full_budget) full_budget -= bytes, use_ffull[i] = 1;
      }
    }
    for (size_t i = 0; i + 2 < q; ++i) {
      PassDesc& ps = pl->passes[i];
      const bool sharded_first = shard_count > 1 && i == 0;
      const int lm = ps.logn + ps.log_inner;
      if (lm > 31) continue;
      const size_t bytes = cells_of(i, sharded_first) * sizeof(Tw);
      if (pl->inv && bytes <= full_budget) full_budget -= bytes, use_ifull[i] = 1;
      if (!pl->fwd || bytes > full_budget) continue;
      if (!sharded_first && bytes <= ((size_t)64 << 20)) {
        full_budget -= bytes, use_ffull[i] = 1;
      } else if (p == kP0 && i + 3 == q && use_ffull[i + 1] && pl->passes[i + 1].logn >= 8 &&
                 !(sharded_first && shard_count > 4)) {
        // (sharded over 8 GPUs the first pass is bound by the links, its twiddle products are free there and the pass
        // behind the exchange would only get longer: 2^30 forward 4.07 -> 4.23 ms; over 2 and 4 GPUs the handover wins,
        // 16.04 -> 15.50 ms and 8.05 -> 7.96 ms)
        // (measured: with a pass of 2^7 behind it the handover loses - that pass is short enough to feel the second
        // 16 B/residue stream: 2^26 forward 1.80 -> 1.96 ms - from 2^8 on it wins: 2^28 7.42 -> 7.13, 2^30 30.8 -> 29.8 ms)
        full_budget -= bytes, use_ffull[i] = 1;
        ps.fwd_by_next = true;
      }
    }
  }
  pl->arena_bytes = words * sizeof(Tw);
  {
    const int rc = be::dev_malloc(&pl->arena, pl->arena_bytes);
    if (rc != 0) {
      delete pl;
      g_err = be::last_error();
      return XNTT_ERR_ALLOC;
    }
  }
  Tw* base = static_cast<Tw*>(pl->arena);
  int rc = XNTT_OK;
  for (size_t i = 0; i < q && rc == XNTT_OK; ++i) {
    PassDesc& ps = pl->passes[i];
    const u64 n = 1ull << ps.logn;
    const u64 root_n = h_pow(gen % p, (p - 1) / n, p);
    ps.fwd_tw = base + off_fwd[i];
    ps.inv_tw = base + off_inv[i];
    rc = gen_table(pl->field, base + off_fwd[i], (u32)(n / 2 ? n / 2 : 1), kFwdG, ps.logn, 0, root_n, 1);
    if (rc == XNTT_OK) rc = gen_table(pl->field, base + off_inv[i], (u32)n, kInvI, ps.logn, 0, h_inv(root_n, p), 1);
    if (ps.col && rc == XNTT_OK) {
      const int lm = ps.logn + ps.log_inner;
      const u64 root_big = h_pow(gen % p, (p - 1) >> lm, p);
      const u64 root_big_inv = h_inv(root_big, p);
      const u32 nlo = 1u << ps.twist_shift, nhi = 1u << (lm - ps.twist_shift);
      ps.fwd_lo = base + off_flo[i];
      ps.fwd_hi = base + off_fhi[i];
      ps.inv_lo = base + off_ilo[i];
      ps.inv_hi = base + off_ihi[i];
      rc = gen_table(pl->field, base + off_flo[i], nlo, kPowers, 0, 0, root_big, 1);
      if (rc == XNTT_OK) rc = gen_table(pl->field, base + off_fhi[i], nhi, kPowers, 0, ps.twist_shift, root_big, 1);
      if (rc == XNTT_OK) rc = gen_table(pl->field, base + off_ilo[i], nlo, kPowers, 0, 0, root_big_inv, 1);
      // the outermost column pass runs last in the inverse: fold 1/inverse_factor into its table
      if (rc == XNTT_OK)
        rc = gen_table(pl->field, base + off_ihi[i], nhi, kPowers, 0, ps.twist_shift, root_big_inv, i == 0 ? finv : 1);
      // the matrices get allocations of their own: if the device cannot spare one, that pass and direction simply
      // keep the compact form
      const bool sharded_first = shard_count > 1 && i == 0;
      for (int dir = 0; dir < 2 && rc == XNTT_OK; ++dir) {
        if (!(dir ? use_ifull[i] : use_ffull[i])) continue;
        // whole matrix; for a column-sharded first pass the rank's column block (inverse: this pass applies it) or its
        // row block (forward: the pass behind the exchange applies it)
        const bool row_block = sharded_first && !dir;
        const int log_cols = ps.log_inner - ((sharded_first && dir) ? shard_log : 0);
        const u32 cells = (1u << (ps.logn + ps.log_inner)) >> (sharded_first ? shard_log : 0);
        const u32 col0 = (sharded_first && dir) ? (u32)(((u64)1 << log_cols) * pl->shard_rank) : 0u;
        const u32 row0 = row_block ? (u32)(((1u << ps.logn) / shard_count) * pl->shard_rank) : 0u;
        void* mem = nullptr;
        if (be::dev_malloc(&mem, (size_t)cells * sizeof(Tw)) != 0) {
          if (!dir) ps.fwd_by_next = false;  // no room: this pass keeps the compact form
          continue;
        }
        pl->matrices.push_back(mem);
        Tw* t = static_cast<Tw*>(mem);
        rc = gen_table(pl->field, t, cells, kTwist, ps.logn, log_cols, dir ? root_big_inv : root_big,
                       (dir && i == 0) ? finv : 1, col0, row0);
        // the kernels index the matrix by GLOBAL column: hand them the table shifted back by its first column
        (dir ? ps.inv_full : ps.fwd_full) = t - col0;
      }
    }
  }
  if (rc == XNTT_OK) {
    const int brc = be::stream_sync(nullptr);
    if (brc != 0) rc = be_fail(brc);
  }
  if (rc != XNTT_OK) {
    be::dev_free(pl->arena);
    for (void* mtx : pl->matrices) be::dev_free(mtx);
    delete pl;
    return rc;
  }
  *out = pl;
  return XNTT_OK;
}

int xntt_plan_destroy(xntt_plan* pl) {
  if (!pl) return XNTT_OK;
  {
    DeviceGuard g(pl->device);
    if (pl->arena) be::dev_free(pl->arena);
    for (void* mtx : pl->matrices) be::dev_free(mtx);
    if (pl->staging) be::dev_free(pl->staging);
    for (void* ev : pl->pipe_events) be::event_destroy(ev);
    for (void* st : pl->pipe_streams)
      if (st) be::stream_destroy(st);
  }
  delete pl;
  return XNTT_OK;
}

uint64_t xntt_plan_m(const xntt_plan* pl) { return pl ? (1ull << pl->log2_m) : 0; }
uint32_t xntt_plan_batch(const xntt_plan* pl) { return pl ? pl->batch : 0; }
uint32_t xntt_plan_launches(const xntt_plan* pl, int) { return pl ? (uint32_t)pl->passes.size() : 0; }
uint32_t xntt_plan_modmul(const xntt_plan* pl) { return pl ? pl->field.kind : 0; }
uint32_t xntt_plan_twiddle_form(const xntt_plan* pl, uint32_t pass, int inverse) {
  if (!pl || pass >= pl->passes.size() || !pl->passes[pass].col) return 0;
  const PassDesc& ps = pl->passes[pass];
  if (row_applies_twist(pl, pass, inverse != 0)) return 3;
  if (!inverse && ps.fwd_by_next) return 3;
  return (inverse ? ps.inv_full : ps.fwd_full) != nullptr ? 2 : 1;
}
uint32_t xntt_plan_tile_log2(const xntt_plan* pl, uint32_t pass) {
  if (!pl || pass >= pl->passes.size()) return 0;
  const PassDesc& ps = pl->passes[pass];
  return (uint32_t)(ps.logn + pass_logw(ps.logn, ps.narrow));
}
uint32_t xntt_plan_splits(const xntt_plan* pl, uint32_t* out, uint32_t n) {
  if (!pl) return 0;
  for (uint32_t i = 0; out && i < n && i < pl->passes.size(); ++i) out[i] = (uint32_t)pl->passes[i].logn;
  return (uint32_t)pl->passes.size();
}

int xntt_forward(const xntt_plan* pl, uint64_t* dst, const uint64_t* src, void* stream) {
  if (!pl) return XNTT_ERR_INVALID;
  if (pl->shard_count > 1) return XNTT_ERR_STATE;
  return run_range(pl, false, 0, pl->passes.size(), (u64*)dst, (const u64*)src, stream, false);
}
int xntt_inverse(const xntt_plan* pl, uint64_t* dst, const uint64_t* src, void* stream) {
  if (!pl) return XNTT_ERR_INVALID;
  if (pl->shard_count > 1) return XNTT_ERR_STATE;
  return run_range(pl, true, 0, pl->passes.size(), (u64*)dst, (const u64*)src, stream, false);
}

int xntt_forward_multiply(const xntt_plan* pl, uint64_t* dst, const uint64_t* src, const uint64_t* b_mont,
                          void* stream) {
  if (!pl || !b_mont) return XNTT_ERR_INVALID;
  if (pl->shard_count > 1) return XNTT_ERR_STATE;
  return run_range(pl, false, 0, pl->passes.size(), (u64*)dst, (const u64*)src, stream, false, (const u64*)b_mont);
}

int xntt_run_pass(const xntt_plan* pl, uint32_t pass, int inverse, uint64_t* dst, const uint64_t* src,
                  void* stream) {
  if (!pl || pass >= pl->passes.size()) return XNTT_ERR_INVALID;
  if (pl->shard_count > 1) return XNTT_ERR_STATE;
  return run_range(pl, inverse != 0, pass, pass + 1, (u64*)dst, (const u64*)src, stream, false);
}

int xntt_shard_forward_cols(const xntt_plan* pl, uint64_t* dst, const uint64_t* src, void* stream) {
  if (!pl || pl->shard_count < 2) return XNTT_ERR_STATE;
  return run_range(pl, false, 0, 1, (u64*)dst, (const u64*)src, stream, false);
}
int xntt_shard_forward_rows(const xntt_plan* pl, uint64_t* dst, const uint64_t* src, void* stream) {
  if (!pl || pl->shard_count < 2) return XNTT_ERR_STATE;
  return run_range(pl, false, 1, pl->passes.size(), (u64*)dst, (const u64*)src, stream, true);
}
int xntt_shard_inverse_rows(const xntt_plan* pl, uint64_t* dst, const uint64_t* src, void* stream) {
  if (!pl || pl->shard_count < 2) return XNTT_ERR_STATE;
  return run_range(pl, true, 1, pl->passes.size(), (u64*)dst, (const u64*)src, stream, true);
}
int xntt_shard_inverse_cols(const xntt_plan* pl, uint64_t* dst, const uint64_t* src, void* stream) {
  if (!pl || pl->shard_count < 2) return XNTT_ERR_STATE;
  return run_range(pl, true, 0, 1, (u64*)dst, (const u64*)src, stream, false);
}

int xntt_shard_forward_cols_chunk(const xntt_plan* pl, uint64_t* tiles, const uint64_t* src, uint32_t chunk,
                                  uint32_t nchunks, void* stream) {
  return shard_cols_chunk(pl, false, (u64*)tiles, (const u64*)src, chunk, nchunks, stream);
}
int xntt_shard_inverse_cols_chunk(const xntt_plan* pl, uint64_t* dst, const uint64_t* tiles, uint32_t chunk,
                                  uint32_t nchunks, void* stream) {
  return shard_cols_chunk(pl, true, (u64*)dst, (const u64*)tiles, chunk, nchunks, stream);
}
int xntt_shard_forward_rows_tiled(const xntt_plan* pl, uint64_t* dst, const uint64_t* tiles, uint32_t nchunks,
                                  void* stream) {
  if (!pl || !dst || !tiles || pl->shard_count < 2) return XNTT_ERR_STATE;
  if (!pl->fwd) return XNTT_ERR_STATE;
  if (pl->batch != 1) return XNTT_ERR_INVALID;
  StrideMap tm;
  int rc = tiled_map(pl, nchunks, tm);
  if (rc != XNTT_OK) return rc;
  DeviceGuard g(pl->device);
  if (!g.ok) return be_fail(1);
  const PassDesc& p1 = pl->passes[1];
  const u64 rows0 = (1ull << pl->passes[0].logn) / pl->shard_count;  // local rows of the n0 x n1 matrix
  const u64 units = p1.col ? rows0 : rows0;  // pass 1: one outer block (column mode) or one row per local row
  const u32 tpo = p1.col ? (u32)((1ull << p1.log_inner) >> tile_logw(p1.logn)) : 0u;
  rc = run_pass_mapped(pl, 1, false, (u64*)dst, (const u64*)tiles, tm, natural_map(pl, p1), units, tpo, 0, stream);
  if (rc != XNTT_OK) return rc;
  if (pl->passes.size() > 2)
    return run_range(pl, false, 2, pl->passes.size(), (u64*)dst, (const u64*)dst, stream, true);
  return XNTT_OK;
}
int xntt_shard_inverse_rows_tiled(const xntt_plan* pl, uint64_t* tiles, const uint64_t* src, uint64_t* work,
                                  uint32_t nchunks, void* stream) {
  if (!pl || !tiles || !src || pl->shard_count < 2) return XNTT_ERR_STATE;
  if (!pl->inv) return XNTT_ERR_STATE;
  if (pl->batch != 1) return XNTT_ERR_INVALID;
  StrideMap tm;
  int rc = tiled_map(pl, nchunks, tm);
  if (rc != XNTT_OK) return rc;
  DeviceGuard g(pl->device);
  if (!g.ok) return be_fail(1);
  const u64* cur = (const u64*)src;
  if (pl->passes.size() > 2) {
    if (!work) return XNTT_ERR_INVALID;
    rc = run_range(pl, true, 2, pl->passes.size(), (u64*)work, cur, stream, true);
    if (rc != XNTT_OK) return rc;
    cur = (const u64*)work;
  }
  const PassDesc& p1 = pl->passes[1];
  const u64 rows0 = (1ull << pl->passes[0].logn) / pl->shard_count;
  const u32 tpo = p1.col ? (u32)((1ull << p1.log_inner) >> tile_logw(p1.logn)) : 0u;
  return run_pass_mapped(pl, 1, true, (u64*)tiles, cur, natural_map(pl, p1), tm, rows0, tpo, 0, stream);
}

// Fused exchange: the pass next to the all-to-all stores straight into every rank's tiled buffer.
// peers[s] = base of rank s's buffer of m / G words, laid out [G (source rank)][n0/G][n1/G].
int xntt_shard_forward_cols_peer(const xntt_plan* pl, uint64_t* const* peers, const uint64_t* src, void* stream) {
  if (!pl || !peers || !src || pl->shard_count < 2 || pl->shard_count > 8) return XNTT_ERR_STATE;
  if (!pl->fwd) return XNTT_ERR_STATE;
  if (pl->batch != 1) return XNTT_ERR_INVALID;
  const PassDesc& p0 = pl->passes[0];
  const u64 G = pl->shard_count, n0 = 1ull << p0.logn, n1 = (1ull << pl->log2_m) / n0, block = n1 / G;
  if (n0 / G < 1) return XNTT_ERR_INVALID;
  DeviceGuard g(pl->device);
  if (!g.ok) return be_fail(1);
  StrideMap full{}, tile{};
  full.b1 = full.b2 = tile.b1 = tile.b2 = 31;
  full.s0 = block;
  full.outer = n0 * block;
  tile.s0 = block;  // row rho of the (n0/G) x (n1/G) tile
  tile.outer = 0;
  const u32 tpo = (u32)(block >> tile_logw(p0.logn));
  const u32 col0 = (u32)(block * pl->shard_rank);
  // my tile inside every destination buffer starts at rank * (n0/G) * (n1/G)
  return run_pass_mapped(pl, 0, false, nullptr, (const u64*)src, full, tile, 1, tpo, col0, stream, (u64* const*)peers,
                         (u32)log2u(n0 / G), (n0 / G) * block * pl->shard_rank);
}
int xntt_shard_inverse_rows_peer(const xntt_plan* pl, uint64_t* const* peers, const uint64_t* src, uint64_t* work,
                                 void* stream) {
  if (!pl || !peers || !src || pl->shard_count < 2 || pl->shard_count > 8) return XNTT_ERR_STATE;
  if (!pl->inv) return XNTT_ERR_STATE;
  if (pl->batch != 1) return XNTT_ERR_INVALID;
  DeviceGuard g(pl->device);
  if (!g.ok) return be_fail(1);
  const u64* cur = (const u64*)src;
  int rc;
  if (pl->passes.size() > 2) {
    if (!work) return XNTT_ERR_INVALID;
    rc = run_range(pl, true, 2, pl->passes.size(), (u64*)work, cur, stream, true);
    if (rc != XNTT_OK) return rc;
    cur = (const u64*)work;
  }
  const PassDesc& p1 = pl->passes[1];
  const u64 G = pl->shard_count, n0 = 1ull << pl->passes[0].logn, n1 = (1ull << pl->log2_m) / n0, block = n1 / G;
  const u64 rows0 = n0 / G;
  // destination (rank s): chunk layout [n0][n1/G], my rows start at rank * (n0/G); element (rho, c')
  StrideMap dm{};
  dm.b1 = dm.b2 = 31;
  dm.outer = block;  // next local row rho
  u32 peer_bits;
  if (p1.col) {
    const u64 n1b = 1ull << p1.log_inner;
    if (block % n1b) return XNTT_ERR_UNSUPPORTED;
    dm.s0 = n1b;
    peer_bits = (u32)log2u(block / n1b);
  } else {
    dm.s0 = 1;
    peer_bits = (u32)log2u(block);
  }
  const u32 tpo = p1.col ? (u32)((1ull << p1.log_inner) >> tile_logw(p1.logn)) : 0u;
  return run_pass_mapped(pl, 1, true, nullptr, cur, natural_map(pl, p1), dm, rows0, tpo, 0, stream, (u64* const*)peers,
                         peer_bits, rows0 * block * pl->shard_rank);
}

int xntt_forward_host(const xntt_plan* pl, uint64_t* dst, const uint64_t* src) {
  return host_roundtrip(pl, dst, src, false);
}
int xntt_inverse_host(const xntt_plan* pl, uint64_t* dst, const uint64_t* src) {
  return host_roundtrip(pl, dst, src, true);
}

int xntt_to_montgomery(const xntt_plan* pl, uint64_t* dst, const uint64_t* src, size_t n, void* st) {
  if (!pl || !dst || !src) return XNTT_ERR_INVALID;
  if (n == 0) return XNTT_OK;
  DeviceGuard g(pl->device);
  BE(be::launch_to_mont(pl->field, (u64*)dst, (const u64*)src, n, pl->r2, st));
  return XNTT_OK;
}
int xntt_from_montgomery(const xntt_plan* pl, uint64_t* dst, const uint64_t* src, size_t n, void* st) {
  if (!pl || !dst || !src) return XNTT_ERR_INVALID;
  if (n == 0) return XNTT_OK;
  DeviceGuard g(pl->device);
  BE(be::launch_from_mont(pl->field, (u64*)dst, (const u64*)src, n, st));
  return XNTT_OK;
}
int xntt_multiply_normalize(const xntt_plan* pl, uint64_t* dst, const uint64_t* a, const uint64_t* b, size_t n,
                            void* st) {
  if (!pl || !dst || !a || !b) return XNTT_ERR_INVALID;
  if (n == 0) return XNTT_OK;
  DeviceGuard g(pl->device);
  BE(be::launch_mulnorm(pl->field, (u64*)dst, (const u64*)a, (const u64*)b, n, st));
  return XNTT_OK;
}

int xntt_alloc_device(void** ptr, size_t bytes, int device) {
  if (!ptr) return XNTT_ERR_INVALID;
  *ptr = nullptr;
  if (device >= 0) BE(be::set_device(device));
  BE(be::dev_malloc(ptr, bytes));
  return XNTT_OK;
}
int xntt_free_device(void* ptr) {
  BE(be::dev_free(ptr));
  return XNTT_OK;
}
int xntt_alloc_pinned(void** ptr, size_t bytes) {
  if (!ptr) return XNTT_ERR_INVALID;
  *ptr = nullptr;
  BE(be::host_malloc_pinned(ptr, bytes));
  return XNTT_OK;
}
int xntt_free_pinned(void* ptr) {
  BE(be::host_free_pinned(ptr));
  return XNTT_OK;
}
int xntt_memcpy_h2d(void* dst, const void* src, size_t bytes, void* st) {
  BE(be::memcpy_h2d(dst, src, bytes, st));
  return XNTT_OK;
}
int xntt_memcpy_d2h(void* dst, const void* src, size_t bytes, void* st) {
  BE(be::memcpy_d2h(dst, src, bytes, st));
  return XNTT_OK;
}
int xntt_stream_synchronize(void* st) {
  BE(be::stream_sync(st));
  return XNTT_OK;
}

int xntt_transpose(uint64_t* dst, const uint64_t* src, uint64_t rows, uint64_t cols, uint64_t ld_dst, uint64_t ld_src,
                   void* stream) {
  if (!dst || !src) return XNTT_ERR_INVALID;
  if (rows == 0 || cols == 0) return XNTT_OK;
  if (ld_src < cols || ld_dst < rows) return XNTT_ERR_INVALID;
  if (dst == src && (rows != cols || ld_dst != ld_src)) return XNTT_ERR_INVALID;  // in place: square only
  if (rows > 0xffffffffull || cols > 0xffffffffull) return XNTT_ERR_INVALID;
  BE(be::launch_transpose((u64*)dst, (const u64*)src, rows, cols, ld_dst, ld_src, stream));
  return XNTT_OK;
}

int xntt_pointer_is_device(const void* ptr) {
  int k = 0;
  BE(be::pointer_is_device(ptr, &k));
  return k;
}

const char* xntt_strerror(int s) {
  switch (s) {
    case XNTT_OK:
      return "ok";
    case XNTT_ERR_INVALID:
      return "invalid argument";
    case XNTT_ERR_UNSUPPORTED:
      return "unsupported configuration";
    case XNTT_ERR_ALLOC:
      return "allocation failed";
    case XNTT_ERR_CUDA:
      return "CUDA error";
    case XNTT_ERR_STATE:
      return "operation not enabled for this plan";
    default:
      return "unknown status";
  }
}
const char* xntt_last_cuda_error(void) { return g_err.c_str(); }
const char* xntt_version(void) { return "xntt 0.1 sm_100a"; }
int xntt_device_count(void) {
  int n = 0;
  const int rc = be::device_count(&n);
  if (rc != 0) return be_fail(rc);
  return n;
}

// ---- Kinnaes' formula (examples/magic-series-kinnaes/kinnaes.hpp) --------------------------------------------
namespace {
// one small result buffer per device, created on first use, released by xntt_release_scratch()
std::mutex g_kinnaes_mu;
void* g_kinnaes_scratch[64] = {};

int kinnaes_sum_impl(u64 p, u64 gen, u64 m, u64 n, u64 j_begin, u64 j_end, int device, u64* result) {
  if ((p & 1) == 0 || p < 3 || gen == 0 || !h_is_prime(p)) return XNTT_ERR_INVALID;
  if (m < 2 || m >= (1ull << 20) || n == 0 || (p - 1) % n != 0) return XNTT_ERR_INVALID;
  if (j_begin > j_end || j_end > n / 2) return XNTT_ERR_INVALID;
  // the kernel builds w^J from ladder[i] = w^(2^i), i < kKinnaesLadder: J must fit
  if (j_end >> kKinnaesLadder) return XNTT_ERR_INVALID;
  *result = 0;
  if (j_begin == j_end) return XNTT_OK;  // empty sum: 0 / 1
  int dev = device;
  if (dev < 0) BE(be::get_device(&dev));
  DeviceGuard g(dev);
  if (!g.ok) return be_fail(1);
  KinnaesParams prm{};
  prm.field.p = p;
  prm.field.pinv = h_montgomery_inverse(p);
  prm.field.one = h_to_mont(1, p);
  prm.m = m;
  prm.j_first = j_begin + 1;
  prm.count = j_end - j_begin;
  prm.exp_num = m * m - m + 1;
  prm.exp_r = m * (m - 1) / 2 * m;
  u64 w = h_pow(gen % p, (p - 1) / n, p);  // Modulus::get_root_forward(n)
  for (int i = 0; i < kKinnaesLadder; ++i) {
    prm.ladder[i] = h_to_mont(w, p);
    w = h_mul(w, w, p);
  }
  u64 blocks = (prm.count + kKinnaesThreads - 1) / kKinnaesThreads;
  if (blocks > kKinnaesMaxBlocks) blocks = kKinnaesMaxBlocks;
  // one small result buffer per device, kept for the life of the process (a cudaMalloc / cudaFree pair per call
  // costs 5-7 ms, thirty times the kernel); calls are serialised on it
  if (dev >= 64) return XNTT_ERR_INVALID;
  std::lock_guard<std::mutex> lock(g_kinnaes_mu);
  void** scratch = g_kinnaes_scratch;
  if (!scratch[dev]) BE(be::dev_malloc(&scratch[dev], kKinnaesMaxBlocks * 2 * sizeof(u64)));
  void* dpart = scratch[dev];
  prm.partial = static_cast<u64*>(dpart);
  std::vector<u64> part(blocks * 2);
  int rc = XNTT_OK, brc;
  if ((brc = be::launch_kinnaes(prm, (unsigned)blocks, nullptr)) != 0) rc = be_fail(brc);
  if (rc == XNTT_OK && (brc = be::memcpy_d2h(part.data(), dpart, part.size() * sizeof(u64), nullptr)) != 0) rc = be_fail(brc);
  if (rc == XNTT_OK && (brc = be::stream_sync(nullptr)) != 0) rc = be_fail(brc);
  if (rc != XNTT_OK) return rc;
  // fold the per-CTA fractions and divide once (kinnaes.hpp:143-156); from Montgomery form: x * 2^-64
  const u64 rinv = h_inv(h_to_mont(1, p), p);
  u64 num = 0, den = 1;
  for (u64 b = 0; b < blocks; ++b) {
    const u64 nb = h_mul(part[2 * b], rinv, p), db = h_mul(part[2 * b + 1], rinv, p);
    num = (u64)(((u128)h_mul(den, nb, p) + h_mul(num, db, p)) % p);
    den = h_mul(den, db, p);
  }
  if (den == 0) return XNTT_ERR_INVALID;  // n is not the order of a usable root for this m
  *result = h_mul(num, h_inv(den, p), p);
  return XNTT_OK;
}
}  // namespace

int xntt_kinnaes_sum(uint64_t modulus, uint64_t generator, uint64_t m, uint64_t n, uint64_t j_begin, uint64_t j_end,
                     int device, uint64_t* result) {
  u64 r = 0;
  const int rc = kinnaes_sum_impl(modulus, generator, m, n, j_begin, j_end, device, &r);
  if (result) *result = r;
  return result ? rc : XNTT_ERR_INVALID;
}

int xntt_kinnaes_compute(uint64_t modulus, uint64_t generator, uint64_t m, uint64_t n, int device, uint64_t* result) {
  if (!result) return XNTT_ERR_INVALID;
  u64 sum = 0;
  const int rc = kinnaes_sum_impl(modulus, generator, m, n, 0, n / 2, device, &sum);
  if (rc != XNTT_OK) return rc;
  const u64 p = modulus;
  sum = (u64)(((u128)sum + sum) % p);
  // compute_comb(m * m, m) (kinnaes.hpp:36-47)
  const u64 a = m * m;
  u64 num = a % p, den = m % p;
  for (u64 i = 1; i < m; ++i) num = h_mul(num, (a - i) % p, p);
  for (u64 i = 2; i < m; ++i) den = h_mul(den, i % p, p);
  sum = (u64)(((u128)sum + h_mul(num, h_inv(den, p), p)) % p);
  *result = h_mul(sum, h_inv(n % p, p), p);
  return XNTT_OK;
}

int xntt_release_scratch(void) {
  std::lock_guard<std::mutex> lock(g_kinnaes_mu);
  int rc = XNTT_OK;
  for (int dev = 0; dev < 64; ++dev) {
    if (!g_kinnaes_scratch[dev]) continue;
    DeviceGuard g(dev);
    if (!g.ok || be::dev_free(g_kinnaes_scratch[dev]) != 0) rc = be_fail(1);
    g_kinnaes_scratch[dev] = nullptr;
  }
  return rc;
}

int xntt_microbench(int kind, int iters, double* gops, double* ms) {
  if (kind < 0 || kind > 5 || iters <= 0 || !gops) return XNTT_ERR_INVALID;
  BE(be::microbench(kind, iters, gops, ms));
  return XNTT_OK;
}

#pragma GCC visibility pop
}  // extern "C"
