// SPDX-License-Identifier: Apache-2.0
// Plain-C++ types shared by the planner (plan.cpp, no CUDA headers) and the kernels.
#pragma once
#include <cstddef>
#include <cstdint>

namespace xntt {

typedef unsigned long long u64;
typedef unsigned int u32;

// A twiddle in the reference's PAdic64 storage format (include/sventt/modmul/sve/p-adic-64.hpp:64-95):
struct alignas(16) Tw {
  u64 w;   // omega * 2^64 mod P         (to_montgomery)
  u64 wp;  // w * P^-1 mod 2^64          (precompute)
};

constexpr int kThreads = 256;
constexpr int kTileLog = 13;    // a tile holds 2^13 residues = 64 KiB of shared memory
constexpr int kMaxRowLog = 13;  // longest transform one CTA keeps in shared memory (row mode)
constexpr int kMaxColLog = 12;  // longest column transform (column mode, 2 columns wide)
constexpr int kMaxTileWLog = 5;

constexpr int tile_logw(int logn) {
  return logn >= kTileLog ? 0 : (kTileLog - logn < kMaxTileWLog ? kTileLog - logn : kMaxTileWLog);
}
// Tuning knobs (compile time): residues per thread-task column group and resident CTAs per SM.
#ifndef XNTT_FORCE_C1
#define XNTT_FORCE_C1 0
#endif
#ifndef XNTT_MINB
#define XNTT_MINB 2
#endif
// ... except plain row passes with one residue per slot (2^13 rows, narrow 2^11 rows; no twiddle matrix, no point-wise
// product): three resident CTAs at 80 registers.  Measured on B200 (profiles/r4_variants_row3.log): 2048 x 2^13 forward
// 191.9 -> 183.6 us, inverse 212.9 -> 203.1 us; the same rows with the twiddle matrix lose (218.9 -> 227.1 us forward)
// and every two-residue kernel spills, so those stay at two.
#ifndef XNTT_MINB_ROW1
#define XNTT_MINB_ROW1 3
#endif
constexpr int pass_minb(bool col, int c, bool plain) { return (!col && c == 1 && plain) ? XNTT_MINB_ROW1 : XNTT_MINB; }
constexpr int tile_c(int logn) { return (!XNTT_FORCE_C1 && tile_logw(logn) >= 1) ? 2 : 1; }

// Narrow tiles: a quarter of the residues of a whole tile, i.e. four times the CTAs and a quarter of the work per
// thread - for transforms too small to fill the GPU with whole tiles (a 2^17 transform is 16 whole tiles on 148 SMs;
// what it costs then is the life of ONE CTA, about 2 us per butterfly level).  Column tiles stay at least four columns
// (32 bytes) wide.
constexpr int kNarrowShift = 2;
constexpr bool has_narrow_tile(int logn, bool col) { return tile_logw(logn) - kNarrowShift >= (col ? 2 : 0); }
constexpr int pass_logw(int logn, bool narrow) { return narrow ? tile_logw(logn) - kNarrowShift : tile_logw(logn); }
constexpr int pass_c(int logn, bool narrow) { return (!XNTT_FORCE_C1 && pass_logw(logn, narrow) >= 1) ? 2 : 1; }

// The production prime of the reference README (README.md:19): 2^64 - 1827*2^31 + 1.
constexpr u64 kP0 = 0xfffffc6e80000001ULL;
// Goldilocks, 2^64 - 2^32 + 1: the other 64-bit modulus of the reference's tests (tests/test-ntt-reference.cpp:17-23,
// tests/test-modulus.cpp, examples/magic-series/test-magic-series.cpp:22-39) - also gets kernels with the modulus baked in
constexpr u64 kPGold = 0xffffffff00000001ULL;

// Generalised addressing of the transform index k (used by the passes next to the all-to-all of a
// sharded plan, where the exchange leaves / expects the data tiled): k is cut into three bit fields
// [0, b1), [b1, b2), [b2, ...) with their own strides; `outer` is the stride of the outer block (column
// mode) or of a row (row mode).
struct StrideMap {
  u64 s0, s1, s2;
  u64 outer;
  u32 b1, b2;
};

// Field constants of a runtime modulus (ignored by the kernels specialised for kP0).
struct FieldConsts {
  u64 p;     // modulus
  u64 pinv;  // p^-1 mod 2^64
  u64 one;   // 2^64 mod p
  // arithmetic of the pass kernels: 0 = Montgomery (PAdic64), 1 = Shoup / FixedPoint64 (moduli below 2^62 only;
  // twiddle tables then hold (omega, floor(omega * 2^64 / p)) instead of the Montgomery pair)
  u32 kind;
  u32 reserved_;
};
constexpr u32 kFieldMontgomery = 0, kFieldShoup = 1;

struct PassParams {
  const u64* src;
  u64* dst;
  const Tw* tw;        // forward: G[0 .. N/2), inverse: I[1 .. N)
  const Tw* twist_lo;  // omega_M^(+-e), e < 2^twist_shift            (null when no twist)
  const Tw* twist_hi;  // omega_M^(+-e * 2^twist_shift) (* 1/inverse_factor on the inverse side)
  const Tw* twist_full;  // optional: the whole twiddle matrix, entry (k << twist_full_shift) + column; when set
                         // the twist is one streamed 16-byte load and one Montgomery product per residue
  u64 inner;           // column mode: elements between consecutive k; row mode: unused
  u64 outer_stride;    // column mode: elements between consecutive outer blocks (= N * inner)
  u32 tiles_per_outer; // column mode: inner / W
  const Tw* pre_twist;   // row pass: the twiddle matrix of the column pass next to it (entry ((row & pre_rows_mask) <<
                         // log2 N) + k), applied while the rows are loaded (forward) or before they are stored
                         // (inverse); that column pass then runs without a twiddle
  u32 pre_rows_mask;
  // forward column pass applying the matrix of the OUTER pass before it (kColPre): entry
  // ((outer block & pre_rows_mask) << pre_shift) + (k << pre_kshift) + column
  u32 pre_shift, pre_kshift;
  u32 twist_shift;
  u32 twist_full_shift;  // log2 of the number of columns of twist_full
  u32 twist_col0;      // column mode: global index of this buffer's first column (sharded plans)
  u32 scale_on;        // inverse row mode: multiply outputs by `scale` (else just canonicalise)
  u32 rows;            // row mode: number of valid rows in the buffer (tiles may be ragged)
  Tw scale;
  FieldConsts field;
  StrideMap smap, dmap;  // MAP kernels only: address maps of src and dst
  // MAP kernels only, peer_on != 0: output word with transform index k is stored through
  // peer[k >> peer_bits] (another GPU's buffer, mapped over NVLink) at dmap(k & (2^peer_bits - 1)) -
  // the all-to-all of a sharded transform fused into the pass that produces the data
  // (peer_bits == 0 is a legal value: every k of the pass belongs to another rank)
  u64* peer[8];
  u32 peer_bits;
  u32 peer_on;
  const u64* pointwise;  // forward row pass: multiply output word i by pointwise[i] * 2^-64 (fused
                         // PAdic64::multiply_normalize against a to_montgomery'd spectrum), or null
  u32 narrow;            // run the narrow-tile kernel of this pass length (pass_logw(logn, true); plain addressing only)
  u32 lazy_out;          // inverse column pass: store lazy residues (any u64 standing for v mod P) - set by the planner
                         // when the pass that consumes them begins with the Montgomery product by its own twiddle
                         // (the outer column pass of a three-pass plan), which takes any 64-bit value
};
// fields whose kernels exist in the narrow-tile form: the production prime and runtime Montgomery
inline bool field_has_narrow(const FieldConsts& fc) {
  return fc.p == kP0 || (fc.kind == kFieldMontgomery && fc.p != kPGold);
}

// Input of the on-device table generator: out[idx] = scale * root^e(idx), Montgomery pair.
struct PowTable {
  u64 sq[32];  // root^(2^i) in Montgomery form
  u64 scale;   // Montgomery form of the extra factor (2^64 mod P for none)
  u32 col0;    // kTwist: global index of the table's first column (a rank's column block of a sharded plan)
  u32 row0;    // kTwist: index of the table's first row (a rank's row block)
};
// Kinnaes sum (kinnaes_kernel.cuh)
constexpr int kKinnaesThreads = 256;
constexpr int kKinnaesLadder = 40;
constexpr unsigned kKinnaesMaxBlocks = 148 * 8;  // one resident wave of 256-thread CTAs

struct KinnaesParams {
  FieldConsts field;
  u64 m;          // order of the magic series
  u64 j_first;    // first J (= j_begin + 1)
  u64 count;      // number of consecutive J
  u64 exp_num;    // m^2 - m + 1
  u64 exp_r;      // r = m (m-1)/2 * m
  u64 ladder[kKinnaesLadder];  // w^(2^i) in Montgomery form
  u64* partial;   // out: [2 * gridDim.x] (numerator, denominator) per CTA, Montgomery form
};

enum TableKind { kFwdG = 0, kInvI = 1, kPowers = 2, kTwist = 3 };

}  // namespace xntt
