// SPDX-License-Identifier: Apache-2.0
//
// One "pass" of the (blocked) six-step NTT: every CTA owns a tile of W independent length-N
// sub-transforms, keeps it resident in shared memory, and runs all log2(N) butterfly levels on it
// as radix-8 (first stage radix 4/8/16) register passes with swizzled shared-memory exchanges and
// barriers that only span the threads an exchange connects.
//
//   column mode : data is [outer][N][inner]; the tile is N rows x W contiguous columns
//                 (element (k, c) at base + k*inner + c).  This is the column phase of
//                 sventt::BlockedGenericSVELayer::compute_forward_without_twiddle
//                 (include/sventt/layer/sve/blocked-generic.hpp:121-155): its tile transpose-in /
//                 inner NTT / transpose-out becomes "strided tile load -> NTT -> strided store",
//                 and the row twiddle of sventt::GenericSVELayer::twiddle_rows_forward
//                 (include/sventt/layer/sve/generic.hpp:169-268) is fused into a pass boundary (TwistKind).
//   row mode    : data is [rows][N]; the tile is W whole rows (element (k, c) at base + c*N + k).
//                 This is the inner kernel call of sventt::RecursiveNTT::compute_forward
//                 (include/sventt/kernel/recursive.hpp:69-74).
//
// Butterfly network.  Forward uses Cooley-Tukey butterflies on natural-order input producing
// bit-reversed output: level s (half-length N/2^(s+1)) multiplies the upper half of block b by
// G[b] = omega_N^bitrev_{log2N-1}(b) - one table of N/2 entries serves every level.  Inverse is
// decimation-in-time on bit-reversed input: level with half-length l multiplies position j by
// I[l + j] = omega_{2l}^-j.  Both are "multiply, then add/sub", which is what lets every
// intermediate value stay lazy (see field.cuh).  The function computed is exactly that of
// NTTReference::compute_forward / compute_inverse (tests/ntt-reference.hpp:43-83): same
// bit-reversed order, canonical outputs.
#pragma once
#if !defined(XNTT_HOST_EMU)
#include <cuda_runtime.h>
#if defined(XNTT_TMA_ROWS) && XNTT_TMA_ROWS
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#endif
#endif

#include <cstdint>
#include <utility>

#include "field.cuh"
#include "params.h"

namespace xntt {

// What a pass fuses besides its butterflies - a template parameter (tested at run time, either costs the plain
// path 3-4 %):
//   column passes : the six-step twiddle after the last forward / before the first inverse stage, from two
//                   sqrt(M)-entry tables (kCompactTwist) or from the whole matrix (kFullTwist); kNoTwist when the
//                   row pass behind it applies the matrix instead
//   forward rows  : kPreTwist = multiply by the preceding column pass's twiddle matrix while loading (contiguous
//                   rows of the matrix next to contiguous rows of data, at the start of a tile where the latency
//                   overlaps the data load - the forward twin of what the inverse column pass does);
//                   kPointwise = the point-wise product of a polynomial multiply before storing; or both
//   inverse rows  : kPostTwist = the mirror image: multiply by the matrix of the column pass that follows, before
//                   storing (which also canonicalises); that column pass then runs twist-free
//   forward columns: kColPre = multiply by the matrix of the OUTER pass before this one while loading (that pass then
//                   runs twist-free, like the last column pass does when the row pass applies its matrix): every pass of
//                   a three-pass plan at one modular product per residue
enum TwistKind {
  kNoTwist = 0, kCompactTwist = 1, kFullTwist = 2, kPointwise = 3, kPreTwist = 4, kPrePointwise = 5, kPostTwist = 6,
  kColPre = 7
};

// the one place that decides (CUDA dispatcher and host emulator both call it)
inline int pass_kind(bool col, bool inverse, bool map, const PassParams& prm) {
  if (col) {
    if (!inverse && prm.pre_twist != nullptr) return kColPre;
    if (prm.twist_full != nullptr) return kFullTwist;
    if (prm.twist_lo != nullptr) return kCompactTwist;
    return kNoTwist;
  }
  if (!inverse && !map) {
    if (prm.pre_twist != nullptr) return prm.pointwise != nullptr ? kPrePointwise : kPreTwist;
    if (prm.pointwise != nullptr) return kPointwise;
  }
  if (inverse && !map && prm.pre_twist != nullptr) return kPostTwist;
  return kNoTwist;
}

// Stage radices of a pass of length 2^LOGN: the first forward (= last inverse) stage has radix 2^LOGR1, every other
// stage radix 2^LOGRN.  LOGRN = 3 (radix 8) everywhere, except - with XNTT_RADIX16 - tiles with one residue per slot
// (C = 1, in the default build the 2^13 rows), which run radix 16 with a first stage of up to radix 32: 13 = 5 + 4 + 4,
// one shared-memory exchange less than 4 + 3 + 3 + 3.  Measured on B200: no gain (2^13 rows 229.5 vs 230.1 us forward,
// 204.5 vs 205.2 us inverse at 128 instead of 88-108 registers; with every tile at C = 1 the 2^12 rows gain 3 % forward
// and lose 5 % inverse) - the exchanges are not what the passes wait for - so it stays off.
#ifndef XNTT_RADIX16
#define XNTT_RADIX16 0
#endif
template <int LOGN, int C>
struct Stages {
  static constexpr int LOGRN = (XNTT_RADIX16 && C == 1) ? 4 : 3;
  static constexpr int kRem = LOGN % 3;
  static constexpr int LOGR1 = LOGRN == 3 ? (LOGN <= 4 ? LOGN : (kRem == 0 ? 3 : (kRem == 1 ? 4 : 2)))
                                          : (LOGN <= 5 ? LOGN : LOGN - 4 * ((LOGN - 5 + 3) / 4));
  static constexpr int NS = 1 + (LOGN - LOGR1) / LOGRN;
};

template <int C>
struct Slot;
template <>
struct Slot<1> {
  typedef u64 type;
  static constexpr int kSwzMask = 15;
};
template <>
struct Slot<2> {
  typedef ulonglong2 type;
  static constexpr int kSwzMask = 7;
};

template <int C, int LOGRN>
__device__ __forceinline__ int swz(int k) {
  return k ^ ((k >> LOGRN) & Slot<C>::kSwzMask);
}

template <int LOGN_, int LOGW_, int C_, bool COL_, bool MAP_ = false, bool TMA_ = false>
struct PassCfg {
  static constexpr int LOGN = LOGN_, LOGW = LOGW_, C = C_;
  static constexpr bool COL = COL_, MAP = MAP_, TMA = TMA_;
  static constexpr int N = 1 << LOGN, W = 1 << LOGW;
  static constexpr int NP = W / C;  // column groups per tile
  static constexpr int LOGNP = LOGW - (C == 2 ? 1 : 0);
  static constexpr int LOGRN = Stages<LOGN, C>::LOGRN;
  static constexpr int LOGR1 = Stages<LOGN, C>::LOGR1;
  static constexpr int NS = Stages<LOGN, C>::NS;
  static constexpr size_t kSmemBytes = NS > 1 ? (size_t)N * W * sizeof(u64) : 0;
  static_assert(W >= C, "tile narrower than a column group");
};

// Slot of element k0 + (r << logs) of a task.  The swizzle is linear over GF(2) and k0 has no bits where
// r << logs has any, so swz(k0 + (r << logs)) = swz(k0) ^ swz(r << logs) with the second term a compile-time
// constant: one swizzle per task, then an XOR with the constant's low bits (often none) plus a constant offset
// that folds into the address immediate - instead of shift / xor / shift-add per slot.
template <class Cfg, int LOGS, int R>
__device__ __forceinline__ void task_slots(int k0, int p, int (&idx)[R]) {
  constexpr int MB = Slot<Cfg::C>::kSwzMask == 15 ? 4 : 3;  // the swizzle rewrites bits [0, MB)
  constexpr int LOW = (1 << MB) - 1;
  const int b = swz<Cfg::C, Cfg::LOGRN>(k0);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int cr = (r << LOGS) ^ (((r << LOGS) >> Cfg::LOGRN) & Slot<Cfg::C>::kSwzMask);
    // above bit MB the constant only has bits of r << logs, which b does not have: add (-> address immediate)
    const int s = LOGS >= MB ? ((b ^ (cr & LOW)) + (cr & ~LOW)) : (b ^ cr);
    if constexpr (Cfg::COL)
      idx[r] = (s << Cfg::LOGNP) + p;
    else
      idx[r] = (p << Cfg::LOGN) + s;
  }
}

template <class Cfg, int LOGS, int R>
__device__ __forceinline__ void smem_load(const typename Slot<Cfg::C>::type* sm, int k0, int p, u64 (&x)[R][Cfg::C]) {
  int idx[R];
  task_slots<Cfg, LOGS, R>(k0, p, idx);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    auto v = sm[idx[r]];
    if constexpr (Cfg::C == 2) {
      x[r][0] = v.x;
      x[r][1] = v.y;
    } else {
      x[r][0] = v;
    }
  }
}

template <class Cfg, int LOGS, int R>
__device__ __forceinline__ void smem_store(typename Slot<Cfg::C>::type* sm, int k0, int p, const u64 (&x)[R][Cfg::C]) {
  int idx[R];
  task_slots<Cfg, LOGS, R>(k0, p, idx);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if constexpr (Cfg::C == 2) {
      sm[idx[r]] = make_ulonglong2(x[r][0], x[r][1]);
    } else {
      sm[idx[r]] = x[r][0];
    }
  }
}

__device__ __forceinline__ u64 map_k(const StrideMap& m, int k) {
  const u32 uk = (u32)k;
  return (u64)(uk & ((1u << m.b1) - 1u)) * m.s0 + (u64)((uk >> m.b1) & ((1u << (m.b2 - m.b1)) - 1u)) * m.s1 +
         (u64)(uk >> m.b2) * m.s2;
}

// Global addressing of element (k, column group p, lane c) relative to the tile base.
template <class Cfg>
__device__ __forceinline__ u64 gofs(const PassParams& prm, const StrideMap& map, int k, int p, int c) {
  if constexpr (Cfg::MAP) {
    if constexpr (Cfg::COL)
      return map_k(map, k) + (u64)(p * Cfg::C + c);
    else
      return (u64)(p * Cfg::C + c) * map.outer + map_k(map, k);
  } else {
    if constexpr (Cfg::COL)
      return (u64)k * prm.inner + (u64)(p * Cfg::C + c);
    else
      return ((u64)(p * Cfg::C + c) << Cfg::LOGN) + (u64)k;
  }
}

// Optional streaming (L1::no_allocate) loads of tile data.  Measured on B200: no gain over plain loads
// (2^24 forward 433.6 vs 434.4 us, some inverse shapes 3-7 % slower), so it stays off.
#ifndef XNTT_LD_NA
#define XNTT_LD_NA 0
#endif
__device__ __forceinline__ u64 ld_stream(const u64* p) {
#if XNTT_LD_NA && !defined(XNTT_HOST_EMU)
  u64 v;
  asm volatile("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
#else
  return *p;
#endif
}
__device__ __forceinline__ ulonglong2 ld_stream2(const u64* p) {
#if XNTT_LD_NA && !defined(XNTT_HOST_EMU)
  ulonglong2 v;
  asm volatile("ld.global.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
  return v;
#else
  return *reinterpret_cast<const ulonglong2*>(p);
#endif
}

// XNTT_PROBE_NOMEM (development probe, never a product build): compile the global loads and stores of tile data out -
// loads become register arithmetic, stores are predicated on a value that never occurs - to see what the memory phases
// of a pass cost ("Why not TMA"; tools/gpu_variants.py times the variant next to the product).
#ifndef XNTT_PROBE_NOMEM
#define XNTT_PROBE_NOMEM 0
#endif
// XNTT_TMA_ROWS (measured variant, off in the product build): the 2^13 row pass stages its tile through a bulk tensor
// copy (cp.async.bulk.tensor.2d + mbarrier, SASS UTMALDG) instead of per-thread LDG, see row_tile_tma_load below.
#ifndef XNTT_TMA_ROWS
#define XNTT_TMA_ROWS 0
#endif
// XNTT_PREFETCH_TABLE: prefetch the pass's twiddle table into L1 ahead of griddepcontrol.wait.  Measured on B200: no gain
// (one 2^10 .. 2^19 transform within 0.1 us either way, 2^20 27.8 -> 28.5 us, 2^24 inverse 382.5 -> 387.7 us), so it stays off.
#ifndef XNTT_PREFETCH_TABLE
#define XNTT_PREFETCH_TABLE 0
#endif

#if XNTT_TMA_ROWS && !defined(XNTT_HOST_EMU)
// The tile (2^13 contiguous residues = 64 KiB) is described to the TMA unit as 512 rows of 16 residues (128 bytes) with
// the 128-byte swizzle: 16-byte chunk c of row r lands at chunk c ^ (r & 7).  As a slot index (8-byte slots):
//   pos_T(k) = k ^ (((k >> 4) & 7) << 1)
// Both first stages read conflict-free from that layout (forward: a warp reads 32 consecutive k; inverse: every lane
// reads 8 consecutive k, 16-byte accesses of 8 lanes spread over all 8 chunks) and write the kernel's own swizzle
// swz(k) = k ^ ((k >> 3) & 15) in place: both permutations stay inside aligned groups of 16 slots, and the tasks that
// own one group in stage 0 sit in the same warp, so a __syncwarp between the reads and the writes is all it takes.
__device__ __forceinline__ int tma_slot(int k) { return k ^ (((k >> 4) & 7) << 1); }

__device__ __forceinline__ void row_tile_tma_load(void* smem, unsigned long long* mbar, const void* tmap, u32 tile) {
  const unsigned sm = (unsigned)__cvta_generic_to_shared(smem), mb = (unsigned)__cvta_generic_to_shared(mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(65536u) : "memory");
#pragma unroll
    for (int j = 0; j < 2; ++j)  // two boxes of 16 x 256 residues
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
              sm + j * 32768u),
          "l"(tmap), "r"(mb), "r"(0), "r"((int)(tile * 512u + j * 256u))
          : "memory");
  }
  // every thread waits for the bytes (phase 0)
  asm volatile(
      "{\n\t.reg .pred p;\n\tTMA_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra TMA_DONE;\n\t"
      "bra TMA_WAIT;\n\tTMA_DONE:\n\t}" ::"r"(mb)
      : "memory");
}
#endif

template <class Cfg, int R>
__device__ __forceinline__ void gmem_load(const PassParams& prm, const u64* base, u32 row0, int k0, int logs,
                                          int p, u64 (&x)[R][Cfg::C]) {
#if XNTT_PROBE_NOMEM
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int c = 0; c < Cfg::C; ++c) x[r][c] = (u64)(threadIdx.x + 7 * r + c) * 0x9e3779b97f4a7c15ull + (u64)(k0 + p + row0 + logs);
  (void)prm;
  (void)base;
  return;
#endif
  if constexpr (Cfg::COL) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int k = k0 + (r << logs);
      if constexpr (Cfg::C == 2) {
        ulonglong2 v = ld_stream2(base + gofs<Cfg>(prm, prm.smap, k, p, 0));
        x[r][0] = v.x;
        x[r][1] = v.y;
      } else {
        x[r][0] = ld_stream(base + gofs<Cfg>(prm, prm.smap, k, p, 0));
      }
    }
  } else {
    // row mode: the last tile may hold fewer than W rows
#pragma unroll
    for (int c = 0; c < Cfg::C; ++c) {
      const bool ok = row0 + (u32)(p * Cfg::C + c) < prm.rows;
#pragma unroll
      for (int r = 0; r < R; ++r) x[r][c] = ok ? ld_stream(base + gofs<Cfg>(prm, prm.smap, k0 + (r << logs), p, c)) : 0ull;
    }
  }
}

template <class Cfg, int R>
__device__ __forceinline__ void gmem_store(const PassParams& prm, u64* base, u32 row0, int k0, int logs, int p,
                                           const u64 (&x)[R][Cfg::C]) {
#if XNTT_PROBE_NOMEM
  // keeps the arithmetic alive, stores nothing (the value never occurs)
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int c = 0; c < Cfg::C; ++c)
      if (x[r][c] == 0x0123456789abcdefull) base[(u64)(k0 + (r << logs)) + p + c + row0] = x[r][c];
  (void)prm;
  return;
#endif
  if constexpr (Cfg::MAP) {
    if (prm.peer_on != 0) {
      // fused exchange: the owner of output index k is rank k >> peer_bits; all ranks' buffers share one
      // layout, so the tile offset (base - prm.dst) carries over
      const u64 tile_ofs = (u64)(base - prm.dst);
      const int kmask = (1 << prm.peer_bits) - 1;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int k = k0 + (r << logs);
        u64* pb = prm.peer[(u32)k >> prm.peer_bits] + tile_ofs;
        if constexpr (Cfg::COL && Cfg::C == 2) {
          *reinterpret_cast<ulonglong2*>(pb + gofs<Cfg>(prm, prm.dmap, k & kmask, p, 0)) =
              make_ulonglong2(x[r][0], x[r][1]);
        } else {
#pragma unroll
          for (int c = 0; c < Cfg::C; ++c)
            if (Cfg::COL || row0 + (u32)(p * Cfg::C + c) < prm.rows)
              pb[gofs<Cfg>(prm, prm.dmap, k & kmask, p, c)] = x[r][c];
        }
      }
      return;
    }
  }
  if constexpr (Cfg::COL) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if constexpr (Cfg::C == 2)
        *reinterpret_cast<ulonglong2*>(base + gofs<Cfg>(prm, prm.dmap, k0 + (r << logs), p, 0)) =
            make_ulonglong2(x[r][0], x[r][1]);
      else
        base[gofs<Cfg>(prm, prm.dmap, k0 + (r << logs), p, 0)] = x[r][0];
    }
  } else {
#pragma unroll
    for (int c = 0; c < Cfg::C; ++c) {
      if (row0 + (u32)(p * Cfg::C + c) < prm.rows) {
        if (logs == 0 && R >= 2) {
          // a thread owns R consecutive outputs of this row: 128-bit stores
#pragma unroll
          for (int r = 0; r + 1 < R; r += 2)
            *reinterpret_cast<ulonglong2*>(base + gofs<Cfg>(prm, prm.dmap, k0 + r, p, c)) =
                make_ulonglong2(x[r][c], x[r + 1][c]);
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) base[gofs<Cfg>(prm, prm.dmap, k0 + (r << logs), p, c)] = x[r][c];
        }
      }
    }
  }
}

__device__ __forceinline__ Tw ld_tw(const Tw* t) {
  ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(t));
  Tw r;
  r.w = v.x;
  r.wp = v.y;
  return r;
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
#if !defined(XNTT_HOST_EMU)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
// Six-step twiddle of element (k, global column col): omega_M^(bitrev_LOGN(k) * col).  Two forms:
//   * prm.twist_full set: the plan holds the whole twiddle matrix in the layout of the data (entry (k, col)
//     next to entry (k, col + 1)); the twist is one streamed load and ONE Montgomery product.  The load is as
//     coalesced as the tile store itself and never touches L1 (the random hi/lo look-ups below are what made the
//     kernels L1-size sensitive: 2^24 forward 433 -> 506 us when L1 shrinks from 96 to 32 KiB).
//   * otherwise hi[e >> shift] * lo[e & mask], two Montgomery products (any size, tables of 2 * sqrt(M) entries).
// Both results are canonical.
__device__ __forceinline__ Tw ld_tw_stream(const Tw* t) {
#if !defined(XNTT_HOST_EMU)
  Tw r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(r.w), "=l"(r.wp) : "l"(t));
  return r;
#else
  return *t;
#endif
}
// all R x C residues of a task; the table form is a template parameter (a run-time test costs the compact form
// 3.5 %: 2^24 forward 446 instead of 430 us)
template <class F, class Cfg, int R, int TWIST>
__device__ __forceinline__ void apply_twist(const F& f, const PassParams& prm, u64 (&x)[R][Cfg::C], int k0, int logs,
                                            u32 col) {
  if constexpr (TWIST == kFullTwist) {
    Tw t[R][Cfg::C];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < Cfg::C; ++c)
        t[r][c] = ld_tw_stream(prm.twist_full + (((u64)(u32)(k0 + (r << logs)) << prm.twist_full_shift) + (col + c)));
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < Cfg::C; ++c) x[r][c] = f.mont(x[r][c], t[r][c]);
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < Cfg::C; ++c) {
        const u32 e = (__brev((u32)(k0 + (r << logs))) >> (32 - Cfg::LOGN)) * (col + c);
        const Tw lo = ld_tw(prm.twist_lo + (e & ((1u << prm.twist_shift) - 1u)));
        const Tw hi = ld_tw(prm.twist_hi + (e >> prm.twist_shift));
        x[r][c] = f.mont(f.mont(x[r][c], hi), lo);
      }
  }
}

// forward column pass, stage 0 (kColPre): element (k, column col) of outer block o times the entry of the previous
// pass's matrix that lies at the same place as the element itself: row (o & mask), column k * inner + col
template <class F, class Cfg, int R>
__device__ __forceinline__ void apply_col_pre(const F& f, const PassParams& prm, u64 (&x)[R][Cfg::C], int k0, int logs,
                                              u32 o, u32 col) {
  const Tw* q = prm.pre_twist + (((u64)(o & prm.pre_rows_mask) << prm.pre_shift) + col);
  constexpr int CH = R < 8 ? R : 8;  // rows in flight at a time
#pragma unroll
  for (int r0 = 0; r0 < R; r0 += CH) {
    Tw t[CH][Cfg::C];
#pragma unroll
    for (int r = 0; r < CH; ++r)
#pragma unroll
      for (int c = 0; c < Cfg::C; ++c) t[r][c] = ld_tw_stream(q + (((u64)(u32)(k0 + ((r0 + r) << logs)) << prm.pre_kshift) + c));
#pragma unroll
    for (int r = 0; r < CH; ++r)
#pragma unroll
      for (int c = 0; c < Cfg::C; ++c) x[r0 + r][c] = f.mont(x[r0 + r][c], t[r][c]);
  }
}

// forward row pass, stage 0: residue k of row `row` times entry ((row & mask) << LOGN) + k of the matrix
template <class F, class Cfg, int R>
__device__ __forceinline__ void apply_pre_twist(const F& f, const PassParams& prm, u64 (&x)[R][Cfg::C], int k0, int logs,
                                                u32 row) {
  constexpr int CH = R < 16 ? R : 16;  // entries in flight at a time (4 registers each)
#pragma unroll
  for (int c = 0; c < Cfg::C; ++c) {
    const Tw* q = prm.pre_twist + ((u64)((row + c) & prm.pre_rows_mask) << Cfg::LOGN);  // the mask keeps ragged tiles in range
#pragma unroll
    for (int r0 = 0; r0 < R; r0 += CH) {
      Tw t[CH];
#pragma unroll
      for (int r = 0; r < CH; ++r) t[r] = ld_tw_stream(q + (k0 + ((r0 + r) << logs)));
#pragma unroll
      for (int r = 0; r < CH; ++r) x[r0 + r][c] = f.mont(x[r0 + r][c], t[r]);
    }
  }
}

// XNTT_SHFL_FUSE (measured variant, off in the product build): the innermost exchange of a row pass with one residue per
// slot - the 8 x 8 transposition between the stride-8 and the stride-1 radix-8 stage, which connects 8 consecutive lanes
// - goes through warp shuffles instead of shared memory, and the two stages run back to back on the same registers
// (a radix-64 step: one shared-memory round trip and one __syncwarp less, 24 SHFL + their selects more).  Measured on
// B200 (tools/gpu_variants.py, profiles/r4_variants_shfl.log).
#ifndef XNTT_SHFL_FUSE
#define XNTT_SHFL_FUSE 0
#endif
template <class Cfg>
__host__ __device__ constexpr bool shfl_fused() {
#if XNTT_SHFL_FUSE && !defined(XNTT_HOST_EMU)
  return !Cfg::COL && Cfg::C == 1 && Cfg::LOGRN == 3 && Cfg::NS >= 3 && !Cfg::TMA;
#else
  return false;
#endif
}
#if XNTT_SHFL_FUSE && !defined(XNTT_HOST_EMU)
// new x[r] of lane j = old x[j] of lane r, over every aligned group of 8 lanes: three rounds, each swapping one bit of
// the register index with one bit of the lane index
__device__ __forceinline__ void shfl_transpose8(u64 (&x)[8][1]) {
  const unsigned lane = threadIdx.x;
#pragma unroll
  for (int s = 1; s < 8; s <<= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r & s) continue;
      const u64 send = up ? x[r][0] : x[r | s][0];
      const u64 recv = __shfl_xor_sync(0xffffffffu, send, s);
      if (up)
        x[r][0] = recv;
      else
        x[r | s][0] = recv;
    }
  }
}
#else
__device__ __forceinline__ void shfl_transpose8(u64 (&)[8][1]) {}
#endif

// ---------------------------------------------------------------------------------------------
// Forward radix-R register network (Cooley-Tukey, block-indexed twiddles).
template <class F, int LOGR, int C, bool FIRST>
__device__ __forceinline__ void fwd_network(const F& f, u64 (&x)[1 << LOGR][C], const Tw* __restrict__ G, int B) {
  constexpr int R = 1 << LOGR;
#pragma unroll
  for (int lam = 0; lam < LOGR; ++lam) {
    const int h = R >> (lam + 1);
#pragma unroll
    for (int g = 0; g < (1 << lam); ++g) {
      if (FIRST && g == 0) {
        // omega = 1.  Level 0 sees canonical input; deeper levels must canonicalise x1 first.
#pragma unroll
        for (int r0 = 0; r0 < h; ++r0)
#pragma unroll
          for (int c = 0; c < C; ++c) {
            if (lam > 0) x[r0 + h][c] = f.canon(x[r0 + h][c]);
            f.ct_butterfly_one(x[r0][c], x[r0 + h][c]);
          }
      } else {
        const Tw t = ld_tw(G + ((B << lam) + g));
#pragma unroll
        for (int r0 = 0; r0 < h; ++r0)
#pragma unroll
          for (int c = 0; c < C; ++c) f.ct_butterfly(x[g * 2 * h + r0][c], x[g * 2 * h + r0 + h][c], t);
      }
    }
  }
}

// Inverse radix-R register network (decimation in time, position-indexed twiddles).
// Level lam pairs (r, r + 2^lam); element stride is S, task offset inside its block is i.
template <class F, int LOGR, int C, bool FIRST>
__device__ __forceinline__ void inv_network(const F& f, u64 (&x)[1 << LOGR][C], const Tw* __restrict__ I, int logs,
                                            int i) {
  constexpr int R = 1 << LOGR;
#pragma unroll
  for (int lam = 0; lam < LOGR; ++lam) {
    const int h = 1 << lam;
#pragma unroll
    for (int j = 0; j < h; ++j) {
      if (FIRST && j == 0) {
        // first stage has S = 1, i = 0: position 0 -> omega = 1
#pragma unroll
        for (int r = j; r < R; r += 2 * h)
#pragma unroll
          for (int c = 0; c < C; ++c) {
            if (lam > 0) x[r + h][c] = f.canon(x[r + h][c]);
            f.ct_butterfly_one(x[r][c], x[r + h][c]);
          }
      } else {
        const Tw t = ld_tw(I + ((h << logs) + i + (j << logs)));
#pragma unroll
        for (int r = j; r < R; r += 2 * h)
#pragma unroll
          for (int c = 0; c < C; ++c) f.ct_butterfly(x[r][c], x[r + h][c], t);
      }
    }
  }
}

// Barrier between two adjacent stages of the same radix R.  Their exchange is closed over blocks of R * 2^LOGS consecutive k
// (LOGS = the smaller of the two strides), and in both stages the tasks of one block sit on the same GT
// consecutive threads (GT = 2^LOGS, times the column groups in column mode) in the same loop iteration - so only
// those threads have to meet: a warp-level barrier for the innermost exchange, a named barrier for the next one.
#ifndef XNTT_GROUP_BARRIERS
#define XNTT_GROUP_BARRIERS 1
#endif
// number of consecutive threads that have to meet after stage J (kThreads = the whole CTA)
template <class Cfg, bool INVERSE, int J>
__host__ __device__ constexpr int barrier_group() {
  constexpr int NS = Cfg::NS, LR = Cfg::LOGRN;
  constexpr int NTASKR = (1 << (Cfg::LOGN - LR)) * Cfg::NP;  // tasks of a stage of the common radix
  // stride of the finer of the two stages around the exchange
  constexpr int LOGS = INVERSE ? LR * (J + 1) : LR * (NS - 1 - J);
  // both stages of the common radix: the first forward / last inverse stage has radix 2^LOGR1
  constexpr bool same = (INVERSE ? (J + 1 < NS - 1 || Cfg::LOGR1 == LR) : (J >= 1 || Cfg::LOGR1 == LR)) && NTASKR >= kThreads;
  if (!XNTT_GROUP_BARRIERS || !same || LOGS > 8) return kThreads;
  const int gt = (1 << LOGS) * (Cfg::COL ? Cfg::NP : 1);
  return gt < kThreads ? gt : kThreads;
}

template <int GT>
__device__ __forceinline__ void stage_barrier() {
#if !defined(XNTT_HOST_EMU)
  if constexpr (GT <= 32) {
    __syncwarp();
  } else if constexpr (GT < kThreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)threadIdx.x / GT), "n"(GT) : "memory");
  } else {
    __syncthreads();
  }
#endif
}

// ---------------------------------------------------------------------------------------------
// what the last forward stage does with its R results (stride 2^logs) before they leave the tile
template <class F, class Cfg, int TWIST, int R>
__device__ __forceinline__ void fwd_finish(const F& f, const PassParams& prm, u64 (&x)[R][Cfg::C], u64* gdst, u32 col0,
                                           u32 row0, int k0, int logs, int p) {
  if constexpr (TWIST == kCompactTwist || TWIST == kFullTwist) {
    apply_twist<F, Cfg, R, TWIST>(f, prm, x, k0, logs, col0 + p * Cfg::C);
  } else if constexpr (TWIST == kPointwise || TWIST == kPrePointwise) {
    // fused point-wise product of a polynomial multiply
    // (examples/magic-series/gaussian-polynomial.hpp:201-212); the Montgomery product is canonical
    u64 b[R][Cfg::C];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < Cfg::C; ++c) {
        // rows past the end of a ragged last tile hold nothing to multiply with
        const bool ok = Cfg::COL || row0 + (u32)(p * Cfg::C + c) < prm.rows;
        b[r][c] = ok ? (prm.pointwise + (gdst - prm.dst))[gofs<Cfg>(prm, prm.dmap, k0 + (r << logs), p, c)] : 0ull;
      }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < Cfg::C; ++c) x[r][c] = f.mont(x[r][c], b[r][c], f.companion(b[r][c]));
  } else if constexpr (Cfg::COL && (TWIST == kNoTwist || TWIST == kColPre)) {
    // A forward column pass that leaves its six-step twiddle to the pass behind it stores LAZY residues: that pass
    // (row pass with kPreTwist / kPrePointwise, column pass with kColPre) begins with the Montgomery product by the
    // matrix entry, which takes any 64-bit value and returns a canonical one - canonicalising here would be six
    // instructions per residue for nothing (2^11 columns of a 2^24 plan: 3.6 % of the pass).  Forward column passes
    // never produce a plan's final output (the last pass is a row pass).
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < Cfg::C; ++c) x[r][c] = f.canon(x[r][c]);
  }
  gmem_store<Cfg, R>(prm, gdst, row0, k0, logs, p, x);
}

template <class F, class Cfg, int TWIST, int J>
__device__ __forceinline__ void fwd_stage(const PassParams& prm, typename Slot<Cfg::C>::type* sm,
                                          const u64* gsrc, u64* gdst, u32 col0, u32 row0) {
  const F f = make_field<F>(prm.field);
  constexpr int NS = Cfg::NS;
  constexpr bool FUSE = shfl_fused<Cfg>();       // stages NS-2 and NS-1 back to back, exchange by warp shuffles
  if constexpr (FUSE && J == NS - 1) return;     // ran inside stage NS-2
  constexpr int LOGR = (J == 0) ? Cfg::LOGR1 : Cfg::LOGRN;
  constexpr int R = 1 << LOGR;
  constexpr int LOGS = Cfg::LOGRN * (NS - 1 - J);
  constexpr int LOGT = Cfg::LOGN - LOGR;  // tasks per sub-transform
  constexpr int NTASK = (1 << LOGT) * Cfg::NP;
#pragma unroll 1
  for (int task = threadIdx.x; task < NTASK; task += kThreads) {
    int p, t;
    if constexpr (Cfg::COL) {
      p = task & (Cfg::NP - 1);
      t = task >> Cfg::LOGNP;
    } else {
      p = task >> LOGT;
      t = task & ((1 << LOGT) - 1);
    }
    const int B = t >> LOGS, i = t & ((1 << LOGS) - 1);
    const int k0 = (B << (LOGS + LOGR)) + i;
    u64 x[R][Cfg::C];
    if constexpr (J == 0) {
#if XNTT_TMA_ROWS && !defined(XNTT_HOST_EMU)
      if constexpr (Cfg::TMA) {
#pragma unroll
        for (int r = 0; r < R; ++r) x[r][0] = sm[tma_slot(k0 + (r << LOGS))];
        __syncwarp();  // stage 0 rewrites the tile in place, in the kernel's own swizzle (see tma_slot)
      } else
#endif
        gmem_load<Cfg, R>(prm, gsrc, row0, k0, LOGS, p, x);
      if constexpr (TWIST == kPreTwist || TWIST == kPrePointwise)
        apply_pre_twist<F, Cfg, R>(f, prm, x, k0, LOGS, row0 + (u32)(p * Cfg::C));
      if constexpr (TWIST == kColPre) apply_col_pre<F, Cfg, R>(f, prm, x, k0, LOGS, row0, col0 + p * Cfg::C);
    } else {
      smem_load<Cfg, LOGS, R>(sm, k0, p, x);
    }
    fwd_network<F, LOGR, Cfg::C, J == 0>(f, x, prm.tw, B);
    if constexpr (FUSE && J == NS - 2) {
      // task t of the last stage owns the 8 consecutive residues 8 t .. 8 t + 7: exactly what the 8 lanes of this
      // task's group hold between them
      if constexpr (R == 8 && Cfg::C == 1) {
        shfl_transpose8(x);
        fwd_network<F, 3, Cfg::C, false>(f, x, prm.tw, t);
        fwd_finish<F, Cfg, TWIST, R>(f, prm, x, gdst, col0, row0, t << 3, 0, p);
      }
    } else if constexpr (J == NS - 1) {
      fwd_finish<F, Cfg, TWIST, R>(f, prm, x, gdst, col0, row0, k0, LOGS, p);
    } else {
      smem_store<Cfg, LOGS, R>(sm, k0, p, x);
    }
  }
  if constexpr (J != NS - 1 && !(FUSE && J == NS - 2)) stage_barrier<barrier_group<Cfg, false, J>()>();
}

template <class F, class Cfg, int TWIST, int J>
__device__ __forceinline__ void inv_stage(const PassParams& prm, typename Slot<Cfg::C>::type* sm,
                                          const u64* gsrc, u64* gdst, u32 col0, u32 row0) {
  const F f = make_field<F>(prm.field);
  constexpr int NS = Cfg::NS;
  constexpr bool FUSE = shfl_fused<Cfg>();  // stages 0 and 1 back to back, exchange by warp shuffles
  if constexpr (FUSE && J == 1) {
    stage_barrier<barrier_group<Cfg, true, 1>()>();  // ran inside stage 0; its results are in shared memory
    return;
  }
  // inverse stage J mirrors forward stage NS-1-J
  constexpr int LOGR = (J == NS - 1) ? Cfg::LOGR1 : Cfg::LOGRN;
  constexpr int R = 1 << LOGR;
  constexpr int LOGS = Cfg::LOGRN * J;
  constexpr int LOGT = Cfg::LOGN - LOGR;
  constexpr int NTASK = (1 << LOGT) * Cfg::NP;
#pragma unroll 1
  for (int task = threadIdx.x; task < NTASK; task += kThreads) {
    int p, t;
    if constexpr (Cfg::COL) {
      p = task & (Cfg::NP - 1);
      t = task >> Cfg::LOGNP;
    } else {
      p = task >> LOGT;
      t = task & ((1 << LOGT) - 1);
    }
    const int B = t >> LOGS, i = t & ((1 << LOGS) - 1);
    const int k0 = (B << (LOGS + LOGR)) + i;
    u64 x[R][Cfg::C];
    if constexpr (J == 0) {
#if XNTT_TMA_ROWS && !defined(XNTT_HOST_EMU)
      if constexpr (Cfg::TMA) {
        // 8 consecutive residues per task: four 16-byte reads from the TMA-swizzled tile
#pragma unroll
        for (int r = 0; r < R; r += 2) {
          const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(&sm[tma_slot(k0 + r)]);
          x[r][0] = v.x;
          x[r + 1][0] = v.y;
        }
        __syncwarp();
      } else
#endif
        gmem_load<Cfg, R>(prm, gsrc, row0, k0, LOGS, p, x);
      if constexpr (TWIST == kCompactTwist || TWIST == kFullTwist)
        apply_twist<F, Cfg, R, TWIST>(f, prm, x, k0, LOGS, col0 + p * Cfg::C);
    } else {
      smem_load<Cfg, LOGS, R>(sm, k0, p, x);
    }
    inv_network<F, LOGR, Cfg::C, J == 0>(f, x, prm.tw, LOGS, i);
    if constexpr (FUSE && J == 0) {
      // stage 1 (stride 8): task t owns residues 64 (t >> 3) + (t & 7) + 8 r - one from each lane of its group of 8
      if constexpr (R == 8 && Cfg::C == 1) {
        shfl_transpose8(x);
        inv_network<F, 3, Cfg::C, false>(f, x, prm.tw, 3, t & 7);
        smem_store<Cfg, 3, R>(sm, ((t >> 3) << 6) + (t & 7), p, x);
      }
    } else if constexpr (J == NS - 1 && TWIST == kPostTwist) {
      // the twiddle matrix of the column pass behind this row pass (and 1/inverse_factor with it); canonical result
      apply_pre_twist<F, Cfg, R>(f, prm, x, k0, LOGS, row0 + (u32)(p * Cfg::C));
      gmem_store<Cfg, R>(prm, gdst, row0, k0, LOGS, p, x);
    } else if constexpr (J == NS - 1) {
      if (prm.scale_on) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int c = 0; c < Cfg::C; ++c) x[r][c] = f.mont(x[r][c], prm.scale);
      } else {
        // an inner column pass of a three-pass plan may store lazy residues (PassParams::lazy_out)
        const bool on = !(Cfg::COL && prm.lazy_out != 0);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int c = 0; c < Cfg::C; ++c) x[r][c] = Cfg::COL ? f.canon_if(x[r][c], on) : f.canon(x[r][c]);
      }
      gmem_store<Cfg, R>(prm, gdst, row0, k0, LOGS, p, x);
    } else {
      smem_store<Cfg, LOGS, R>(sm, k0, p, x);
    }
  }
  if constexpr (J != NS - 1 && !(FUSE && J == 0)) stage_barrier<barrier_group<Cfg, true, J>()>();
}

template <class F, class Cfg, bool INVERSE, int TWIST, int... Js>
__device__ __forceinline__ void run_stages(const PassParams& prm, typename Slot<Cfg::C>::type* sm,
                                           const u64* gsrc, u64* gdst, u32 col0, u32 row0,
                                           std::integer_sequence<int, Js...>) {
  if constexpr (INVERSE)
    (inv_stage<F, Cfg, TWIST, Js>(prm, sm, gsrc, gdst, col0, row0), ...);
  else
    (fwd_stage<F, Cfg, TWIST, Js>(prm, sm, gsrc, gdst, col0, row0), ...);
}

// Where tile number `tile` starts in src and dst, and which global column / row it begins with.
template <class Cfg>
__device__ __forceinline__ void tile_origin(const PassParams& prm, u32 tile, u64& sbase, u64& dbase, u32& col0,
                                            u32& row0) {
  if constexpr (Cfg::COL) {
    const u32 o = tile / prm.tiles_per_outer, cb = tile - o * prm.tiles_per_outer;
    if constexpr (Cfg::MAP) {
      sbase = (u64)o * prm.smap.outer + (u64)cb * Cfg::W;
      dbase = (u64)o * prm.dmap.outer + (u64)cb * Cfg::W;
    } else {
      sbase = dbase = (u64)o * prm.outer_stride + (u64)cb * Cfg::W;
    }
    col0 = prm.twist_col0 + cb * Cfg::W;
    row0 = o;  // column mode: the outer block (kColPre indexes the previous pass's matrix with it)
  } else {
    row0 = tile << Cfg::LOGW;
    if constexpr (Cfg::MAP) {
      sbase = (u64)row0 * prm.smap.outer;
      dbase = (u64)row0 * prm.dmap.outer;
    } else {
      sbase = dbase = ((u64)tile << (Cfg::LOGN + Cfg::LOGW));
    }
  }
}

#if !defined(XNTT_HOST_EMU)
template <class F, int LOGN, int LOGW, int C, bool COL, bool INVERSE, int TWIST, bool MAP = false>
__global__ void __launch_bounds__(kThreads, pass_minb(COL, C, TWIST == kNoTwist)) pass_kernel(const __grid_constant__ PassParams prm) {
  typedef PassCfg<LOGN, LOGW, C, COL, MAP> Cfg;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  auto* sm = reinterpret_cast<typename Slot<C>::type*>(smem_raw);
  const u32 tile = blockIdx.x;
  u64 sbase, dbase;
  u32 col0 = 0, row0 = 0;
  tile_origin<Cfg>(prm, tile, sbase, dbase, col0, row0);
  // Programmatic dependent launch (dispatch.cuh launches every pass with programmatic stream serialisation): the next
  // kernel in the stream may become resident now, while this grid is still running - it blocks in its own
  // griddepcontrol.wait below until this grid has completed and its stores are visible.  What a pass of a small plan
  // saves is the launch latency between two dependent kernels.  Nothing above or below this point and before the wait
  // touches data another kernel writes (tables and twiddle matrices are read-only once the plan exists).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if constexpr (TWIST == kPostTwist) {
    // consumed by the last stage: start pulling this tile's rows of the matrix (16 N W bytes, contiguous per row)
    // into L2 now
    for (int e = threadIdx.x * 8; e < Cfg::N * Cfg::W; e += kThreads * 8) {  // one 128-byte line = 8 entries
      const u32 row = row0 + (u32)(e >> Cfg::LOGN);
      prefetch_l2(prm.pre_twist + (((u64)(row & prm.pre_rows_mask) << Cfg::LOGN) + (u64)(e & (Cfg::N - 1))));
    }
  }
  if constexpr (TWIST == kFullTwist && !INVERSE) {
    // forward: the twiddle matrix is consumed by the last stage - start pulling this tile's N segments of it
    // (W entries = 16 W bytes each) into L2 now, so that the streamed loads there do not wait for HBM
    {
      const Tw* q = prm.twist_full + col0;
      constexpr int SEG = Cfg::W >= 2 ? Cfg::W / 2 : 1;  // 32-byte sectors per segment
      for (int i = threadIdx.x; i < Cfg::N * SEG; i += kThreads)
        prefetch_l2(q + (((u64)(i / SEG) << prm.twist_full_shift) + (u64)(i % SEG) * 2));
    }
  }
#if XNTT_PREFETCH_TABLE
  // the pass's own twiddle table (read-only) into L1 while waiting for the kernel before: one 128-byte line per thread step
  for (int e = threadIdx.x * 8; e < (INVERSE ? Cfg::N : Cfg::N / 2); e += kThreads * 8)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(prm.tw + e));
#endif
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the kernel before this one has completed
  run_stages<F, Cfg, INVERSE, TWIST>(prm, sm, prm.src + sbase, prm.dst + dbase, col0, row0,
                                     std::make_integer_sequence<int, Cfg::NS>{});
}
#if XNTT_TMA_ROWS
// Row pass of length 2^13 (one row per tile) with the tile staged by the TMA unit.  Same stages, same tables, same
// results as pass_kernel<F, 13, 0, 1, false, ...>; only where stage 0 takes its input from differs.
template <class F, bool INVERSE, int TWIST>
__global__ void __launch_bounds__(kThreads, XNTT_MINB)
    row_kernel_tma(const __grid_constant__ PassParams prm, const __grid_constant__ CUtensorMap tmap) {
  typedef PassCfg<13, 0, 1, false, false, true> Cfg;
  // the 128-byte swizzle is a function of the shared-memory address: the tile has to start on a 1024-byte boundary,
  // which the launch only guarantees if we round up ourselves (1 KiB of slack is allocated for it); the mbarrier
  // sits behind the tile
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem_raw = smem_dyn + ((1024u - ((unsigned)__cvta_generic_to_shared(smem_dyn) & 1023u)) & 1023u);
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(smem_raw + 65536);
  auto* sm = reinterpret_cast<typename Slot<1>::type*>(smem_raw);
  const u32 tile = blockIdx.x;
  u64 sbase, dbase;
  u32 col0 = 0, row0 = 0;
  tile_origin<Cfg>(prm, tile, sbase, dbase, col0, row0);
  if constexpr (TWIST == kPostTwist) {
    for (int e = threadIdx.x * 8; e < Cfg::N * Cfg::W; e += kThreads * 8) {
      const u32 row = row0 + (u32)(e >> Cfg::LOGN);
      prefetch_l2(prm.pre_twist + (((u64)(row & prm.pre_rows_mask) << Cfg::LOGN) + (u64)(e & (Cfg::N - 1))));
    }
  }
  row_tile_tma_load(smem_raw, mbar, &tmap, tile);
  run_stages<F, Cfg, INVERSE, TWIST>(prm, sm, prm.src + sbase, prm.dst + dbase, col0, row0,
                                     std::make_integer_sequence<int, Cfg::NS>{});
}
#endif
#endif  // !XNTT_HOST_EMU

}  // namespace xntt
