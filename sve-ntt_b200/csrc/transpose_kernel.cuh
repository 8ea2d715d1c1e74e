// SPDX-License-Identifier: Apache-2.0
//
// Stand-alone tiled transposition with the contract of the reference's 17 Transpose* classes
// (include/sventt/transposition/sve/*.hpp, e.g. in-register.hpp:16-109 and
// in-register-explicit-blocking-row-first.hpp:28-124):
//     out of place : dst[ld_dst * c + r] = src[ld_src * r + c]   for r < rows, c < cols
//     in place     : square, dst[ld * c + r] <-> dst[ld * r + c]
// One CTA moves a 64 x 64 tile of u64 through padded shared memory: coalesced 512-byte row segments on
// both the load and the store side.  (Inside the transforms no transposition runs at all - the column
// pass reads its strided tile directly; this kernel exists for callers that use the classes on their own,
// as tests/bench-transpose.cpp does.)
#pragma once
#include "params.h"

namespace xntt {

constexpr int kTrTile = 64;
constexpr int kTrPad = 1;
constexpr int kTrThreads = 256;
constexpr size_t kTrSmemWords = (size_t)kTrTile * (kTrTile + kTrPad);

struct TransposeParams {
  u64* dst;
  const u64* src;
  u64 rows, cols, ld_dst, ld_src;
  u32 tiles_c;  // tiles along the column direction of src
};

// phase 1: src tile (r0.., c0..) -> smem[r][c]
__device__ __forceinline__ void tr_load(const u64* __restrict__ src, u64 ld, u64 rows, u64 cols, u64 r0, u64 c0,
                                        u64* sm, int tid) {
  const int tx = tid & (kTrTile - 1), ty = tid >> 6;  // 64 columns x 4 rows per sweep
  // all 16 loads of a thread are issued before the first value is needed (memory-level parallelism)
  constexpr int kSweeps = kTrTile / (kTrThreads / kTrTile);
  u64 v[kSweeps];
#pragma unroll
  for (int s = 0; s < kSweeps; ++s) {
    const u64 r = r0 + ty + s * (kTrThreads / kTrTile), c = c0 + tx;
    v[s] = (r < rows && c < cols) ? src[ld * r + c] : 0ull;
  }
#pragma unroll
  for (int s = 0; s < kSweeps; ++s) sm[(ty + s * (kTrThreads / kTrTile)) * (kTrTile + kTrPad) + tx] = v[s];
}
// phase 2: smem[r][c] -> dst tile at (c0.., r0..), i.e. transposed
__device__ __forceinline__ void tr_store(u64* __restrict__ dst, u64 ld, u64 rows, u64 cols, u64 r0, u64 c0,
                                         const u64* sm, int tid) {
  const int tx = tid & (kTrTile - 1), ty = tid >> 6;
#pragma unroll
  for (int i = ty; i < kTrTile; i += kTrThreads / kTrTile) {
    const u64 c = c0 + i, r = r0 + tx;  // dst row = src column
    if (r < rows && c < cols) dst[ld * c + r] = sm[tx * (kTrTile + kTrPad) + i];
  }
}

#if !defined(XNTT_HOST_EMU)
__global__ void __launch_bounds__(kTrThreads) transpose_kernel(const __grid_constant__ TransposeParams p) {
  extern __shared__ __align__(16) unsigned char tr_smem_raw[];
  u64* sm = reinterpret_cast<u64*>(tr_smem_raw);
  const u64 tr = blockIdx.x / p.tiles_c, tc = blockIdx.x % p.tiles_c;
  const u64 r0 = tr * kTrTile, c0 = tc * kTrTile;
  tr_load(p.src, p.ld_src, p.rows, p.cols, r0, c0, sm, threadIdx.x);
  __syncthreads();
  tr_store(p.dst, p.ld_dst, p.rows, p.cols, r0, c0, sm, threadIdx.x);
}

// in place, square: one CTA per unordered tile pair {(i, j), (j, i)}, i <= j - a 1-D grid over the upper triangle
// (row i of the triangle starts at i * T - i (i - 1) / 2); the loads of both tiles are in flight together
__global__ void __launch_bounds__(kTrThreads) transpose_inplace_kernel(const __grid_constant__ TransposeParams p) {
  extern __shared__ __align__(16) unsigned char tr_smem_raw[];
  u64* sa = reinterpret_cast<u64*>(tr_smem_raw);
  u64* sb = sa + kTrSmemWords;
  const u64 T = p.tiles_c, idx = blockIdx.x;
  // largest i with i * T - i (i - 1) / 2 <= idx
  u64 i = (u64)(((double)(2 * T + 1) - sqrt((double)(2 * T + 1) * (double)(2 * T + 1) - 8.0 * (double)idx)) * 0.5);
  while (i > 0 && i * T - i * (i - 1) / 2 > idx) --i;
  while ((i + 1) * T - (i + 1) * i / 2 <= idx) ++i;
  const u64 j = i + (idx - (i * T - i * (i - 1) / 2));
  const u64 r0 = i * kTrTile, c0 = j * kTrTile;
  {
    const int tx = threadIdx.x & (kTrTile - 1), ty = threadIdx.x >> 6;
    constexpr int kStep = kTrThreads / kTrTile, kSweeps = kTrTile / kStep;
    u64 va[kSweeps], vb[kSweeps];
#pragma unroll
    for (int s = 0; s < kSweeps; ++s) {
      const u64 ra = r0 + ty + s * kStep, ca = c0 + tx, rb = c0 + ty + s * kStep, cb = r0 + tx;
      va[s] = (ra < p.rows && ca < p.cols) ? p.dst[p.ld_dst * ra + ca] : 0ull;
      vb[s] = (i != j && rb < p.rows && cb < p.cols) ? p.dst[p.ld_dst * rb + cb] : 0ull;
    }
#pragma unroll
    for (int s = 0; s < kSweeps; ++s) {
      sa[(ty + s * kStep) * (kTrTile + kTrPad) + tx] = va[s];
      sb[(ty + s * kStep) * (kTrTile + kTrPad) + tx] = vb[s];
    }
  }
  __syncthreads();
  tr_store(p.dst, p.ld_dst, p.rows, p.cols, r0, c0, sa, threadIdx.x);
  if (i != j) tr_store(p.dst, p.ld_dst, p.rows, p.cols, c0, r0, sb, threadIdx.x);
}
#endif

}  // namespace xntt
