// SPDX-License-Identifier: Apache-2.0
// Launch-side view of the pass kernels: the dispatchers (one translation unit per direction / mode
// / field flavour so the heavy template instantiations compile in parallel).
#pragma once
#include <atomic>
#include <mutex>

#include "pass_kernel.cuh"

namespace xntt {

// static-modulus kernels (p = kP0) and runtime-modulus kernels (any odd prime < 2^64)
cudaError_t launch_fwd_row(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_row(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_fwd_col(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_col(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
// kP0 kernels with generalised src/dst address maps (passes next to the all-to-all of a sharded plan)
cudaError_t launch_fwd_row_map(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_row_map(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_fwd_col_map(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_col_map(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_fwd_row_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_row_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_fwd_col_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_col_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);

template <class F, int LOGN, bool COL, bool INV, bool MAP, int TWIST>
cudaError_t launch_kernel(const PassParams& prm, unsigned grid, cudaStream_t st) {
  constexpr int LOGW = tile_logw(LOGN), C = tile_c(LOGN);
  typedef PassCfg<LOGN, LOGW, C, COL> Cfg;
  auto kern = pass_kernel<F, LOGN, LOGW, C, COL, INV, TWIST, MAP>;
  // function attributes live in the context of the device they were set on: once per kernel and device, and
  // safe against two host threads launching the same kernel for the first time (the reference's compute_* are
  // const and re-entrant, wrapper.hpp:50-82)
  static std::atomic<bool> attr_done_on[64];
  static std::mutex attr_mu;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::atomic<bool>& attr_done = attr_done_on[dev & 63];
  if (!attr_done.load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lock(attr_mu);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
#ifndef XNTT_CARVE_TILES
#define XNTT_CARVE_TILES XNTT_MINB
#endif
#if XNTT_CARVE_TILES > 2
    // room for XNTT_MINB resident tiles and not more: what is left of the 228 KiB stays L1, which the twiddle
    // tables live in (measured: 2^24 forward 433 us with the default split, 506 us with L1 squeezed to 32 KiB)
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                             (int)((XNTT_CARVE_TILES * (Cfg::kSmemBytes + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)));
    if (e != cudaSuccess) return e;
#endif
    attr_done.store(true, std::memory_order_release);
  }
  kern<<<grid, kThreads, Cfg::kSmemBytes, st>>>(prm);
  return cudaGetLastError();
}

// one instantiation per kind a pass can take (pass_kernel.cuh: pass_kind).  The generalised-addressing kernels of
// sharded plans serve the pass next to the exchange: a column pass there has the compact form, plus what the planner
// uses for an inner column pass - the whole matrix in the inverse, none at all in the forward direction (the row pass
// applies it) - and a row pass is always plain.
template <class F, int LOGN, bool COL, bool INV, bool MAP = false>
cudaError_t launch_one(const PassParams& prm, unsigned grid, cudaStream_t st) {
  const int kind = pass_kind(COL, INV, MAP, prm);
  if constexpr (COL) {
    if (kind == kCompactTwist) return launch_kernel<F, LOGN, COL, INV, MAP, kCompactTwist>(prm, grid, st);
    if constexpr (INV || !MAP) {
      if (kind == kFullTwist) return launch_kernel<F, LOGN, COL, INV, MAP, kFullTwist>(prm, grid, st);
    }
    // twist-free: the row pass next to it applies the matrix
    if (kind == kNoTwist) return launch_kernel<F, LOGN, COL, INV, MAP, kNoTwist>(prm, grid, st);
    return cudaErrorInvalidValue;
  } else {
    if constexpr (!INV && !MAP) {
      if (kind == kPointwise) return launch_kernel<F, LOGN, COL, INV, MAP, kPointwise>(prm, grid, st);
      if (kind == kPreTwist) return launch_kernel<F, LOGN, COL, INV, MAP, kPreTwist>(prm, grid, st);
      if (kind == kPrePointwise) return launch_kernel<F, LOGN, COL, INV, MAP, kPrePointwise>(prm, grid, st);
    }
    if constexpr (INV && !MAP) {
      if (kind == kPostTwist) return launch_kernel<F, LOGN, COL, INV, MAP, kPostTwist>(prm, grid, st);
    }
    return launch_kernel<F, LOGN, COL, INV, MAP, kNoTwist>(prm, grid, st);
  }
}

#define XNTT_CASE(F, L, COL, INV) \
  case L:                         \
    return launch_one<F, L, COL, INV>(prm, grid, st);
#define XNTT_CASE_MAP(F, L, COL, INV) \
  case L:                             \
    return launch_one<F, L, COL, INV, true>(prm, grid, st);

}  // namespace xntt
