// SPDX-License-Identifier: Apache-2.0
// Launch-side view of the pass kernels: tile shapes per pass length and the four dispatchers
// (one translation unit each so the heavy template instantiations compile in parallel).
#pragma once
#include "pass_kernel.cuh"

namespace xntt {

typedef Field<kP0> F0;

cudaError_t launch_fwd_row(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_row(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_fwd_col(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_col(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);

template <int LOGN, bool COL, bool INV>
cudaError_t launch_one(const PassParams& prm, unsigned grid, cudaStream_t st) {
  constexpr int LOGW = tile_logw(LOGN), C = tile_c(LOGN);
  typedef PassCfg<LOGN, LOGW, C, COL> Cfg;
  auto kern = pass_kernel<F0, LOGN, LOGW, C, COL, INV, COL>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  kern<<<grid, kThreads, Cfg::kSmemBytes, st>>>(prm);
  return cudaGetLastError();
}

#define XNTT_CASE(L, COL, INV) \
  case L:                      \
    return launch_one<L, COL, INV>(prm, grid, st);

}  // namespace xntt
