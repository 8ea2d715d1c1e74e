// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: fwd_col, field F0, generalised address maps.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_fwd_col_map(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE_MAP(F0, 1, true, false)
    XNTT_CASE_MAP(F0, 2, true, false)
    XNTT_CASE_MAP(F0, 3, true, false)
    XNTT_CASE_MAP(F0, 4, true, false)
    XNTT_CASE_MAP(F0, 5, true, false)
    XNTT_CASE_MAP(F0, 6, true, false)
    XNTT_CASE_MAP(F0, 7, true, false)
    XNTT_CASE_MAP(F0, 8, true, false)
    XNTT_CASE_MAP(F0, 9, true, false)
    XNTT_CASE_MAP(F0, 10, true, false)
    XNTT_CASE_MAP(F0, 11, true, false)
    XNTT_CASE_MAP(F0, 12, true, false)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
