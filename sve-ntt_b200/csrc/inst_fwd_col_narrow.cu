// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: fwd_col, narrow tiles, field F0.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_fwd_col_narrow(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE_NARROW(F0, 1, true, false)
    XNTT_CASE_NARROW(F0, 2, true, false)
    XNTT_CASE_NARROW(F0, 3, true, false)
    XNTT_CASE_NARROW(F0, 4, true, false)
    XNTT_CASE_NARROW(F0, 5, true, false)
    XNTT_CASE_NARROW(F0, 6, true, false)
    XNTT_CASE_NARROW(F0, 7, true, false)
    XNTT_CASE_NARROW(F0, 8, true, false)
    XNTT_CASE_NARROW(F0, 9, true, false)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
