// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: fwd_row, narrow tiles, field F0.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_fwd_row_narrow(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE_NARROW(F0, 1, false, false)
    XNTT_CASE_NARROW(F0, 2, false, false)
    XNTT_CASE_NARROW(F0, 3, false, false)
    XNTT_CASE_NARROW(F0, 4, false, false)
    XNTT_CASE_NARROW(F0, 5, false, false)
    XNTT_CASE_NARROW(F0, 6, false, false)
    XNTT_CASE_NARROW(F0, 7, false, false)
    XNTT_CASE_NARROW(F0, 8, false, false)
    XNTT_CASE_NARROW(F0, 9, false, false)
    XNTT_CASE_NARROW(F0, 10, false, false)
    XNTT_CASE_NARROW(F0, 11, false, false)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
