// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: inv_row, narrow tiles, field F0.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_inv_row_narrow(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE_NARROW(F0, 1, false, true)
    XNTT_CASE_NARROW(F0, 2, false, true)
    XNTT_CASE_NARROW(F0, 3, false, true)
    XNTT_CASE_NARROW(F0, 4, false, true)
    XNTT_CASE_NARROW(F0, 5, false, true)
    XNTT_CASE_NARROW(F0, 6, false, true)
    XNTT_CASE_NARROW(F0, 7, false, true)
    XNTT_CASE_NARROW(F0, 8, false, true)
    XNTT_CASE_NARROW(F0, 9, false, true)
    XNTT_CASE_NARROW(F0, 10, false, true)
    XNTT_CASE_NARROW(F0, 11, false, true)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
