// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: inv_row, narrow tiles, field FieldRT.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_inv_row_narrow_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE_NARROW(FieldRT, 1, false, true)
    XNTT_CASE_NARROW(FieldRT, 2, false, true)
    XNTT_CASE_NARROW(FieldRT, 3, false, true)
    XNTT_CASE_NARROW(FieldRT, 4, false, true)
    XNTT_CASE_NARROW(FieldRT, 5, false, true)
    XNTT_CASE_NARROW(FieldRT, 6, false, true)
    XNTT_CASE_NARROW(FieldRT, 7, false, true)
    XNTT_CASE_NARROW(FieldRT, 8, false, true)
    XNTT_CASE_NARROW(FieldRT, 9, false, true)
    XNTT_CASE_NARROW(FieldRT, 10, false, true)
    XNTT_CASE_NARROW(FieldRT, 11, false, true)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
