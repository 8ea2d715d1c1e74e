// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: fwd_row, narrow tiles, field FieldRT.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_fwd_row_narrow_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE_NARROW(FieldRT, 1, false, false)
    XNTT_CASE_NARROW(FieldRT, 2, false, false)
    XNTT_CASE_NARROW(FieldRT, 3, false, false)
    XNTT_CASE_NARROW(FieldRT, 4, false, false)
    XNTT_CASE_NARROW(FieldRT, 5, false, false)
    XNTT_CASE_NARROW(FieldRT, 6, false, false)
    XNTT_CASE_NARROW(FieldRT, 7, false, false)
    XNTT_CASE_NARROW(FieldRT, 8, false, false)
    XNTT_CASE_NARROW(FieldRT, 9, false, false)
    XNTT_CASE_NARROW(FieldRT, 10, false, false)
    XNTT_CASE_NARROW(FieldRT, 11, false, false)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
