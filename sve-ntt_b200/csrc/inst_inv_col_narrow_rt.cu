// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: inv_col, narrow tiles, field FieldRT.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_inv_col_narrow_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE_NARROW(FieldRT, 1, true, true)
    XNTT_CASE_NARROW(FieldRT, 2, true, true)
    XNTT_CASE_NARROW(FieldRT, 3, true, true)
    XNTT_CASE_NARROW(FieldRT, 4, true, true)
    XNTT_CASE_NARROW(FieldRT, 5, true, true)
    XNTT_CASE_NARROW(FieldRT, 6, true, true)
    XNTT_CASE_NARROW(FieldRT, 7, true, true)
    XNTT_CASE_NARROW(FieldRT, 8, true, true)
    XNTT_CASE_NARROW(FieldRT, 9, true, true)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
