// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: fwd_col, field FieldRT.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_fwd_col_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE(FieldRT, 1, true, false)
    XNTT_CASE(FieldRT, 2, true, false)
    XNTT_CASE(FieldRT, 3, true, false)
    XNTT_CASE(FieldRT, 4, true, false)
    XNTT_CASE(FieldRT, 5, true, false)
    XNTT_CASE(FieldRT, 6, true, false)
    XNTT_CASE(FieldRT, 7, true, false)
    XNTT_CASE(FieldRT, 8, true, false)
    XNTT_CASE(FieldRT, 9, true, false)
    XNTT_CASE(FieldRT, 10, true, false)
    XNTT_CASE(FieldRT, 11, true, false)
    XNTT_CASE(FieldRT, 12, true, false)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
