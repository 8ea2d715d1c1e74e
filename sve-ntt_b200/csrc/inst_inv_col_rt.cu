// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: inv_col, field FieldRT.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_inv_col_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE(FieldRT, 1, true, true)
    XNTT_CASE(FieldRT, 2, true, true)
    XNTT_CASE(FieldRT, 3, true, true)
    XNTT_CASE(FieldRT, 4, true, true)
    XNTT_CASE(FieldRT, 5, true, true)
    XNTT_CASE(FieldRT, 6, true, true)
    XNTT_CASE(FieldRT, 7, true, true)
    XNTT_CASE(FieldRT, 8, true, true)
    XNTT_CASE(FieldRT, 9, true, true)
    XNTT_CASE(FieldRT, 10, true, true)
    XNTT_CASE(FieldRT, 11, true, true)
    XNTT_CASE(FieldRT, 12, true, true)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
