// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: inv_row, field FieldRT.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_inv_row_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE(FieldRT, 1, false, true)
    XNTT_CASE(FieldRT, 2, false, true)
    XNTT_CASE(FieldRT, 3, false, true)
    XNTT_CASE(FieldRT, 4, false, true)
    XNTT_CASE(FieldRT, 5, false, true)
    XNTT_CASE(FieldRT, 6, false, true)
    XNTT_CASE(FieldRT, 7, false, true)
    XNTT_CASE(FieldRT, 8, false, true)
    XNTT_CASE(FieldRT, 9, false, true)
    XNTT_CASE(FieldRT, 10, false, true)
    XNTT_CASE(FieldRT, 11, false, true)
    XNTT_CASE(FieldRT, 12, false, true)
    XNTT_CASE(FieldRT, 13, false, true)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
