// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: fwd_row, field FieldShoup.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_fwd_row_sh(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE(FieldShoup, 1, false, false)
    XNTT_CASE(FieldShoup, 2, false, false)
    XNTT_CASE(FieldShoup, 3, false, false)
    XNTT_CASE(FieldShoup, 4, false, false)
    XNTT_CASE(FieldShoup, 5, false, false)
    XNTT_CASE(FieldShoup, 6, false, false)
    XNTT_CASE(FieldShoup, 7, false, false)
    XNTT_CASE(FieldShoup, 8, false, false)
    XNTT_CASE(FieldShoup, 9, false, false)
    XNTT_CASE(FieldShoup, 10, false, false)
    XNTT_CASE(FieldShoup, 11, false, false)
    XNTT_CASE(FieldShoup, 12, false, false)
    XNTT_CASE(FieldShoup, 13, false, false)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
