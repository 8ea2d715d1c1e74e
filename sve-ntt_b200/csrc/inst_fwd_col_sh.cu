// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: fwd_col, field FieldShoup.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_fwd_col_sh(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE(FieldShoup, 1, true, false)
    XNTT_CASE(FieldShoup, 2, true, false)
    XNTT_CASE(FieldShoup, 3, true, false)
    XNTT_CASE(FieldShoup, 4, true, false)
    XNTT_CASE(FieldShoup, 5, true, false)
    XNTT_CASE(FieldShoup, 6, true, false)
    XNTT_CASE(FieldShoup, 7, true, false)
    XNTT_CASE(FieldShoup, 8, true, false)
    XNTT_CASE(FieldShoup, 9, true, false)
    XNTT_CASE(FieldShoup, 10, true, false)
    XNTT_CASE(FieldShoup, 11, true, false)
    XNTT_CASE(FieldShoup, 12, true, false)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
