// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: inv_col, field FieldShoup.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_inv_col_sh(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE(FieldShoup, 1, true, true)
    XNTT_CASE(FieldShoup, 2, true, true)
    XNTT_CASE(FieldShoup, 3, true, true)
    XNTT_CASE(FieldShoup, 4, true, true)
    XNTT_CASE(FieldShoup, 5, true, true)
    XNTT_CASE(FieldShoup, 6, true, true)
    XNTT_CASE(FieldShoup, 7, true, true)
    XNTT_CASE(FieldShoup, 8, true, true)
    XNTT_CASE(FieldShoup, 9, true, true)
    XNTT_CASE(FieldShoup, 10, true, true)
    XNTT_CASE(FieldShoup, 11, true, true)
    XNTT_CASE(FieldShoup, 12, true, true)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
