// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: inv_row, field FieldShoup, generalised address maps.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_inv_row_map_sh(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE_MAP(FieldShoup, 1, false, true)
    XNTT_CASE_MAP(FieldShoup, 2, false, true)
    XNTT_CASE_MAP(FieldShoup, 3, false, true)
    XNTT_CASE_MAP(FieldShoup, 4, false, true)
    XNTT_CASE_MAP(FieldShoup, 5, false, true)
    XNTT_CASE_MAP(FieldShoup, 6, false, true)
    XNTT_CASE_MAP(FieldShoup, 7, false, true)
    XNTT_CASE_MAP(FieldShoup, 8, false, true)
    XNTT_CASE_MAP(FieldShoup, 9, false, true)
    XNTT_CASE_MAP(FieldShoup, 10, false, true)
    XNTT_CASE_MAP(FieldShoup, 11, false, true)
    XNTT_CASE_MAP(FieldShoup, 12, false, true)
    XNTT_CASE_MAP(FieldShoup, 13, false, true)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
