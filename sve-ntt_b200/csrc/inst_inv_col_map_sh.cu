// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: inv_col, field FieldShoup, generalised address maps.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_inv_col_map_sh(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE_MAP(FieldShoup, 1, true, true)
    XNTT_CASE_MAP(FieldShoup, 2, true, true)
    XNTT_CASE_MAP(FieldShoup, 3, true, true)
    XNTT_CASE_MAP(FieldShoup, 4, true, true)
    XNTT_CASE_MAP(FieldShoup, 5, true, true)
    XNTT_CASE_MAP(FieldShoup, 6, true, true)
    XNTT_CASE_MAP(FieldShoup, 7, true, true)
    XNTT_CASE_MAP(FieldShoup, 8, true, true)
    XNTT_CASE_MAP(FieldShoup, 9, true, true)
    XNTT_CASE_MAP(FieldShoup, 10, true, true)
    XNTT_CASE_MAP(FieldShoup, 11, true, true)
    XNTT_CASE_MAP(FieldShoup, 12, true, true)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
