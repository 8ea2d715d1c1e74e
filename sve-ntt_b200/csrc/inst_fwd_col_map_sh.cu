// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: fwd_col, field FieldShoup, generalised address maps.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_fwd_col_map_sh(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE_MAP(FieldShoup, 1, true, false)
    XNTT_CASE_MAP(FieldShoup, 2, true, false)
    XNTT_CASE_MAP(FieldShoup, 3, true, false)
    XNTT_CASE_MAP(FieldShoup, 4, true, false)
    XNTT_CASE_MAP(FieldShoup, 5, true, false)
    XNTT_CASE_MAP(FieldShoup, 6, true, false)
    XNTT_CASE_MAP(FieldShoup, 7, true, false)
    XNTT_CASE_MAP(FieldShoup, 8, true, false)
    XNTT_CASE_MAP(FieldShoup, 9, true, false)
    XNTT_CASE_MAP(FieldShoup, 10, true, false)
    XNTT_CASE_MAP(FieldShoup, 11, true, false)
    XNTT_CASE_MAP(FieldShoup, 12, true, false)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
