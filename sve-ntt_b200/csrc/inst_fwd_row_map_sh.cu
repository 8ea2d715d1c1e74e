// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: fwd_row, field FieldShoup, generalised address maps.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_fwd_row_map_sh(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE_MAP(FieldShoup, 1, false, false)
    XNTT_CASE_MAP(FieldShoup, 2, false, false)
    XNTT_CASE_MAP(FieldShoup, 3, false, false)
    XNTT_CASE_MAP(FieldShoup, 4, false, false)
    XNTT_CASE_MAP(FieldShoup, 5, false, false)
    XNTT_CASE_MAP(FieldShoup, 6, false, false)
    XNTT_CASE_MAP(FieldShoup, 7, false, false)
    XNTT_CASE_MAP(FieldShoup, 8, false, false)
    XNTT_CASE_MAP(FieldShoup, 9, false, false)
    XNTT_CASE_MAP(FieldShoup, 10, false, false)
    XNTT_CASE_MAP(FieldShoup, 11, false, false)
    XNTT_CASE_MAP(FieldShoup, 12, false, false)
    XNTT_CASE_MAP(FieldShoup, 13, false, false)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
