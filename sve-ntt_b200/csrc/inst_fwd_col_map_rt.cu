// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: fwd_col, field FieldRT, generalised address maps.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_fwd_col_map_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE_MAP(FieldRT, 1, true, false)
    XNTT_CASE_MAP(FieldRT, 2, true, false)
    XNTT_CASE_MAP(FieldRT, 3, true, false)
    XNTT_CASE_MAP(FieldRT, 4, true, false)
    XNTT_CASE_MAP(FieldRT, 5, true, false)
    XNTT_CASE_MAP(FieldRT, 6, true, false)
    XNTT_CASE_MAP(FieldRT, 7, true, false)
    XNTT_CASE_MAP(FieldRT, 8, true, false)
    XNTT_CASE_MAP(FieldRT, 9, true, false)
    XNTT_CASE_MAP(FieldRT, 10, true, false)
    XNTT_CASE_MAP(FieldRT, 11, true, false)
    XNTT_CASE_MAP(FieldRT, 12, true, false)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
