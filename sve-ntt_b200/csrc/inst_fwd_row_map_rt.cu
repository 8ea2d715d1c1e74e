// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: fwd_row, field FieldRT, generalised address maps.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_fwd_row_map_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE_MAP(FieldRT, 1, false, false)
    XNTT_CASE_MAP(FieldRT, 2, false, false)
    XNTT_CASE_MAP(FieldRT, 3, false, false)
    XNTT_CASE_MAP(FieldRT, 4, false, false)
    XNTT_CASE_MAP(FieldRT, 5, false, false)
    XNTT_CASE_MAP(FieldRT, 6, false, false)
    XNTT_CASE_MAP(FieldRT, 7, false, false)
    XNTT_CASE_MAP(FieldRT, 8, false, false)
    XNTT_CASE_MAP(FieldRT, 9, false, false)
    XNTT_CASE_MAP(FieldRT, 10, false, false)
    XNTT_CASE_MAP(FieldRT, 11, false, false)
    XNTT_CASE_MAP(FieldRT, 12, false, false)
    XNTT_CASE_MAP(FieldRT, 13, false, false)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
