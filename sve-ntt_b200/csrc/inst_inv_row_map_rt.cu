// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: inv_row, field FieldRT, generalised address maps.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_inv_row_map_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE_MAP(FieldRT, 1, false, true)
    XNTT_CASE_MAP(FieldRT, 2, false, true)
    XNTT_CASE_MAP(FieldRT, 3, false, true)
    XNTT_CASE_MAP(FieldRT, 4, false, true)
    XNTT_CASE_MAP(FieldRT, 5, false, true)
    XNTT_CASE_MAP(FieldRT, 6, false, true)
    XNTT_CASE_MAP(FieldRT, 7, false, true)
    XNTT_CASE_MAP(FieldRT, 8, false, true)
    XNTT_CASE_MAP(FieldRT, 9, false, true)
    XNTT_CASE_MAP(FieldRT, 10, false, true)
    XNTT_CASE_MAP(FieldRT, 11, false, true)
    XNTT_CASE_MAP(FieldRT, 12, false, true)
    XNTT_CASE_MAP(FieldRT, 13, false, true)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
