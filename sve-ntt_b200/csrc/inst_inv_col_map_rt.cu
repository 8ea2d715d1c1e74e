// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: inv_col, field FieldRT, generalised address maps.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_inv_col_map_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE_MAP(FieldRT, 1, true, true)
    XNTT_CASE_MAP(FieldRT, 2, true, true)
    XNTT_CASE_MAP(FieldRT, 3, true, true)
    XNTT_CASE_MAP(FieldRT, 4, true, true)
    XNTT_CASE_MAP(FieldRT, 5, true, true)
    XNTT_CASE_MAP(FieldRT, 6, true, true)
    XNTT_CASE_MAP(FieldRT, 7, true, true)
    XNTT_CASE_MAP(FieldRT, 8, true, true)
    XNTT_CASE_MAP(FieldRT, 9, true, true)
    XNTT_CASE_MAP(FieldRT, 10, true, true)
    XNTT_CASE_MAP(FieldRT, 11, true, true)
    XNTT_CASE_MAP(FieldRT, 12, true, true)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
