// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: fwd_row, field F0, generalised address maps.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_fwd_row_map(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE_MAP(F0, 1, false, false)
    XNTT_CASE_MAP(F0, 2, false, false)
    XNTT_CASE_MAP(F0, 3, false, false)
    XNTT_CASE_MAP(F0, 4, false, false)
    XNTT_CASE_MAP(F0, 5, false, false)
    XNTT_CASE_MAP(F0, 6, false, false)
    XNTT_CASE_MAP(F0, 7, false, false)
    XNTT_CASE_MAP(F0, 8, false, false)
    XNTT_CASE_MAP(F0, 9, false, false)
    XNTT_CASE_MAP(F0, 10, false, false)
    XNTT_CASE_MAP(F0, 11, false, false)
    XNTT_CASE_MAP(F0, 12, false, false)
    XNTT_CASE_MAP(F0, 13, false, false)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
