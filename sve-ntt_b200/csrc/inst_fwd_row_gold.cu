// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: fwd_row, field FGold.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_fwd_row_gold(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE(FGold, 1, false, false)
    XNTT_CASE(FGold, 2, false, false)
    XNTT_CASE(FGold, 3, false, false)
    XNTT_CASE(FGold, 4, false, false)
    XNTT_CASE(FGold, 5, false, false)
    XNTT_CASE(FGold, 6, false, false)
    XNTT_CASE(FGold, 7, false, false)
    XNTT_CASE(FGold, 8, false, false)
    XNTT_CASE(FGold, 9, false, false)
    XNTT_CASE(FGold, 10, false, false)
    XNTT_CASE(FGold, 11, false, false)
    XNTT_CASE(FGold, 12, false, false)
    XNTT_CASE(FGold, 13, false, false)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
