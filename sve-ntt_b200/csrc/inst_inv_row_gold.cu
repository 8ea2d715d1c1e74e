// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: inv_row, field FGold.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_inv_row_gold(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE(FGold, 1, false, true)
    XNTT_CASE(FGold, 2, false, true)
    XNTT_CASE(FGold, 3, false, true)
    XNTT_CASE(FGold, 4, false, true)
    XNTT_CASE(FGold, 5, false, true)
    XNTT_CASE(FGold, 6, false, true)
    XNTT_CASE(FGold, 7, false, true)
    XNTT_CASE(FGold, 8, false, true)
    XNTT_CASE(FGold, 9, false, true)
    XNTT_CASE(FGold, 10, false, true)
    XNTT_CASE(FGold, 11, false, true)
    XNTT_CASE(FGold, 12, false, true)
    XNTT_CASE(FGold, 13, false, true)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
