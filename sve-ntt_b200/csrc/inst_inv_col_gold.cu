// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: inv_col, field FGold.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_inv_col_gold(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE(FGold, 1, true, true)
    XNTT_CASE(FGold, 2, true, true)
    XNTT_CASE(FGold, 3, true, true)
    XNTT_CASE(FGold, 4, true, true)
    XNTT_CASE(FGold, 5, true, true)
    XNTT_CASE(FGold, 6, true, true)
    XNTT_CASE(FGold, 7, true, true)
    XNTT_CASE(FGold, 8, true, true)
    XNTT_CASE(FGold, 9, true, true)
    XNTT_CASE(FGold, 10, true, true)
    XNTT_CASE(FGold, 11, true, true)
    XNTT_CASE(FGold, 12, true, true)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
