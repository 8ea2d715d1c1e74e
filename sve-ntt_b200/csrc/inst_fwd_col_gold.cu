// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: fwd_col, field FGold.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_fwd_col_gold(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE(FGold, 1, true, false)
    XNTT_CASE(FGold, 2, true, false)
    XNTT_CASE(FGold, 3, true, false)
    XNTT_CASE(FGold, 4, true, false)
    XNTT_CASE(FGold, 5, true, false)
    XNTT_CASE(FGold, 6, true, false)
    XNTT_CASE(FGold, 7, true, false)
    XNTT_CASE(FGold, 8, true, false)
    XNTT_CASE(FGold, 9, true, false)
    XNTT_CASE(FGold, 10, true, false)
    XNTT_CASE(FGold, 11, true, false)
    XNTT_CASE(FGold, 12, true, false)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
