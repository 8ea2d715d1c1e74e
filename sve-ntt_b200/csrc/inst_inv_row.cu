// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: inv_row.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_inv_row(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE(1, false, true)
    XNTT_CASE(2, false, true)
    XNTT_CASE(3, false, true)
    XNTT_CASE(4, false, true)
    XNTT_CASE(5, false, true)
    XNTT_CASE(6, false, true)
    XNTT_CASE(7, false, true)
    XNTT_CASE(8, false, true)
    XNTT_CASE(9, false, true)
    XNTT_CASE(10, false, true)
    XNTT_CASE(11, false, true)
    XNTT_CASE(12, false, true)
    XNTT_CASE(13, false, true)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
