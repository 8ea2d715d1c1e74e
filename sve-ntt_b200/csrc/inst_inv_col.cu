// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: inv_col.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_inv_col(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE(1, true, true)
    XNTT_CASE(2, true, true)
    XNTT_CASE(3, true, true)
    XNTT_CASE(4, true, true)
    XNTT_CASE(5, true, true)
    XNTT_CASE(6, true, true)
    XNTT_CASE(7, true, true)
    XNTT_CASE(8, true, true)
    XNTT_CASE(9, true, true)
    XNTT_CASE(10, true, true)
    XNTT_CASE(11, true, true)
    XNTT_CASE(12, true, true)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
