// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: fwd_col.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_fwd_col(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE(1, true, false)
    XNTT_CASE(2, true, false)
    XNTT_CASE(3, true, false)
    XNTT_CASE(4, true, false)
    XNTT_CASE(5, true, false)
    XNTT_CASE(6, true, false)
    XNTT_CASE(7, true, false)
    XNTT_CASE(8, true, false)
    XNTT_CASE(9, true, false)
    XNTT_CASE(10, true, false)
    XNTT_CASE(11, true, false)
    XNTT_CASE(12, true, false)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
