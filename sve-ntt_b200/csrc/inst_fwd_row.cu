// SPDX-License-Identifier: Apache-2.0
// Instantiations of pass_kernel: fwd_row.
#include "dispatch.cuh"
namespace xntt {
cudaError_t launch_fwd_row(int logn, const PassParams& prm, unsigned grid, cudaStream_t st) {
  switch (logn) {
    XNTT_CASE(1, false, false)
    XNTT_CASE(2, false, false)
    XNTT_CASE(3, false, false)
    XNTT_CASE(4, false, false)
    XNTT_CASE(5, false, false)
    XNTT_CASE(6, false, false)
    XNTT_CASE(7, false, false)
    XNTT_CASE(8, false, false)
    XNTT_CASE(9, false, false)
    XNTT_CASE(10, false, false)
    XNTT_CASE(11, false, false)
    XNTT_CASE(12, false, false)
    XNTT_CASE(13, false, false)
    default:
      return cudaErrorInvalidValue;
  }
}
}  // namespace xntt
