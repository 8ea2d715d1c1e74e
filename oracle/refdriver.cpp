// SPDX-License-Identifier: Apache-2.0
//
// TEST INFRASTRUCTURE.  C entry points around the reference's own oracle class, compiled from the
// reference sources where they lie (/root/reference/tests/ntt-reference.hpp is #included, never
// copied).  Output: oracle/_ref/libnttref.so (git-ignored, travels to the GPU box).
// Used to pin oracle/ntt_oracle.c, to generate tests/golden/, and as the "reference" CPU baseline.
#include <cstdint>

#include "ntt-reference.hpp"

extern "C" {
void ref_ntt_forward(std::uint64_t* dst, const std::uint64_t* src, std::uint64_t m, std::uint64_t N,
                     std::uint64_t g) {
  const NTTReference ntt{m, N, g};
  ntt.compute_forward(dst, src);
}
void ref_ntt_inverse(std::uint64_t* dst, const std::uint64_t* src, std::uint64_t m, std::uint64_t N,
                     std::uint64_t g) {
  const NTTReference ntt{m, N, g};
  ntt.compute_inverse(dst, src);
}
}
