// SPDX-License-Identifier: Apache-2.0
// sventt::bitreverse - 64-bit bit reversal (reference: include/sventt/utility.hpp:12-23).
#ifndef XNTT_SVENTT_UTILITY_HPP
#define XNTT_SVENTT_UTILITY_HPP

#include <cstdint>

namespace sventt {

static inline constexpr std::uint64_t bitreverse(std::uint64_t x) {
  std::uint64_t r = 0;
  for (int i = 0; i < 64; ++i) {
    r = (r << 1) | (x & 1);
    x >>= 1;
  }
  return r;
}

}  // namespace sventt

#endif
