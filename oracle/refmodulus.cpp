// SPDX-License-Identifier: Apache-2.0
//
// TEST INFRASTRUCTURE.  C entry points around sventt::Modulus for the two moduli the hot path is
// pinned on, compiled from /root/reference/include/sventt/modulus.hpp where it lies.
#include <cstdint>

#include "sventt/modulus.hpp"

using P0 = sventt::Modulus<UINT64_C(0xfffffc6e80000001), 3>;
using Goldilocks = sventt::Modulus<UINT64_C(0xffffffff00000001), 7>;

template <class M>
static std::uint64_t root(bool inverse, std::uint64_t order) {
  try {
    return inverse ? M::get_root_inverse(order) : M::get_root_forward(order);
  } catch (const std::invalid_argument&) {
    return 0;
  }
}

extern "C" {
std::uint64_t ref_root(int which, int inverse, std::uint64_t order) {
  return which == 0 ? root<P0>(inverse != 0, order) : root<Goldilocks>(inverse != 0, order);
}
std::uint64_t ref_montgomery_inverse(int which) {
  return which == 0 ? P0::get_montgomery_inverse() : Goldilocks::get_montgomery_inverse();
}
std::uint64_t ref_multiply(int which, std::uint64_t a, std::uint64_t b) {
  return which == 0 ? P0::multiply(a, b) : Goldilocks::multiply(a, b);
}
std::uint64_t ref_power(int which, std::uint64_t a, std::uint64_t e) {
  return which == 0 ? P0::power(a, e) : Goldilocks::power(a, e);
}
}
