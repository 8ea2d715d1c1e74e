// SPDX-License-Identifier: Apache-2.0
//
// sventt::Modulus<p, g> - compile-time field constants with the interface of the reference's
// include/sventt/modulus.hpp:14-133 (get_modulus, get_generator, get_montgomery_inverse,
// get_shoup_inverse, reduce/negate/add/subtract/multiply/divide/power/invert,
// get_root_forward/get_root_inverse throwing std::invalid_argument when order does not divide p-1).
// Host-side and constexpr: it feeds the plan descriptor, never the data path.
#ifndef XNTT_SVENTT_MODULUS_HPP
#define XNTT_SVENTT_MODULUS_HPP

#include <cstdint>
#include <stdexcept>

namespace sventt {

template <std::uint64_t modulus, std::uint64_t generator = 0>
class Modulus {
  using u128 = unsigned __int128;

 public:
  struct shoup_inverse_type {
    std::uint64_t modulus_inverse_lo, modulus_inverse_hi;
  };

  static constexpr std::uint64_t get_modulus() { return modulus; }
  static constexpr std::uint64_t get_generator() { return generator; }

  // floor((2^128 - 1) / p), or 2^128 / p exactly when p is a power of two
  static constexpr shoup_inverse_type get_shoup_inverse() {
    const bool pow2 = (modulus & (modulus - 1)) == 0;
    u128 inv = ~u128{0} / modulus;
    if (pow2) inv += 1;
    return {static_cast<std::uint64_t>(inv), static_cast<std::uint64_t>(inv >> 64)};
  }

  // p^-1 mod 2^64 by Newton iteration (each step doubles the number of correct bits)
  static constexpr std::uint64_t get_montgomery_inverse() {
    std::uint64_t x = modulus;  // correct to 3 bits for odd p
    for (int i = 0; i < 6; ++i) x *= 2 - modulus * x;
    return x;
  }

  static constexpr std::uint64_t reduce(std::uint64_t a) { return a % modulus; }
  static constexpr std::uint64_t negate(std::uint64_t a) { return subtract(0, a); }
  static constexpr std::uint64_t add(std::uint64_t a, std::uint64_t b) {
    a %= modulus;
    b %= modulus;
    const std::uint64_t room = modulus - b;
    return a < room ? a + b : a - room;
  }
  static constexpr std::uint64_t subtract(std::uint64_t a, std::uint64_t b) {
    a %= modulus;
    b %= modulus;
    return a >= b ? a - b : modulus - (b - a);
  }
  static constexpr std::uint64_t multiply(std::uint64_t a, std::uint64_t b) {
    return static_cast<std::uint64_t>(static_cast<u128>(a) * b % modulus);
  }
  static constexpr std::uint64_t power(std::uint64_t a, std::uint64_t e) {
    std::uint64_t acc = 1 % modulus;
    while (e != 0) {
      if (e & 1) acc = multiply(acc, a);
      a = multiply(a, a);
      e >>= 1;
    }
    return acc;
  }
  static constexpr std::uint64_t invert(std::uint64_t a) { return power(a, modulus - 2); }
  static constexpr std::uint64_t divide(std::uint64_t a, std::uint64_t b) { return multiply(a, invert(b)); }

  static constexpr std::uint64_t get_root_forward(std::uint64_t order)
    requires(generator != 0)
  {
    if (order == 0 || (modulus - 1) % order != 0) throw std::invalid_argument{"the field has no such root"};
    return power(generator, (modulus - 1) / order);
  }
  static constexpr std::uint64_t get_root_inverse(std::uint64_t order)
    requires(generator != 0)
  {
    if (order == 0 || (modulus - 1) % order != 0) throw std::invalid_argument{"the field has no such root"};
    // g^(-(p-1)/order) = g^((p-1) - (p-1)/order)
    const std::uint64_t step = (modulus - 1) / order;
    return power(generator, (modulus - 1) - step % (modulus - 1));
  }
};

}  // namespace sventt

#endif
