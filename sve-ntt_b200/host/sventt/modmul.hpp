// SPDX-License-Identifier: Apache-2.0
//
// sventt::PAdic64<modulus_type> - the Montgomery (R = 2^64) modmul tag of the reference
// (include/sventt/modmul/sve/p-adic-64.hpp:13-248, scalar twin modmul/scalar/p-adic-64.hpp) as
// host scalar functions.  The device twin lives in csrc/field.cuh; this class is what user code
// such as examples/magic-series/gaussian-polynomial.hpp:176-193 calls on single words.
// PAdic64SVE / PAdic64Scalar are aliases so reference sources keep compiling.
#ifndef XNTT_SVENTT_MODMUL_HPP
#define XNTT_SVENTT_MODMUL_HPP

#include <cstdint>

namespace sventt {

template <class modulus_type_>
class PAdic64 {
  using u128 = unsigned __int128;

 public:
  using modulus_type = modulus_type_;
  static constexpr bool is_fixed_point = false;

  // b * 2^64 mod N   (p-adic-64.hpp:19-22)
  static constexpr std::uint64_t to_montgomery(std::uint64_t b) {
    constexpr std::uint64_t N = modulus_type::get_modulus();
    return static_cast<std::uint64_t>((static_cast<u128>(b % N) << 64) % N);
  }
  // b * 2^-64 mod N  (p-adic-64.hpp:24-27)
  static constexpr std::uint64_t from_montgomery(std::uint64_t b) {
    return multiply_normalize(b, 1, modulus_type::get_montgomery_inverse());
  }
  // b * N^-1 mod 2^64  (p-adic-64.hpp:64-67)
  static constexpr std::uint64_t precompute(std::uint64_t b) { return b * modulus_type::get_montgomery_inverse(); }

  // a * b * 2^-64 mod N, canonical  (p-adic-64.hpp:101-115).  With b = to_montgomery(w): a * w.
  static constexpr std::uint64_t multiply_normalize(std::uint64_t a, std::uint64_t b, std::uint64_t bp) {
    constexpr std::uint64_t N = modulus_type::get_modulus();
    const std::uint64_t q = a * bp;
    const std::uint64_t hi_ab = static_cast<std::uint64_t>((static_cast<u128>(a) * b) >> 64);
    const std::uint64_t hi_qn = static_cast<std::uint64_t>((static_cast<u128>(q) * N) >> 64);
    return hi_ab >= hi_qn ? hi_ab - hi_qn : hi_ab - hi_qn + N;
  }
  static constexpr std::uint64_t multiply_normalize(std::uint64_t a, std::uint64_t b) {
    return multiply_normalize(a, b, precompute(b));
  }
  // the reference's multiply() may return a lazily reduced value for <= 63-bit moduli; every value
  // it can return is congruent to this canonical one
  static constexpr std::uint64_t multiply(std::uint64_t a, std::uint64_t b, std::uint64_t bp) {
    return multiply_normalize(a, b, bp);
  }
  static constexpr std::uint64_t multiply(std::uint64_t a, std::uint64_t b) { return multiply_normalize(a, b); }
  static constexpr std::uint64_t add(std::uint64_t a, std::uint64_t b) { return modulus_type::add(a, b); }
  static constexpr std::uint64_t subtract(std::uint64_t a, std::uint64_t b) { return modulus_type::subtract(a, b); }
};

// sventt::FixedPoint64<modulus_type> - the Shoup-style alternative modmul tag
// (include/sventt/modmul/{scalar,sve}/fixed-point-64.hpp): to/from_montgomery are the identity,
// precompute(b) = floor(b * 2^64 / N), multiply(a, b, bp) = a * b mod N.  A composition whose layers ALL carry this
// tag runs the device kernels with Shoup arithmetic (csrc/field.cuh: FieldShoup, XNTT_MODMUL_FIXED_POINT) when the
// modulus is below 2^62; a composition that mixes it with PAdic64 (tests/ntt-tests/iterative-scalar-radix2-two10.hpp)
// or a larger modulus runs the Montgomery kernels - the transform computed is the same function either way.
template <class modulus_type_>
class FixedPoint64 {
  using u128 = unsigned __int128;

 public:
  using modulus_type = modulus_type_;
  static constexpr bool is_fixed_point = true;
  static constexpr std::uint64_t to_montgomery(std::uint64_t b) { return b; }
  static constexpr std::uint64_t from_montgomery(std::uint64_t b) { return b; }
  static constexpr std::uint64_t precompute(std::uint64_t b) {
    return static_cast<std::uint64_t>((static_cast<u128>(b) << 64) / modulus_type::get_modulus());
  }
  static constexpr std::uint64_t multiply_normalize(std::uint64_t a, std::uint64_t b, std::uint64_t bp) {
    constexpr std::uint64_t N = modulus_type::get_modulus();
    const std::uint64_t q = static_cast<std::uint64_t>((static_cast<u128>(a) * bp) >> 64);
    // a*b - q*N lies in [0, 2N): one conditional subtraction, carried out in 128 bits
    u128 r = static_cast<u128>(a) * b - static_cast<u128>(q) * N;
    if (r >= N) r -= N;
    return static_cast<std::uint64_t>(r);
  }
  static constexpr std::uint64_t multiply_normalize(std::uint64_t a, std::uint64_t b) {
    return multiply_normalize(a, b, precompute(b));
  }
  static constexpr std::uint64_t multiply(std::uint64_t a, std::uint64_t b, std::uint64_t bp) {
    return multiply_normalize(a, b, bp);
  }
  static constexpr std::uint64_t multiply(std::uint64_t a, std::uint64_t b) { return multiply_normalize(a, b); }
  static constexpr std::uint64_t add(std::uint64_t a, std::uint64_t b) { return modulus_type::add(a, b); }
  static constexpr std::uint64_t subtract(std::uint64_t a, std::uint64_t b) { return modulus_type::subtract(a, b); }
};
template <class modulus_type>
using FixedPoint64SVE = FixedPoint64<modulus_type>;
template <class modulus_type>
using FixedPoint64Scalar = FixedPoint64<modulus_type>;

template <class modulus_type>
using PAdic64SVE = PAdic64<modulus_type>;
template <class modulus_type>
using PAdic64Scalar = PAdic64<modulus_type>;

}  // namespace sventt

#endif
