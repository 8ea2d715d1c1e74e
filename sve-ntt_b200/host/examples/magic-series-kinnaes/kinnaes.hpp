// SPDX-License-Identifier: Apache-2.0
//
// MagicSeriesKinnaes<m, modmul_type, n> - drop-in for the class of the reference's second example
// (examples/magic-series-kinnaes/kinnaes.hpp:10-161): same template parameters, same static members
// (get_m / get_n / get_r, compute, compute_comb, compute_sum), the sum itself running on the GPU through
// xntt_kinnaes_sum (one thread per j instead of one SVE lane per j).  modmul_type only has to name its
// modulus_type; every modmul tag computes the same residues.
#ifndef XNTT_EXAMPLES_MAGIC_SERIES_KINNAES_HPP
#define XNTT_EXAMPLES_MAGIC_SERIES_KINNAES_HPP

#include <cstdint>
#include <stdexcept>

#include <sventt/sventt.hpp>

#include "xntt.h"

template <std::uint64_t m, class modmul_type_, std::uint64_t n>
class MagicSeriesKinnaes {
 public:
  using modmul_type = modmul_type_;
  using modulus_type = typename modmul_type::modulus_type;

  static constexpr std::uint64_t r{m * (m - 1) / 2 * m};

  static constexpr std::uint64_t get_m() { return m; }
  static constexpr std::uint64_t get_n() { return n; }
  static constexpr std::uint64_t get_r() { return r; }

  // (2 * compute_sum() + binomial(m^2, m)) / n   (kinnaes.hpp:27-34)
  static std::uint64_t compute() {
    std::uint64_t sum = compute_sum();
    sum = modulus_type::add(sum, sum);
    sum = modulus_type::add(sum, compute_comb(m * m, m));
    return modulus_type::divide(sum, n % modulus_type::get_modulus());
  }

  // binomial(a, b) mod N   (kinnaes.hpp:36-47)
  static std::uint64_t compute_comb(const std::uint64_t a, const std::uint64_t b) {
    constexpr std::uint64_t N = modulus_type::get_modulus();
    std::uint64_t num{a % N};
    for (std::uint64_t i{1}; i < b; ++i) num = modulus_type::multiply(num, (a - i) % N);
    std::uint64_t den{b % N};
    for (std::uint64_t i{2}; i < b; ++i) den = modulus_type::multiply(den, i % N);
    return modulus_type::divide(num, den);
  }

  static std::uint64_t compute_sum() { return compute_sum(0, n / 2); }

  // kinnaes.hpp:51-157
  static std::uint64_t compute_sum(const std::uint64_t j_begin, const std::uint64_t j_end) {
    std::uint64_t out = 0;
    const int status = xntt_kinnaes_sum(modulus_type::get_modulus(), modulus_type::get_generator(), m, n, j_begin,
                                        j_end, -1, &out);
    if (status == XNTT_ERR_INVALID) throw std::invalid_argument{"MagicSeriesKinnaes: invalid parameters"};
    if (status != XNTT_OK) throw std::runtime_error{xntt_last_cuda_error()};
    return out;
  }
};

#endif
