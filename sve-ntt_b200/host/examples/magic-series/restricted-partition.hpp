// SPDX-License-Identifier: Apache-2.0
//
// RestrictedPartition<modulus_type> - drop-in for examples/magic-series/restricted-partition.hpp:11-51 of the
// reference: p(n, k) = number of partitions of n into parts of size at most k, modulo the modulus, enumerated
// n = 0, 1, 2, ... by advance().  (Coefficients of prod_{i=1..k} 1 / (1 - q^i).)
#ifndef XNTT_EXAMPLES_MAGIC_SERIES_RESTRICTED_PARTITION_HPP
#define XNTT_EXAMPLES_MAGIC_SERIES_RESTRICTED_PARTITION_HPP

#include <cstdint>
#include <vector>

template <class modulus_type_>
class RestrictedPartition {
 public:
  using modulus_type = modulus_type_;

  RestrictedPartition() = default;
  // window[j][n mod (k+1)] = p(n, j) for the last k + 1 values of n; p(0, j) = 1 for j >= 1 and, like the
  // reference (restricted-partition.hpp:24-27), p(0, 0) = 0
  explicit RestrictedPartition(std::uint64_t k) : n_{0}, k_{k}, window_((k + 1) * (k + 1), 0) {
    for (std::uint64_t j = 1; j <= k; ++j) at(j, 0) = 1 % modulus_type::get_modulus();
  }

  std::uint64_t get_n() const { return n_; }
  std::uint64_t get_k() const { return k_; }
  std::uint64_t operator()() const { return window_[k_ * (k_ + 1) + n_ % (k_ + 1)]; }

  // p(n, j) = p(n, j - 1) + p(n - j, j)
  void advance() {
    ++n_;
    at(0, n_) = 0;
    for (std::uint64_t j = 1; j <= k_; ++j) {
      const std::uint64_t fewer = at(j - 1, n_);
      const std::uint64_t shorter = n_ >= j ? at(j, n_ - j) : 0;
      at(j, n_) = modulus_type::add(fewer, shorter);
    }
  }

 private:
  std::uint64_t& at(std::uint64_t j, std::uint64_t n) { return window_[j * (k_ + 1) + n % (k_ + 1)]; }
  std::uint64_t n_{}, k_{};
  std::vector<std::uint64_t> window_;
};

#endif
