// SPDX-License-Identifier: Apache-2.0
//
// Drop-in for examples/magic-series/gaussian-polynomial.hpp of the reference: one coefficient of the Gaussian
// polynomial [n choose k]_q = prod_{i=1..k} (1 - q^(n-k+i)) / prod_{i=1..k} (1 - q^i) by power-series division in
// blocks of half the NTT length - the polynomial multiply (forward, point-wise multiply_normalize against a
// to_montgomery'd spectrum, inverse; gaussian-polynomial.hpp:148-244) running through sventt::NTT on the GPU.
// Same entry points: calculate_q_pochhammer, GaussianPolynomialNumerator(Segment),
// calculate_gaussian_polynomial_coefficient(n, k, d, ntt), calculate_number_of_magic_series(m, ntt).
//
// Division scheme (written for this port): with c = m / 2 coefficients per block, quotient S = N / D, E = 1 / D
// mod q^c and deg D < c, block j of S is  S_j = ((N_j - carry_j) * E) mod q^c  with
// carry_{j+1} = the upper half of S_j * D; both products are cyclic convolutions of length 2c.
#ifndef XNTT_EXAMPLES_MAGIC_SERIES_GAUSSIAN_POLYNOMIAL_HPP
#define XNTT_EXAMPLES_MAGIC_SERIES_GAUSSIAN_POLYNOMIAL_HPP

#include <algorithm>
#include <cstdint>
#include <ranges>
#include <stdexcept>
#include <vector>

#include <sventt/sventt.hpp>

#include "restricted-partition.hpp"

// coefficients of prod_{i=1..k} (1 - q^i)   (gaussian-polynomial.hpp:19-45)
template <class modulus_type, class range_type>
static void calculate_q_pochhammer(range_type&& coefficients, const std::uint64_t k)
  requires(std::ranges::random_access_range<range_type> &&
           std::same_as<std::ranges::range_value_t<range_type>, std::uint64_t>)
{
  const std::uint64_t degree = k * (k + 1) / 2;
  if (std::ranges::size(coefficients) < degree + 1) throw std::invalid_argument{"coefficient vector is too small"};
  for (std::uint64_t i = 0; i <= degree; ++i) coefficients[i] = 0;
  coefficients[0] = 1 % modulus_type::get_modulus();
  std::uint64_t top = 0;  // current degree
  for (std::uint64_t i = 1; i <= k; ++i) {
    // times (1 - q^i): from the top down so that every source coefficient is still the old one
    for (std::uint64_t t = top + i; t >= i; --t)
      coefficients[t] = modulus_type::subtract(coefficients[t], coefficients[t - i]);
    top += i;
  }
}

// [k choose j]_q for j = 0, 1, ..., one advance() at a time  (gaussian-polynomial.hpp:52-107):
// [k j] = [k j-1] * (1 - q^(k-j+1)) / (1 - q^j)
template <class modulus_type_>
class GaussianPolynomialNumeratorSegment {
 public:
  using modulus_type = modulus_type_;
  GaussianPolynomialNumeratorSegment() = default;
  explicit GaussianPolynomialNumeratorSegment(std::uint64_t k) : k_{k}, j_{0} {}
  std::uint64_t get_k() const { return k_; }
  std::uint64_t get_j() const { return j_; }  // the segment held is [k choose j - 1]_q once advanced
  const std::vector<std::uint64_t>& get_coefficients() const { return c_; }

  void advance() {
    if (j_ == 0 || j_ == k_) {  // [k 0] = [k k] = 1
      c_.assign(1, 1 % modulus_type::get_modulus());
      if (j_ == 0) ++j_;
      return;
    }
    const std::uint64_t a = k_ - j_ + 1, deg = j_ * (k_ - j_);
    c_.resize(deg + 1, 0);
    for (std::uint64_t t = deg; t >= a; --t) c_[t] = modulus_type::subtract(c_[t], c_[t - a]);   // * (1 - q^a)
    for (std::uint64_t t = j_; t <= deg; ++t) c_[t] = modulus_type::add(c_[t], c_[t - j_]);      // / (1 - q^j)
    ++j_;
  }

 private:
  std::uint64_t k_{}, j_{};
  std::vector<std::uint64_t> c_;
};

// prod_{i=1..k} (1 - q^(n-k+i)) = sum_j (-1)^j q^(j (n-k+1) + j (j-1) / 2) [k choose j]_q, streamed in order of
// increasing degree; subtract_next(minuend, size) subtracts the next `size` coefficients from minuend
// (gaussian-polynomial.hpp:109-146).  The segments must not overlap (checked by the caller).
template <class modulus_type_>
class GaussianPolynomialNumerator {
 public:
  using modulus_type = modulus_type_;
  GaussianPolynomialNumerator() = default;
  GaussianPolynomialNumerator(std::uint64_t n, std::uint64_t k) : n_{n}, k_{k}, segment_{k} {}

  void subtract_next(std::uint64_t* const minuend, const std::uint64_t size) {
    for (std::uint64_t pos = 0; pos < size; ++pos, ++degree_) {
      // move on to the segment that starts at or before this degree
      while (j_ <= k_ && (!loaded_ || (j_ < k_ && degree_ >= start(j_ + 1)))) {
        if (loaded_) ++j_;
        segment_.advance();
        loaded_ = true;
      }
      if (j_ > k_ || degree_ < start(j_)) continue;
      const std::uint64_t t = degree_ - start(j_);
      const auto& c = segment_.get_coefficients();
      if (t >= c.size()) continue;
      minuend[pos] = (j_ % 2 == 1 ? modulus_type::add : modulus_type::subtract)(minuend[pos], c[t]);
    }
  }

 private:
  std::uint64_t start(std::uint64_t j) const { return j * (n_ - k_ + 1) + j * (j - 1) / 2; }
  std::uint64_t n_{}, k_{}, j_{}, degree_{};
  bool loaded_{false};
  GaussianPolynomialNumeratorSegment<modulus_type> segment_;
};

template <class ntt_type>
static std::uint64_t calculate_gaussian_polynomial_coefficient(const std::uint64_t n, const std::uint64_t k,
                                                               const std::uint64_t d, const ntt_type& ntt) {
  using modulus_type = typename ntt_type::modulus_type;
  using modmul_type = sventt::PAdic64SVE<modulus_type>;
  if (d > k * (n - k)) throw std::invalid_argument{"d is out of range"};
  if (n < (k * k + 2 * k + k % 2 + 3) / 4) throw std::invalid_argument{"n is too small; segments will overlap"};
  if (ntt_type::get_m() < (k * (k + 1) / 2 + 1) * 2) throw std::invalid_argument{"NTT length is too small"};
  constexpr std::uint64_t m = ntt_type::get_m(), c = m / 2;

  // spectra of D = prod (1 - q^i) and of E = 1 / D mod q^c, in Montgomery form for multiply_normalize
  sventt::PageMemory<std::uint64_t> d_hat(m), e_hat(m), x(m);
  std::fill(d_hat.begin(), d_hat.end(), 0);
  calculate_q_pochhammer<modulus_type>(std::ranges::subrange(d_hat.begin(), d_hat.end()), k);
  ntt.compute_forward(d_hat.data());
  std::fill(e_hat.begin(), e_hat.end(), 0);
  {
    RestrictedPartition<modulus_type> partition(k);
    for (std::uint64_t i = 0; i < c; ++i, partition.advance()) e_hat[i] = i == 0 ? 1 % modulus_type::get_modulus() : partition();
  }
  ntt.compute_forward(e_hat.data());
  for (std::uint64_t i = 0; i < m; ++i) {
    d_hat[i] = modmul_type::to_montgomery(d_hat[i]);
    e_hat[i] = modmul_type::to_montgomery(e_hat[i]);
  }

  GaussianPolynomialNumerator<modulus_type> numerator(n, k);
  std::fill(x.begin(), x.end(), 0);  // low half: the carry, starts at 0
  for (std::uint64_t base = 0; base <= d; base += c) {
    // x[0..c) = N_block - carry, upper half zero
    numerator.subtract_next(x.data(), c);  // carry - N_block
    for (std::uint64_t i = 0; i < c; ++i) x[i] = modulus_type::negate(x[i]);
    // S_block = (x * E) mod q^c
    ntt.compute_forward(x.data());
    for (std::uint64_t i = 0; i < m; ++i) x[i] = modmul_type::multiply_normalize(x[i], e_hat[i]);
    ntt.compute_inverse(x.data());
    if (d < base + c) return x[d - base];
    // carry = upper half of S_block * D
    std::fill(x.begin() + c, x.end(), 0);
    ntt.compute_forward(x.data());
    for (std::uint64_t i = 0; i < m; ++i) x[i] = modmul_type::multiply_normalize(x[i], d_hat[i]);
    ntt.compute_inverse(x.data());
    for (std::uint64_t i = 0; i < c; ++i) x[i] = x[c + i];
    std::fill(x.begin() + c, x.end(), 0);
  }
  throw std::runtime_error{"internal error"};
}

// number of magic series of order m = coefficient of q^(m^2 (m-1) / 2) of [m^2 choose m]_q
// (gaussian-polynomial.hpp:246-251)
template <class ntt_type>
static std::uint64_t calculate_number_of_magic_series(const std::uint64_t m, const ntt_type& ntt) {
  return calculate_gaussian_polynomial_coefficient(m * m, m, m * m * (m - 1) / 2, ntt);
}

#endif
