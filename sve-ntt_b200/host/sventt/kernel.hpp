// SPDX-License-Identifier: Apache-2.0
//
// sventt::IterativeNTT / sventt::RecursiveNTT - the composition classes of the reference
// (include/sventt/kernel/iterative.hpp:17-107, kernel/recursive.hpp:15-145) with the same template
// parameters and static_asserts.  Instead of carrying compute loops they describe themselves to the
// planner: transform length, accumulated inverse factor and - for six-step compositions - the
// n0 x n1 decomposition, which becomes xntt_desc::split_log2.
#ifndef XNTT_SVENTT_KERNEL_HPP
#define XNTT_SVENTT_KERNEL_HPP

#include <bit>
#include <cstdint>
#include <type_traits>
#include <vector>

#include "layer.hpp"

namespace sventt {

namespace detail {
constexpr std::uint32_t log2_exact(std::uint64_t v) { return static_cast<std::uint32_t>(std::countr_zero(v)); }
}  // namespace detail

template <class modulus_type_, std::uint64_t m, class... layer_types>
class IterativeNTT {
 public:
  using modulus_type = modulus_type_;

 private:
  static_assert(sizeof...(layer_types) >= 1);
  static_assert((std::is_same_v<modulus_type, typename layer_types::modulus_type> && ...));
  static_assert(((layer_types::get_m() == m) && ...));
  static_assert((layer_types::get_radix() * ...) == m, "the product of the layer radices must equal m");
  static_assert(std::has_single_bit(m), "transform length must be a power of two");

 public:
  static constexpr std::uint64_t get_m() { return m; }
  // product of the layers' inverse factors (the reference folds each one into that layer's butterflies)
  static constexpr std::uint64_t get_inverse_factor() {
    std::uint64_t f = 1;
    ((f = modulus_type::multiply(f, layer_types::get_inverse_factor())), ...);
    return f;
  }
  // every layer carries the FixedPoint64 tag: run the Shoup kernels (wrapper.hpp sets XNTT_MODMUL_FIXED_POINT)
  static constexpr bool all_fixed_point() { return (layer_types::all_fixed_point() && ...); }
  // one contiguous transform of length m: a single planner unit
  static void append_splits(std::vector<std::uint32_t>& out) { out.push_back(detail::log2_exact(m)); }

};

// separate_twiddle has no default in the reference (kernel/recursive.hpp:15-16); three of its own (disabled) test shapes,
// tests/ntt-tests/recursive-scalar-*.hpp, still spell RecursiveNTT with four arguments, so a default is accepted here.
template <class modulus_type_, std::uint64_t m, class layer_type_, class inner_kernel_type_, bool separate_twiddle = false>
class RecursiveNTT {
 public:
  using modulus_type = modulus_type_;
  using layer_type = layer_type_;
  using inner_kernel_type = inner_kernel_type_;

 private:
  static_assert(std::is_same_v<modulus_type, typename layer_type::modulus_type>);
  static_assert(std::is_same_v<modulus_type, typename inner_kernel_type::modulus_type>);
  static_assert(layer_type::get_m() == m);
  static_assert(inner_kernel_type::get_m() * layer_type::get_radix() == m);
  static_assert(std::has_single_bit(m), "transform length must be a power of two");

 public:
  static constexpr std::uint64_t get_m() { return m; }
  static constexpr std::uint64_t get_inverse_factor() {
    return modulus_type::multiply(layer_type::get_inverse_factor(), inner_kernel_type::get_inverse_factor());
  }
  static constexpr bool all_fixed_point() { return layer_type::all_fixed_point() && inner_kernel_type::all_fixed_point(); }
  static void append_splits(std::vector<std::uint32_t>& out) {
    if constexpr (layer_type::is_six_step()) {
      // (blocked) six-step: n0 = radix column transforms, then the inner kernel on every row
      out.push_back(detail::log2_exact(layer_type::get_radix()));
      inner_kernel_type::append_splits(out);
    } else {
      // a plain radix layer in front of an inner kernel is the same transform of length m
      out.push_back(detail::log2_exact(m));
    }
    (void)separate_twiddle;
  }
};

}  // namespace sventt

#endif
