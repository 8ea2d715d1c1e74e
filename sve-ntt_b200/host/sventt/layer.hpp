// SPDX-License-Identifier: Apache-2.0
//
// Layer descriptors with the template surface of the reference's layers:
//   RadixEightSVELayer / RadixFourSVELayer / RadixTwoSVELayer <modmul, m, n, inverse_factor = 1,
//       store_precomputation = true>     (include/sventt/layer/sve/radix-{eight,four,two}.hpp)
//   GenericSVELayer<modmul, m, inner_kernel, buffer_padding, twiddle_unroll, transposition,
//       transpose_in_place = false>      (include/sventt/layer/sve/generic.hpp:22-30)
//   BlockedGenericSVELayer<modmul, m, inner_kernel, block_padding, twiddle_unroll, block_rows,
//       transposition>                   (include/sventt/layer/sve/blocked-generic.hpp:23-27)
// and the scalar spellings.  On the B200 a layer is pure type-level information: the composition is
// lowered to an xntt plan (kernel.hpp), and the knobs that only exist for SVE cache blocking
// (paddings, unroll counts, block rows, transposition class, store_precomputation) are accepted and
// ignored - the shared-memory tile of csrc/pass_kernel.cuh plays their role.
#ifndef XNTT_SVENTT_LAYER_HPP
#define XNTT_SVENTT_LAYER_HPP

#include <cstdint>
#include <stdexcept>

#include "xntt.h"

namespace sventt {

namespace detail {
template <class modmul_type_, std::uint64_t radix, std::uint64_t m, std::uint64_t n,
          std::uint64_t inverse_factor, bool store_precomputation>
class RadixLayer {
 public:
  using modmul_type = modmul_type_;
  using modulus_type = typename modmul_type::modulus_type;
  static constexpr std::uint64_t get_radix() { return radix; }
  static constexpr std::uint64_t get_m() { return m; }
  static constexpr std::uint64_t get_n() { return n; }
  static constexpr std::uint64_t get_inverse_factor() { return inverse_factor; }
  static constexpr bool is_six_step() { return false; }
  static constexpr bool all_fixed_point() { return modmul_type::is_fixed_point; }
  static_assert(n >= radix, "sub-transform shorter than the radix");
  static_assert(m % n == 0, "n must divide m");
  // radix-two.hpp:208-211: only a terminal layer (n == radix) may carry the inverse factor
  static_assert(inverse_factor == 1 || n == radix, "inverse_factor is only allowed on a terminal layer");
};
}  // namespace detail

#define XNTT_RADIX_LAYER(NAME, RADIX)                                                              \
  template <class modmul_type, std::uint64_t m, std::uint64_t n, std::uint64_t inverse_factor = 1, \
            bool store_precomputation = true>                                                      \
  using NAME = detail::RadixLayer<modmul_type, RADIX, m, n, inverse_factor, store_precomputation>;

XNTT_RADIX_LAYER(RadixEightLayer, 8)
XNTT_RADIX_LAYER(RadixFourLayer, 4)
XNTT_RADIX_LAYER(RadixTwoLayer, 2)
XNTT_RADIX_LAYER(RadixEightSVELayer, 8)
XNTT_RADIX_LAYER(RadixFourSVELayer, 4)
XNTT_RADIX_LAYER(RadixTwoSVELayer, 2)
XNTT_RADIX_LAYER(RadixEightScalarLayer, 8)
XNTT_RADIX_LAYER(RadixFourScalarLayer, 4)
XNTT_RADIX_LAYER(RadixTwoScalarLayer, 2)
#undef XNTT_RADIX_LAYER

// Six-step outer layer: its radix is a whole inner transform of length n0 = inner_kernel::get_m().
template <class modmul_type_, std::uint64_t m, class inner_kernel_type_, std::uint64_t buffer_padding_elements = 0,
          std::uint64_t twiddle_unroll_count = 1, class transposition_type = void, bool transpose_in_place = false>
class GenericLayer {
 public:
  using modmul_type = modmul_type_;
  using modulus_type = typename modmul_type::modulus_type;
  using inner_kernel_type = inner_kernel_type_;
  static constexpr std::uint64_t get_m() { return m; }
  static constexpr std::uint64_t get_radix() { return inner_kernel_type::get_m(); }
  static constexpr std::uint64_t get_inverse_factor() { return inner_kernel_type::get_inverse_factor(); }
  static constexpr bool is_six_step() { return true; }
  static constexpr bool all_fixed_point() { return modmul_type::is_fixed_point && inner_kernel_type::all_fixed_point(); }
  static_assert(m % inner_kernel_type_::get_m() == 0);
};
template <class modmul_type, std::uint64_t m, class inner_kernel_type, std::uint64_t buffer_padding_elements = 0,
          std::uint64_t twiddle_unroll_count = 1, class transposition_type = void, bool transpose_in_place = false>
using GenericSVELayer = GenericLayer<modmul_type, m, inner_kernel_type, buffer_padding_elements,
                                     twiddle_unroll_count, transposition_type, transpose_in_place>;
template <class modmul_type, std::uint64_t m, class inner_kernel_type>
using GenericScalarLayer = GenericLayer<modmul_type, m, inner_kernel_type>;

// Blocked six-step: same mathematics, tile by tile.  block_rows is the reference's tile width; here
// the tile width follows from the 64 KiB shared-memory tile.
template <class modmul_type_, std::uint64_t m, class inner_kernel_type_, std::uint64_t block_padding_elements = 0,
          std::uint64_t twiddle_unroll_count = 1, std::uint64_t block_rows = 0, class transposition_type = void>
class BlockedGenericLayer : public GenericLayer<modmul_type_, m, inner_kernel_type_> {};
template <class modmul_type, std::uint64_t m, class inner_kernel_type, std::uint64_t block_padding_elements = 0,
          std::uint64_t twiddle_unroll_count = 1, std::uint64_t block_rows = 0, class transposition_type = void>
using BlockedGenericSVELayer = BlockedGenericLayer<modmul_type, m, inner_kernel_type, block_padding_elements,
                                                   twiddle_unroll_count, block_rows, transposition_type>;

// Transposition classes of include/sventt/transposition/sve/ (17 variants of one contract,
// transpose(dst, src, src_rows, src_cols, ld_dst, ld_src) => dst[ld_dst * c + r] = src[ld_src * r + c], and
// the in-place square transpose(dst, dim)).  As layer arguments they are tags - inside a transform the
// column pass reads its strided tile directly - and called on their own they run libxntt's tiled
// transposition kernel on DEVICE buffers (block-shape parameters are accepted and ignored).
template <std::uint64_t... params>
struct TransposeTag {
  static void transpose(std::uint64_t* dst, const std::uint64_t* src, std::uint64_t src_rows, std::uint64_t src_cols,
                        std::uint64_t ld_dst, std::uint64_t ld_src, void* stream = nullptr) {
    const int status = xntt_transpose(dst, src, src_rows, src_cols, ld_dst, ld_src, stream);
    if (status == XNTT_ERR_INVALID) throw std::invalid_argument{"transpose: invalid shape"};
    if (status != XNTT_OK) throw std::runtime_error{xntt_last_cuda_error()};
  }
  static void transpose(std::uint64_t* dst, std::uint64_t dim, void* stream = nullptr) {
    transpose(dst, dst, dim, dim, dim, dim, stream);
  }
};
#define XNTT_TRANSPOSE_TAG(NAME)    \
  template <std::uint64_t... params> \
  using NAME = TransposeTag<params...>;
XNTT_TRANSPOSE_TAG(TransposeSVEInRegister)
XNTT_TRANSPOSE_TAG(TransposeParallelSVEInRegister)
XNTT_TRANSPOSE_TAG(TransposeSVEInRegisterRowFirst)
XNTT_TRANSPOSE_TAG(TransposeParallelSVEInRegisterRowFirst)
XNTT_TRANSPOSE_TAG(TransposeSVEInRegisterExplicitBlockingRowFirst)
XNTT_TRANSPOSE_TAG(TransposeParallelSVEInRegisterExplicitBlockingRowFirst)
XNTT_TRANSPOSE_TAG(TransposeSVEInRegisterFullBlockingRowFirst)
XNTT_TRANSPOSE_TAG(TransposeParallelSVEInRegisterFullBlockingRowFirst)
XNTT_TRANSPOSE_TAG(TransposeSVEGatherImmediateIndexRowFirst)
XNTT_TRANSPOSE_TAG(TransposeSVEGatherImmediateIndexColumnFirst)
XNTT_TRANSPOSE_TAG(TransposeSVEGatherVectorIndexRowFirst)
XNTT_TRANSPOSE_TAG(TransposeSVEGatherVectorIndexColumnFirst)
#undef XNTT_TRANSPOSE_TAG

}  // namespace sventt

#endif
