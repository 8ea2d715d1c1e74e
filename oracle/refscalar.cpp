// SPDX-License-Identifier: Apache-2.0
//
// TEST INFRASTRUCTURE / CPU BASELINE.  The reference's own "portable scalar path" - sventt::IterativeNTT over
// RadixEightScalarLayer<PAdic64Scalar> (include/sventt/layer/scalar/radix-eight.hpp:13-445, composition as in
// tests/ntt-tests/iterative-scalar-radix8-two12.hpp) - compiled from the reference headers where they lie
// (/root/reference/include is on the include path, nothing is copied) into oracle/_ref/libnttref_scalar.so.
//
// B2 = this kernel on one core, B3 = the same kernel under `omp parallel for` over the
// transforms of a batch on all host cores.  Restrictions the reference itself imposes:
//   * the scalar layers are only correct for moduli below 2^62 (lazy [0, 2N) arithmetic with + 2N corrections,
//     radix-eight.hpp:75-111), so this baseline runs at the 62-bit test prime 0x3a00000000000001 (g = 3) of
//     tests/ntt-tests/*.hpp, not at the production prime;
//   * sventt::NTT<> cannot wrap the scalar layers (their prepare_* take AuxiliaryVector& only, wrapper.hpp:18-22
//     sizes the arena through a FakeByteVector), so the kernel's static prepare_* / compute_* are driven directly;
//   * outputs are lazily reduced: callers compare `% N` like tests/bench-ntt.cpp:60-64 does.
#include <cstdint>
#include <cstring>
#include <utility>  // wrapper.hpp uses std::cmp_not_equal without including it

#include "sventt/sventt.hpp"

namespace {

using modulus_type = sventt::Modulus<UINT64_C(0x3a00000000000001), UINT64_C(3)>;
using modmul_type = sventt::PAdic64Scalar<modulus_type>;

template <std::uint64_t m, std::uint64_t n, std::uint64_t f = 1>
using R8 = sventt::RadixEightScalarLayer<modmul_type, m, n, f>;
template <std::uint64_t m, std::uint64_t n, std::uint64_t f = 1>
using R4 = sventt::RadixFourScalarLayer<modmul_type, m, n, f>;

constexpr std::uint64_t two(int l) { return std::uint64_t{1} << l; }

// 2^12 = 8^4 (the shipped test shape), 2^20 = 8^6 * 4, 2^24 = 8^8; the terminal layer carries inverse_factor = m
using K12 = sventt::IterativeNTT<modulus_type, two(12), R8<two(12), two(12)>, R8<two(12), two(9)>,
                                 R8<two(12), two(6)>, R8<two(12), two(3), two(12)>>;
using K20 = sventt::IterativeNTT<modulus_type, two(20), R8<two(20), two(20)>, R8<two(20), two(17)>,
                                 R8<two(20), two(14)>, R8<two(20), two(11)>, R8<two(20), two(8)>,
                                 R8<two(20), two(5)>, R4<two(20), two(2), two(20)>>;
using K24 = sventt::IterativeNTT<modulus_type, two(24), R8<two(24), two(24)>, R8<two(24), two(21)>,
                                 R8<two(24), two(18)>, R8<two(24), two(15)>, R8<two(24), two(12)>,
                                 R8<two(24), two(9)>, R8<two(24), two(6)>, R8<two(24), two(3), two(24)>>;

template <class K>
struct Runner {
  sventt::AuxiliaryVector fwd{4096}, inv{4096};
  Runner() {
    K::prepare_forward(fwd);
    K::prepare_inverse(inv);
  }
  void forward(std::uint64_t* dst, const std::uint64_t* src) const {
    const std::byte* aux = fwd.data();
    K::compute_forward(dst, src, aux);
  }
  void inverse(std::uint64_t* dst, const std::uint64_t* src) const {
    const std::byte* aux = inv.data();
    K::compute_inverse(dst, src, aux);
  }
};

template <class K>
int run(int inverse, std::uint64_t* dst, const std::uint64_t* src, std::uint64_t batch, int threads) {
  static const Runner<K> r;
  const std::uint64_t m = K::get_m();
  if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic) num_threads(threads) if (threads > 1 && batch > 1)
  for (std::uint64_t b = 0; b < batch; ++b) {
    if (inverse)
      r.inverse(dst + b * m, src + b * m);
    else
      r.forward(dst + b * m, src + b * m);
  }
  return 0;
}

}  // namespace

extern "C" {
std::uint64_t refscalar_modulus(void) { return modulus_type::get_modulus(); }
std::uint64_t refscalar_generator(void) { return 3; }
// batch back-to-back transforms of length 2^log2_m (12, 20 or 24) on `threads` OpenMP threads; out of place or in place
int refscalar_run(int log2_m, int inverse, std::uint64_t* dst, const std::uint64_t* src, std::uint64_t batch,
                  int threads) {
  switch (log2_m) {
    case 12:
      return run<K12>(inverse, dst, src, batch, threads);
    case 20:
      return run<K20>(inverse, dst, src, batch, threads);
    case 24:
      return run<K24>(inverse, dst, src, batch, threads);
    default:
      return -1;
  }
}
}
