// SPDX-License-Identifier: Apache-2.0
//
// The reference's benchmark-as-test (tests/bench-ntt.cpp:20-65) re-hosted on the drop-in headers:
// for every kernel composition of tests/ntt-tests/*.hpp (and the README example) build
// sventt::NTT<kernel_type>, transform an iota input out of place with dst poisoned, and require
// dst[i] % N == reference[i] for all i, forward and inverse.  The compositions are spelled exactly
// as in the reference, including its 62-bit test modulus 0x3a00000000000001 (README example and the
// 2^24 shape: the production modulus of README.md:19).  The expected vectors come from the CPU
// oracle (oracle/ntt_oracle.c) - this file is test code, the only place allowed to link it.
#include <sventt/sventt.hpp>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

extern "C" {
void oracle_ntt_forward(std::uint64_t* dst, const std::uint64_t* src, std::uint64_t m, std::uint64_t N,
                        std::uint64_t g);
void oracle_ntt_inverse(std::uint64_t* dst, const std::uint64_t* src, std::uint64_t m, std::uint64_t N,
                        std::uint64_t g);
}

using namespace sventt;

constexpr std::uint64_t one = 1;
namespace prod {
using modulus_type = Modulus<UINT64_C(0xfffffc6e80000001), UINT64_C(3)>;  // README.md:19
using modmul_type = PAdic64SVE<modulus_type>;
}  // namespace prod
// tests/ntt-tests/*.hpp
using modulus_type = Modulus<UINT64_C(0x3a00000000000001), UINT64_C(3)>;
using modmul_type = PAdic64SVE<modulus_type>;

// ---- tests/ntt-tests/iterative-sve-radix2-two10.hpp
namespace it_r2 {
constexpr std::uint64_t m{one << 10};
template <std::uint64_t n, std::uint64_t f = 1>
using L = RadixTwoSVELayer<modmul_type, m, n, f>;
using kernel_type = IterativeNTT<modulus_type, m, L<one << 10>, L<one << 9>, L<one << 8>, L<one << 7>, L<one << 6>,
                                 L<one << 5>, L<one << 4>, L<one << 3>, L<one << 2>, L<one << 1, m>>;
}  // namespace it_r2
// ---- tests/ntt-tests/iterative-scalar-radix2-two10.hpp shape: PAdic64 and FixedPoint64 layers mixed
namespace it_mixed {
constexpr std::uint64_t m{one << 10};
using fp = FixedPoint64Scalar<modulus_type>;
using pa = PAdic64Scalar<modulus_type>;
using kernel_type =
    IterativeNTT<modulus_type, m, RadixTwoScalarLayer<pa, m, one << 10>, RadixTwoScalarLayer<fp, m, one << 9>,
                 RadixFourScalarLayer<pa, m, one << 8>, RadixFourScalarLayer<fp, m, one << 6>,
                 RadixEightScalarLayer<pa, m, one << 4>, RadixTwoScalarLayer<fp, m, one << 1, m>>;
}  // namespace it_mixed
// ---- every layer FixedPoint64 (modmul/sve/fixed-point-64.hpp): the device runs the Shoup kernels
namespace it_fixed {
constexpr std::uint64_t m{one << 12};
using fp = FixedPoint64SVE<modulus_type>;
using kernel_type = IterativeNTT<modulus_type, m, RadixEightSVELayer<fp, m, one << 12>, RadixEightSVELayer<fp, m, one << 9>,
                                 RadixEightSVELayer<fp, m, one << 6>, RadixEightSVELayer<fp, m, one << 3, m>>;
}  // namespace it_fixed
namespace rec_fixed {  // four-step 2^9 x 2^6 with FixedPoint64 layers
constexpr std::uint64_t m{one << 15}, n0{one << 9}, n1{one << 6};
using fp = FixedPoint64SVE<modulus_type>;
using inner0 = IterativeNTT<modulus_type, n0, RadixEightSVELayer<fp, n0, n0>, RadixEightSVELayer<fp, n0, (n0 >> 3)>,
                            RadixEightSVELayer<fp, n0, (n0 >> 6)>>;
using inner1 = IterativeNTT<modulus_type, n1, RadixEightSVELayer<fp, n1, n1>, RadixEightSVELayer<fp, n1, (n1 >> 3), m>>;
using kernel_type = RecursiveNTT<modulus_type, m, GenericSVELayer<fp, m, inner0, 0, 1>, inner1, true>;
}  // namespace rec_fixed
// ---- tests/ntt-tests/iterative-sve-radix4-two12.hpp
namespace it_r4 {
constexpr std::uint64_t m{one << 12};
template <std::uint64_t n, std::uint64_t f = 1>
using L = RadixFourSVELayer<modmul_type, m, n, f>;
using kernel_type = IterativeNTT<modulus_type, m, L<one << 12>, L<one << 10>, L<one << 8>, L<one << 6>, L<one << 4>,
                                 L<one << 2, m>>;
}  // namespace it_r4
// ---- tests/ntt-tests/iterative-sve-radix8-two12.hpp
namespace it_r8 {
constexpr std::uint64_t m{one << 12};
using kernel_type =
    IterativeNTT<modulus_type, m, RadixEightSVELayer<modmul_type, m, one << 12, 1, false>,
                 RadixEightSVELayer<modmul_type, m, one << 9, 1, true>,
                 RadixEightSVELayer<modmul_type, m, one << 6, 1, false>, RadixEightSVELayer<modmul_type, m, one << 3, m, true>>;
}  // namespace it_r8
// ---- tests/ntt-tests/iterative-scalar-radix248-two13.hpp shape: mixed radix 8*4*2*8*... = 2^13
namespace it_r248 {
constexpr std::uint64_t m{one << 13};
using kernel_type =
    IterativeNTT<modulus_type, m, RadixEightScalarLayer<modmul_type, m, one << 13>,
                 RadixFourScalarLayer<modmul_type, m, one << 10>, RadixTwoScalarLayer<modmul_type, m, one << 8>,
                 RadixEightScalarLayer<modmul_type, m, one << 7>, RadixTwoScalarLayer<modmul_type, m, one << 4>,
                 RadixEightScalarLayer<modmul_type, m, one << 3, m>>;
}  // namespace it_r248
// ---- tests/ntt-tests/recursive-sve-radix248-two13.hpp
namespace rec_r248 {
constexpr std::uint64_t m{one << 13};
using inner_inner_kernel_type =
    IterativeNTT<modulus_type, one << 9, RadixEightSVELayer<modmul_type, one << 9, one << 9, 1, true>,
                 RadixFourSVELayer<modmul_type, one << 9, one << 6, 1, true>,
                 RadixTwoSVELayer<modmul_type, one << 9, one << 4, 1, false>,
                 RadixEightSVELayer<modmul_type, one << 9, one << 3, m, false>>;
using inner_kernel_type = RecursiveNTT<modulus_type, one << 10,
                                       RadixTwoSVELayer<modmul_type, one << 10, one << 10, 1, false>,
                                       inner_inner_kernel_type, false>;
using kernel_type =
    RecursiveNTT<modulus_type, m, RadixEightSVELayer<modmul_type, m, m, 1, true>, inner_kernel_type, false>;
}  // namespace rec_r248
// ---- tests/ntt-tests/recursive-sve-fourstep-two13.hpp (m = 2^15 = 2^9 x 2^6, six-step)
namespace rec_four {
constexpr std::uint64_t m{one << 15};
using inner_column_kernel_type =
    IterativeNTT<modulus_type, one << 9, RadixEightSVELayer<modmul_type, one << 9, one << 9>,
                 RadixFourSVELayer<modmul_type, one << 9, one << 6>, RadixTwoSVELayer<modmul_type, one << 9, one << 4>,
                 RadixEightSVELayer<modmul_type, one << 9, one << 3, m>>;
using inner_row_kernel_type = IterativeNTT<modulus_type, one << 6, RadixEightSVELayer<modmul_type, one << 6, one << 6>,
                                           RadixEightSVELayer<modmul_type, one << 6, one << 3>>;
using kernel_type = RecursiveNTT<modulus_type, m,
                                 GenericSVELayer<modmul_type, m, inner_column_kernel_type, 8, 2,
                                                 TransposeParallelSVEInRegister<8, 64>>,
                                 inner_row_kernel_type, true>;
}  // namespace rec_four
// ---- tests/ntt-tests/recursive-scalar-fourstep-two13.hpp (m = 2^13 = 2^9 x 2^4)
namespace rec_four13 {
constexpr std::uint64_t m{one << 13};
using inner_column_kernel_type =
    IterativeNTT<modulus_type, one << 9, RadixEightScalarLayer<modmul_type, one << 9, one << 9>,
                 RadixFourScalarLayer<modmul_type, one << 9, one << 6>,
                 RadixTwoScalarLayer<modmul_type, one << 9, one << 4>,
                 RadixEightScalarLayer<modmul_type, one << 9, one << 3, m>>;
using inner_row_kernel_type = IterativeNTT<modulus_type, one << 4, RadixFourScalarLayer<modmul_type, one << 4, one << 4>,
                                           RadixFourScalarLayer<modmul_type, one << 4, one << 2>>;
using kernel_type = RecursiveNTT<modulus_type, m, GenericScalarLayer<modmul_type, m, inner_column_kernel_type>,
                                 inner_row_kernel_type, true>;
}  // namespace rec_four13
// ---- README.md:14-71: blocked six-step, 2^17 = 2^8 x 2^9, unscaled inverse
namespace readme {
using modulus_type = prod::modulus_type;
using modmul_type = prod::modmul_type;
using transposition_type = TransposeParallelSVEInRegisterExplicitBlockingRowFirst<32, 128, 128 + 32, 3>;
constexpr std::uint64_t n = one << 17, n0 = one << 8, n1 = one << 9;
using ntt0_type = IterativeNTT<modulus_type, n0, RadixEightSVELayer<modmul_type, n0, n0>,
                               RadixEightSVELayer<modmul_type, n0, (n0 >> 3)>, RadixFourSVELayer<modmul_type, n0, (n0 >> 6)>>;
using ntt1_type = RecursiveNTT<modulus_type, n1, RadixEightSVELayer<modmul_type, n1, n1>,
                               IterativeNTT<modulus_type, (n1 >> 3), RadixEightSVELayer<modmul_type, (n1 >> 3), (n1 >> 3)>,
                                            RadixEightSVELayer<modmul_type, (n1 >> 3), (n1 >> 6)>>,
                               false>;
using kernel_type =
    RecursiveNTT<modulus_type, n, BlockedGenericSVELayer<modmul_type, n, ntt0_type, 32, 2, 128, transposition_type>,
                 ntt1_type, true>;
}  // namespace readme
// ---- BASELINE configs[1] spelled with the reference's classes: 2^24 = 2^12 x 2^12 blocked six-step
namespace big24 {
using modulus_type = prod::modulus_type;
using modmul_type = prod::modmul_type;
constexpr std::uint64_t n = one << 24, n0 = one << 12, n1 = one << 12;
template <std::uint64_t len, std::uint64_t f>
using inner = IterativeNTT<modulus_type, len, RadixEightSVELayer<modmul_type, len, len>,
                           RadixEightSVELayer<modmul_type, len, (len >> 3)>, RadixEightSVELayer<modmul_type, len, (len >> 6)>,
                           RadixEightSVELayer<modmul_type, len, (len >> 9), f>>;
using kernel_type = RecursiveNTT<modulus_type, n, BlockedGenericSVELayer<modmul_type, n, inner<n0, 1>, 32, 2, 128>,
                                 inner<n1, n>, true>;
}  // namespace big24

// ---- one transform over several GPUs: production modulus, six-step compositions (2^10 x 2^10 and the README shape)
namespace mg20 {
using modulus_type = prod::modulus_type;
using modmul_type = prod::modmul_type;
constexpr std::uint64_t n = one << 20, n0 = one << 10, n1 = one << 10;
template <std::uint64_t len, std::uint64_t f>
using inner = IterativeNTT<modulus_type, len, RadixTwoSVELayer<modmul_type, len, len>,
                           RadixEightSVELayer<modmul_type, len, (len >> 1)>, RadixEightSVELayer<modmul_type, len, (len >> 4)>,
                           RadixEightSVELayer<modmul_type, len, (len >> 7), f>>;
using kernel_type = RecursiveNTT<modulus_type, n, BlockedGenericSVELayer<modmul_type, n, inner<n0, 1>, 32, 2, 128>,
                                 inner<n1, n>, true>;
}  // namespace mg20
namespace mg17 {
using modulus_type = prod::modulus_type;
using modmul_type = prod::modmul_type;
constexpr std::uint64_t n = one << 17, n0 = one << 8, n1 = one << 9;
using ntt0_type = IterativeNTT<modulus_type, n0, RadixEightSVELayer<modmul_type, n0, n0>,
                               RadixEightSVELayer<modmul_type, n0, (n0 >> 3)>, RadixFourSVELayer<modmul_type, n0, (n0 >> 6)>>;
using ntt1_type = IterativeNTT<modulus_type, n1, RadixEightSVELayer<modmul_type, n1, n1>,
                               RadixEightSVELayer<modmul_type, n1, (n1 >> 3)>, RadixEightSVELayer<modmul_type, n1, (n1 >> 6), n>>;
using kernel_type =
    RecursiveNTT<modulus_type, n, BlockedGenericSVELayer<modmul_type, n, ntt0_type, 32, 2, 128>, ntt1_type, true>;
}  // namespace mg17

static int failures = 0;

// bench-ntt.cpp:20-65 without the Google Benchmark loop
template <class kernel_type, bool is_inverse>
static void check_ntt(const std::string& name) {
  using ntt_type = NTT<kernel_type>;
  using mod = typename ntt_type::modulus_type;
  const std::uint64_t m{ntt_type::get_m()};
  const std::uint64_t N{mod::get_modulus()};
  std::default_random_engine gen;
  PageMemory<std::uint64_t> buffer{m * 3, false};
  std::uint64_t *src{&buffer[m * 0]}, *dst{&buffer[m * 1]}, *dst_ref{&buffer[m * 2]};
  std::iota(&src[0], &src[m], std::uniform_int_distribution<std::uint64_t>{std::uint64_t{}, N - m - 1}(gen));
  std::memset(dst, 0x55, sizeof(std::uint64_t) * m);
  std::memset(dst_ref, 0xaa, sizeof(std::uint64_t) * m);
  if constexpr (is_inverse) {
    oracle_ntt_inverse(dst_ref, src, m, N, mod::get_generator());
    // NTTReference always scales by 1/m; a composition without inverse_factor does not
    const std::uint64_t k = mod::multiply(mod::invert(kernel_type::get_inverse_factor()), m % N);
    for (std::uint64_t i = 0; i < m; ++i) dst_ref[i] = mod::multiply(dst_ref[i], k);
  } else {
    oracle_ntt_forward(dst_ref, src, m, N, mod::get_generator());
  }
  const ntt_type ntt{!is_inverse, is_inverse, false};
  if constexpr (is_inverse)
    ntt.compute_inverse(dst, src);
  else
    ntt.compute_forward(dst, src);
  std::uint64_t bad = 0;
  for (std::uint64_t i = 0; i < m; ++i) bad += (dst[i] % N != dst_ref[i]);
  // the in-place overload on the same data must agree
  std::memcpy(dst_ref, src, sizeof(std::uint64_t) * m);
  if constexpr (is_inverse)
    ntt.compute_inverse(dst_ref);
  else
    ntt.compute_forward(dst_ref);
  bad += std::memcmp(dst, dst_ref, sizeof(std::uint64_t) * m) != 0;
  // a direction that was not enabled is a logic_error (wrapper.hpp:54-56)
  bool threw = false;
  try {
    if constexpr (is_inverse)
      ntt.compute_forward(dst_ref);
    else
      ntt.compute_inverse(dst_ref);
  } catch (const std::logic_error&) {
    threw = true;
  }
  bad += !threw;
  std::printf("%-8s %-44s m=2^%-2d %s\n", is_inverse ? "Inverse," : "Forward,", name.c_str(),
              (int)detail::log2_exact(m), bad ? "MISMATCH" : "ok");
  failures += bad != 0;
}

template <class kernel_type>
static void both(const std::string& name) {
  check_ntt<kernel_type, false>(name);
  check_ntt<kernel_type, true>(name);
}

int main(int argc, char** argv) {
  const bool big = argc > 1 && std::string{argv[1]} == "--big";
  try {
    both<it_r2::kernel_type>("iterative, SVE, radix-2");
    both<it_mixed::kernel_type>("iterative, scalar, mixed modmul tags");
    both<it_fixed::kernel_type>("iterative, FixedPoint64 layers (Shoup kernels)");
    both<rec_fixed::kernel_type>("four-step, FixedPoint64 layers (Shoup kernels)");
    {
      static_assert(it_fixed::kernel_type::all_fixed_point() && rec_fixed::kernel_type::all_fixed_point());
      static_assert(!it_mixed::kernel_type::all_fixed_point() && !it_r8::kernel_type::all_fixed_point());
      const NTT<it_fixed::kernel_type> shoup;
      const NTT<it_mixed::kernel_type> mixed;
      const bool tag_ok = shoup.get_modmul_kind() == 1 && mixed.get_modmul_kind() == 0;
      std::printf("%-53s %s\n", "FixedPoint64 tag selects the Shoup kernels", tag_ok ? "ok" : "MISMATCH");
      failures += !tag_ok;
    }
    both<it_r4::kernel_type>("iterative, SVE, radix-4");
    both<it_r8::kernel_type>("iterative, SVE, radix-8");
    both<it_r248::kernel_type>("iterative, scalar, radix-2,4,8");
    both<rec_r248::kernel_type>("recursive, SVE, radix-2,4,8");
    both<rec_four::kernel_type>("recursive, SVE, four-step");
    both<rec_four13::kernel_type>("recursive, scalar, four-step");
    both<readme::kernel_type>("README blocked six-step 2^8 x 2^9");
    if (big) both<big24::kernel_type>("blocked six-step 2^12 x 2^12");
    // the same user code, one transform spread over several GPUs by the library (xntt_mgpu_*): by default every
    // rank on device 0, `--devices 0,1,2,3` names real ones
    {
      std::vector<int> devs{0, 0, 0, 0};
      for (int i = 1; i + 1 < argc; ++i)
        if (std::string{argv[i]} == "--devices") {
          devs.clear();
          for (const char* p = argv[i + 1]; *p; ++p)
            if (*p >= '0' && *p <= '9') devs.push_back(*p - '0');
        }
      set_default_devices(devs);
      both<mg20::kernel_type>("multi-GPU six-step 2^10 x 2^10, G=" + std::to_string(devs.size()));
      set_default_devices(std::vector<int>(devs.begin(), devs.begin() + 2));
      both<mg17::kernel_type>("multi-GPU README shape 2^8 x 2^9, G=2");
      if (big) {
        set_default_devices(devs);
        both<big24::kernel_type>("multi-GPU six-step 2^12 x 2^12, G=" + std::to_string(devs.size()));
      }
      set_default_devices({});
    }
    // Modulus / PAdic64 scalar identities used by the examples
    using modulus_type = prod::modulus_type;
    using P = PAdic64<modulus_type>;
    static_assert(modulus_type::get_montgomery_inverse() == UINT64_C(0x4000039180000001));
    static_assert(P::to_montgomery(1) == UINT64_C(0x3917fffffff));
    static_assert(P::from_montgomery(P::to_montgomery(12345)) == 12345);
    using FP = FixedPoint64<modulus_type>;
    static_assert(FP::multiply(0xfffffc6e80000000ull, 0xfffffc6e80000000ull) == 1);  // (-1)^2
    static_assert(FP::multiply(P::to_montgomery(7), 3) == modulus_type::multiply(P::to_montgomery(7), 3));
    static_assert(modulus_type::multiply(modulus_type::get_root_forward(one << 31), modulus_type::get_root_inverse(one << 31)) == 1);
    bool threw = false;
    try {
      (void)modulus_type::get_root_forward(7);  // 7 does not divide p - 1
    } catch (const std::invalid_argument&) {
      threw = true;
    }
    failures += !threw;
  } catch (const std::exception& e) {
    std::printf("exception: %s\n", e.what());
    return 2;
  }
  std::printf("%s\n", failures ? "FAILED" : "ALL OK");
  return failures ? 1 : 0;
}
