// SPDX-License-Identifier: Apache-2.0
// The reference's examples/magic-series/test-magic-series.cpp (GaussianPolynomialCoefficient, :295-331) against the
// drop-in headers: the 2^15 IterativeNTT of five radix-8 layers spelled as in the reference, its eight moduli
// (:22-39), the magic-series counts it stores as decimal strings (:315-325) reduced modulo each prime.
#include <cstdint>
#include <cstdio>
#include <string>
#include <tuple>
#include <vector>

#include <sventt/sventt.hpp>

#include "examples/magic-series/gaussian-polynomial.hpp"

static int failures = 0;

template <class modulus_type>
static std::uint64_t decimal_mod(const std::string& s) {
  unsigned __int128 r = 0;
  for (char ch : s) r = (r * 10 + (unsigned)(ch - '0')) % modulus_type::get_modulus();
  return (std::uint64_t)r;
}

template <std::uint64_t N, std::uint64_t g>
static void run_modulus(bool all) {
  using modulus_type = sventt::Modulus<N, g>;
  using modmul_type = sventt::PAdic64SVE<modulus_type>;
  constexpr std::uint64_t len{std::uint64_t{1} << 15};
  using kernel_type = sventt::IterativeNTT<
      modulus_type, len, sventt::RadixEightSVELayer<modmul_type, len, std::uint64_t{1} << 15>,
      sventt::RadixEightSVELayer<modmul_type, len, std::uint64_t{1} << 12>,
      sventt::RadixEightSVELayer<modmul_type, len, std::uint64_t{1} << 9>,
      sventt::RadixEightSVELayer<modmul_type, len, std::uint64_t{1} << 6>,
      sventt::RadixEightSVELayer<modmul_type, len, std::uint64_t{1} << 3, len>>;
  sventt::NTT<kernel_type> ntt{true, true, false};
  const std::vector<std::tuple<std::uint64_t, std::string>> expected{
      {10, "78132541528"},
      {25, "140170526450793924490478768121814869629364"},
      {35, "13872534241478210358349096341203128450357241660871429860873721318"},
      {42, "1195452957914568544628242649935060977711193839443701120065551521757686130217168310"},
      {100, "904300736808894426574793302240693911261234942398748154528052171724"
            "305279045583459861011357813556260746366850646669062169890178280824"
            "885995375485156399921958991796250954308603011799192842071430359668"
            "946052264146938445899732873114858199920"},
  };
  for (const auto& [m, text] : expected) {
    if (!all && m > 35) continue;
    const std::uint64_t got = calculate_number_of_magic_series(m, ntt), want = decimal_mod<modulus_type>(text);
    const bool ok = got == want;
    std::printf("magic series m=%llu N=%016llx %s\n", (unsigned long long)m, (unsigned long long)N, ok ? "ok" : "MISMATCH");
    if (!ok) ++failures;
  }
}

// the helper classes on their own (test-magic-series.cpp: QPochhammer, RestrictedPartition, numerator segments)
static void run_units() {
  using modulus_type = sventt::Modulus<UINT64_C(0xffffffff00000001), 7>;
  constexpr std::uint64_t N = modulus_type::get_modulus();
  bool ok = true;
  {  // (1-q)(1-q^2)(1-q^3) = 1 - q - q^2 + q^4 + q^5 - q^6
    std::vector<std::uint64_t> c(7, 99);
    calculate_q_pochhammer<modulus_type>(c, 3);
    const std::vector<std::uint64_t> want{1, N - 1, N - 1, 0, 1, 1, N - 1};
    ok = ok && c == want;
  }
  {  // partitions of n into parts <= 3: 1 1 2 3 4 5 7 8 10 12 14
    RestrictedPartition<modulus_type> p(3);
    const std::uint64_t want[] = {1, 1, 2, 3, 4, 5, 7, 8, 10, 12, 14};
    for (std::uint64_t w : want) {
      ok = ok && p() == w;
      p.advance();
    }
    ok = ok && p.get_n() == 11 && p.get_k() == 3;
  }
  {  // [4 choose 2]_q = 1 + q + 2 q^2 + q^3 + q^4
    GaussianPolynomialNumeratorSegment<modulus_type> seg(4);
    seg.advance();  // [4 0]
    seg.advance();  // [4 1]
    seg.advance();  // [4 2]
    ok = ok && seg.get_coefficients() == std::vector<std::uint64_t>{1, 1, 2, 1, 1};
  }
  {  // prod_{i=1..2} (1 - q^(5-2+i)) = (1 - q^4)(1 - q^5) = 1 - q^4 - q^5 + q^9, subtracted from zeros
    GaussianPolynomialNumerator<modulus_type> num(5, 2);
    std::vector<std::uint64_t> v(12, 0);
    num.subtract_next(v.data(), 5);
    num.subtract_next(v.data() + 5, 7);
    const std::vector<std::uint64_t> want{N - 1, 0, 0, 0, 1, 1, 0, 0, 0, N - 1, 0, 0};
    ok = ok && v == want;
  }
  std::printf("magic series helper classes %s\n", ok ? "ok" : "MISMATCH");
  if (!ok) ++failures;
}

int main(int argc, char**) {
  const bool all = argc > 1;
  run_units();
  run_modulus<UINT64_C(0xffffffff00000001), 7>(all);
  run_modulus<UINT64_C(0xa3b25f400c7a8001), 5>(all);
  if (all) {
    run_modulus<UINT64_C(0xffffffff00000001), UINT64_C(0xf44872f5ec1c4cc0)>(all);
    run_modulus<UINT64_C(0x41d33d0d1fbf8001), 6>(all);
    run_modulus<UINT64_C(0x3164c5d59b090001), 13>(all);
    run_modulus<UINT64_C(0x1e4a0e19e4548001), 3>(all);
    run_modulus<UINT64_C(0x08aa90297f870001), 3>(all);
    run_modulus<UINT64_C(0x0000000000010001), 3>(all);
  }
  std::printf(failures ? "FAILED\n" : "ALL OK\n");
  return failures ? 1 : 0;
}
