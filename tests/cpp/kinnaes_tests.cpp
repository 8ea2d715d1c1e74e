// SPDX-License-Identifier: Apache-2.0
// The reference's test-magic-series-kinnaes.cpp (typed test over (m, Modulus<N, g>, n), expected counts reduced
// modulo N) against the drop-in MagicSeriesKinnaes; the expected residues come from the oracle library, which
// tests/test_kinnaes.py pins on the reference's decimal strings.  Test code: links the oracle.
#include <cstdint>
#include <cstdio>

#include <sventt/sventt.hpp>

#include "examples/magic-series-kinnaes/kinnaes.hpp"

extern "C" std::uint64_t oracle_kinnaes_compute(std::uint64_t N, std::uint64_t g, std::uint64_t m, std::uint64_t n);
extern "C" std::uint64_t oracle_kinnaes_sum(std::uint64_t N, std::uint64_t g, std::uint64_t m, std::uint64_t n,
                                            std::uint64_t jb, std::uint64_t je);

static int failures = 0;

template <std::uint64_t m, std::uint64_t N, std::uint64_t g, std::uint64_t n>
static void run_case() {
  using modulus_type = sventt::Modulus<N, g>;
  using modmul_type = sventt::PAdic64SVE<modulus_type>;
  using kinnaes_type = MagicSeriesKinnaes<m, modmul_type, n>;
  static_assert(kinnaes_type::get_m() == m && kinnaes_type::get_n() == n && kinnaes_type::get_r() == m * (m - 1) / 2 * m);
  const std::uint64_t got = kinnaes_type::compute(), want = oracle_kinnaes_compute(N, g, m, n);
  const std::uint64_t part = kinnaes_type::compute_sum(3, 40), part_want = oracle_kinnaes_sum(N, g, m, n, 3, 40);
  const bool ok = got == want && part == part_want;
  std::printf("kinnaes m=%llu N=%016llx n=%llu %s\n", (unsigned long long)m, (unsigned long long)N,
              (unsigned long long)n, ok ? "ok" : "MISMATCH");
  if (!ok) ++failures;
}

int main(int argc, char**) {
  // test-magic-series-kinnaes.cpp:18-67 (a 64-bit and a 61-bit modulus of each order; all twelve with any argument)
  run_case<100, UINT64_C(0xfffffffffeca467f), 5, 495017>();
  run_case<101, UINT64_C(0x1ffffffffdce2e99), 3, 510053>();
  if (argc > 1) {
    run_case<100, UINT64_C(0xfffffffffe05e355), 6, 495017>();
    run_case<100, UINT64_C(0x7ffffffffed59fb5), 2, 495017>();
    run_case<100, UINT64_C(0x7ffffffffcc4e37f), 13, 495017>();
    run_case<100, UINT64_C(0x3fffffffff4c9937), 5, 495017>();
    run_case<100, UINT64_C(0x3fffffffff102bef), 3, 495017>();
    run_case<100, UINT64_C(0x1ffffffffff962df), 7, 495017>();
    run_case<100, UINT64_C(0x1fffffffffdb2c3b), 2, 495017>();
    run_case<101, UINT64_C(0xfffffffffe023ec1), 11, 510053>();
    run_case<101, UINT64_C(0x7ffffffffd0f0621), 3, 510053>();
    run_case<101, UINT64_C(0x3ffffffffec5c639), 21, 510053>();
  }
  std::printf(failures ? "FAILED\n" : "ALL OK\n");
  return failures ? 1 : 0;
}
