// SPDX-License-Identifier: Apache-2.0
//
// The reference's benchmark harness (tests/bench-ntt.cpp:20-84) re-hosted on the drop-in headers, timing only: for a
// kernel composition it registers "Forward, <name>" and "Inverse, <name>", runs compute_forward / compute_inverse out of
// place and reports time per transform and items per second (Google Benchmark's SetItemsProcessed, bench-ntt.cpp:57).
// Two columns: device-resident buffers (DeviceMemory, stream-ordered calls - what the kernels do) and host buffers
// (PageMemory, the reference user's call, PCIe copies included).  Correctness lives in ntt_tests.cpp.
//   usage: bench_ntt [--reps N] [--devices 0,1,...]      (devices: also run the multi-GPU line)
#include <sventt/sventt.hpp>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <numeric>
#include <string>
#include <vector>

using namespace sventt;
constexpr std::uint64_t one = 1;
using modulus_type = Modulus<UINT64_C(0xfffffc6e80000001), UINT64_C(3)>;  // README.md:19
using modmul_type = PAdic64SVE<modulus_type>;

template <std::uint64_t len, std::uint64_t f>
using radix8x4 = IterativeNTT<modulus_type, len, RadixEightSVELayer<modmul_type, len, len>,
                              RadixEightSVELayer<modmul_type, len, (len >> 3)>, RadixEightSVELayer<modmul_type, len, (len >> 6)>,
                              RadixEightSVELayer<modmul_type, len, (len >> 9), f>>;
// README.md:28-68 shape, 2^17 = 2^8 x 2^9
namespace readme {
constexpr std::uint64_t n = one << 17, n0 = one << 8, n1 = one << 9;
using ntt0 = IterativeNTT<modulus_type, n0, RadixEightSVELayer<modmul_type, n0, n0>,
                          RadixEightSVELayer<modmul_type, n0, (n0 >> 3)>, RadixFourSVELayer<modmul_type, n0, (n0 >> 6)>>;
using ntt1 = IterativeNTT<modulus_type, n1, RadixEightSVELayer<modmul_type, n1, n1>,
                          RadixEightSVELayer<modmul_type, n1, (n1 >> 3)>, RadixEightSVELayer<modmul_type, n1, (n1 >> 6), n>>;
using kernel_type = RecursiveNTT<modulus_type, n, BlockedGenericSVELayer<modmul_type, n, ntt0, 32, 2, 128>, ntt1, true>;
}  // namespace readme
// BASELINE configs[1]: 2^24 = 2^12 x 2^12 blocked six-step
namespace big24 {
constexpr std::uint64_t n = one << 24, n0 = one << 12, n1 = one << 12;
using kernel_type = RecursiveNTT<modulus_type, n, BlockedGenericSVELayer<modmul_type, n, radix8x4<n0, 1>, 32, 2, 128>,
                                 radix8x4<n1, n>, true>;
}  // namespace big24

static double seconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <class kernel_type, bool is_inverse>
static void bench(const std::string& name, int reps) {
  using ntt_type = NTT<kernel_type>;
  const std::uint64_t m = ntt_type::get_m();
  const ntt_type ntt{!is_inverse, is_inverse, false};
  PageMemory<std::uint64_t> host{m * 2, false};
  std::iota(&host[0], &host[m], std::uint64_t{12345});
  double dev_s = 0;
  if (ntt.get_device_count() == 1) {
    DeviceMemory<std::uint64_t> src{m}, dst{m};
    src.copy_from_host(host.data());
    for (int w = 0; w < 3; ++w)
      is_inverse ? ntt.compute_inverse_async(dst.data(), src.data(), nullptr)
                 : ntt.compute_forward_async(dst.data(), src.data(), nullptr);
    ntt.synchronize();
    const double t0 = seconds();
    for (int r = 0; r < reps; ++r)
      is_inverse ? ntt.compute_inverse_async(dst.data(), src.data(), nullptr)
                 : ntt.compute_forward_async(dst.data(), src.data(), nullptr);
    ntt.synchronize();
    dev_s = (seconds() - t0) / reps;
  }
  for (int w = 0; w < 2; ++w)
    is_inverse ? ntt.compute_inverse(&host[m], &host[0]) : ntt.compute_forward(&host[m], &host[0]);
  const int hreps = reps < 5 ? reps : 5;
  const double t0 = seconds();
  for (int r = 0; r < hreps; ++r)
    is_inverse ? ntt.compute_inverse(&host[m], &host[0]) : ntt.compute_forward(&host[m], &host[0]);
  const double host_s = (seconds() - t0) / hreps;
  std::printf("%-8s %-46s m=2^%-2u gpus=%u  device %10.1f us %8.2f Gitems/s   host buffers %10.1f us %7.2f Gitems/s\n",
              is_inverse ? "Inverse," : "Forward,", name.c_str(), detail::log2_exact(m), ntt.get_device_count(),
              dev_s * 1e6, dev_s > 0 ? m / dev_s / 1e9 : 0.0, host_s * 1e6, m / host_s / 1e9);
}

int main(int argc, char** argv) {
  int reps = 50;
  std::vector<int> devs;
  for (int i = 1; i + 1 < argc; ++i) {
    if (std::string{argv[i]} == "--reps") reps = std::atoi(argv[i + 1]);
    if (std::string{argv[i]} == "--devices")
      for (const char* p = argv[i + 1]; *p; ++p)
        if (*p >= '0' && *p <= '9') devs.push_back(*p - '0');
  }
  try {
    bench<readme::kernel_type, false>("README blocked six-step 2^8 x 2^9", reps);
    bench<readme::kernel_type, true>("README blocked six-step 2^8 x 2^9", reps);
    bench<big24::kernel_type, false>("blocked six-step 2^12 x 2^12", reps);
    bench<big24::kernel_type, true>("blocked six-step 2^12 x 2^12", reps);
    if (devs.size() >= 2) {
      set_default_devices(devs);
      bench<big24::kernel_type, false>("blocked six-step 2^12 x 2^12, multi-GPU", reps);
      bench<big24::kernel_type, true>("blocked six-step 2^12 x 2^12, multi-GPU", reps);
      set_default_devices({});
    }
  } catch (const std::exception& e) {
    std::printf("exception: %s\n", e.what());
    return 2;
  }
  return 0;
}
