// SPDX-License-Identifier: Apache-2.0
//
// sventt::NTT<kernel_type> - the user-facing wrapper of the reference
// (include/sventt/wrapper.hpp:13-83): constructor (enable_forward, enable_inverse,
// allocate_huge_pages), get_m(), compute_forward(dst[, src]), compute_inverse(dst[, src]), all const.
// Construction builds the xntt plan (twiddles are generated on the device); the compute calls accept
// either ordinary host pointers (copied over PCIe, like a drop-in user would have them) or device
// pointers (used in place).  Errors of the C ABI come back as the exception types the reference
// throws: invalid_argument / bad_alloc / logic_error / runtime_error.
#ifndef XNTT_SVENTT_WRAPPER_HPP
#define XNTT_SVENTT_WRAPPER_HPP

#include <cstdint>
#include <new>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "kernel.hpp"
#include "xntt.h"

namespace sventt {

namespace detail {
[[noreturn]] inline void throw_status(int status, const char* what) {
  const std::string msg = std::string{what} + ": " + xntt_strerror(status);
  switch (status) {
    case XNTT_ERR_INVALID:
    case XNTT_ERR_UNSUPPORTED:
      throw std::invalid_argument{msg};
    case XNTT_ERR_ALLOC:
      throw std::bad_alloc{};
    case XNTT_ERR_STATE:
      throw std::logic_error{msg};
    default:
      throw std::runtime_error{msg + " (" + xntt_last_cuda_error() + ")"};
  }
}
inline void check(int status, const char* what) {
  if (status != XNTT_OK) throw_status(status, what);
}
}  // namespace detail

// Devices a subsequently constructed NTT<> spreads ONE transform over (empty or one entry: single GPU).  This is the
// multi-GPU counterpart of the reference's OpenMP thread team (kernel/recursive.hpp:65): user code stays
// `sventt::NTT<kernel> ntt; ntt.compute_forward(a.data());`, the library shards the six-step composition
// (xntt_mgpu_*, one process, peer-to-peer stores, no collective library).  2, 4 or 8 devices; production modulus.
inline std::vector<int>& default_devices() {
  static std::vector<int> devices;
  return devices;
}
inline void set_default_devices(std::vector<int> devices) { default_devices() = std::move(devices); }

template <class kernel_type_>
class NTT {
 public:
  using kernel_type = kernel_type_;
  using modulus_type = typename kernel_type::modulus_type;

  explicit NTT(bool enable_forward = true, bool enable_inverse = true, bool /*allocate_huge_pages*/ = true,
               std::uint32_t batch = 1, int device = -1)
      : NTT(batch == 1 && device < 0 ? default_devices() : std::vector<int>{}, enable_forward, enable_inverse, batch,
            device) {}

  // one transform over several GPUs of this process (see set_default_devices)
  explicit NTT(const std::vector<int>& devices, bool enable_forward = true, bool enable_inverse = true,
               std::uint32_t batch = 1, int device = -1) {
    xntt_desc d{};
    d.modulus = modulus_type::get_modulus();
    d.generator = modulus_type::get_generator();
    d.log2_m = detail::log2_exact(get_m());
    d.batch = batch;
    d.inverse_factor = kernel_type::get_inverse_factor();
    d.flags = (enable_forward ? static_cast<std::uint32_t>(XNTT_ENABLE_FORWARD) : 0u) |
              (enable_inverse ? static_cast<std::uint32_t>(XNTT_ENABLE_INVERSE) : 0u);
    if (d.flags == 0) d.flags = XNTT_ENABLE_FORWARD | XNTT_ENABLE_INVERSE;
    if constexpr (kernel_type::all_fixed_point()) d.flags |= XNTT_MODMUL_FIXED_POINT;  // FixedPoint64 layers throughout
    d.device = device;
    std::vector<std::uint32_t> splits;
    kernel_type::append_splits(splits);
    const bool multi = devices.size() > 1;
    std::vector<std::int32_t> devs(devices.begin(), devices.end());
    auto create = [&]() {
      return multi ? xntt_mgpu_create(&mgpu_, &d, devs.data(), static_cast<std::uint32_t>(devs.size()))
                   : xntt_plan_create(&plan_, &d);
    };
    if (!multi && devices.size() == 1) d.device = devices[0];
    int status = XNTT_ERR_UNSUPPORTED;
    if (splits.size() >= 2 && splits.size() <= XNTT_MAX_SPLITS) {
      d.n_splits = static_cast<std::uint32_t>(splits.size());
      for (std::size_t i = 0; i < splits.size(); ++i) d.split_log2[i] = splits[i];
      status = create();
    }
    if (status == XNTT_ERR_UNSUPPORTED || (multi && status == XNTT_ERR_INVALID && d.n_splits != 0)) {
      // the composition asks for a tile shape the shared-memory kernels do not have (or is a single
      // unit, or a first split the device count does not divide): same transform, planner's own decomposition
      d.n_splits = 0;
      status = create();
    }
    detail::check(status, "sventt::NTT");
  }
  NTT(const NTT&) = delete;
  NTT& operator=(const NTT&) = delete;
  ~NTT() {
    xntt_plan_destroy(plan_);
    xntt_mgpu_destroy(mgpu_);
  }
  // number of GPUs this transform runs on
  std::uint32_t get_device_count() const { return mgpu_ ? xntt_mgpu_devices(mgpu_) : 1u; }

  static constexpr std::uint64_t get_m() { return kernel_type::get_m(); }

  void compute_forward(std::uint64_t* dst, const std::uint64_t* src) const { run(dst, src, false); }
  void compute_forward(std::uint64_t* dst) const { run(dst, dst, false); }
  void compute_inverse(std::uint64_t* dst, const std::uint64_t* src) const { run(dst, src, true); }
  void compute_inverse(std::uint64_t* dst) const { run(dst, dst, true); }

  // stream-ordered variants for device-resident data (no reference counterpart)
  void compute_forward_async(std::uint64_t* dst, const std::uint64_t* src, void* stream) const {
    detail::check(xntt_forward(plan_, dst, src, stream), "compute_forward");
  }
  void compute_inverse_async(std::uint64_t* dst, const std::uint64_t* src, void* stream) const {
    detail::check(xntt_inverse(plan_, dst, src, stream), "compute_inverse");
  }
  // forward transform fused with the point-wise multiply_normalize against a to_montgomery'd spectrum
  // (the loop at examples/magic-series/gaussian-polynomial.hpp:201-212); device pointers
  void compute_forward_multiply_async(std::uint64_t* dst, const std::uint64_t* src, const std::uint64_t* b_mont,
                                      void* stream) const {
    detail::check(xntt_forward_multiply(plan_, dst, src, b_mont, stream), "compute_forward_multiply");
  }
  const xntt_plan* plan() const { return plan_; }
  xntt_mgpu* mgpu() const { return mgpu_; }
  // 1 when the device kernels run FixedPoint64 (Shoup) arithmetic, 0 for PAdic64 (Montgomery)
  std::uint32_t get_modmul_kind() const { return plan_ ? xntt_plan_modmul(plan_) : 0u; }
  // multi-GPU plans, device-resident shards (layouts: include/xntt.h, xntt_mgpu_forward); asynchronous
  void compute_forward_shards(std::uint64_t* const* dst, const std::uint64_t* const* src) const {
    detail::check(mgpu_ ? xntt_mgpu_forward(mgpu_, dst, src) : XNTT_ERR_STATE, "compute_forward_shards");
  }
  void compute_inverse_shards(std::uint64_t* const* dst, const std::uint64_t* const* src) const {
    detail::check(mgpu_ ? xntt_mgpu_inverse(mgpu_, dst, src) : XNTT_ERR_STATE, "compute_inverse_shards");
  }
  void synchronize() const {
    detail::check(mgpu_ ? xntt_mgpu_synchronize(mgpu_) : xntt_stream_synchronize(nullptr), "synchronize");
  }

 private:
  void run(std::uint64_t* dst, const std::uint64_t* src, bool inverse) const {
    if (mgpu_) {
      // whole transform on host buffers, scattered over the GPUs (device-resident data: compute_*_shards)
      const int kd = xntt_pointer_is_device(dst), ks = xntt_pointer_is_device(src);
      int status = (kd == 0 && ks == 0)
                       ? (inverse ? xntt_mgpu_inverse_host(mgpu_, dst, src) : xntt_mgpu_forward_host(mgpu_, dst, src))
                       : XNTT_ERR_INVALID;
      detail::check(status, inverse ? "compute_inverse" : "compute_forward");
      return;
    }
    const int kd = xntt_pointer_is_device(dst), ks = xntt_pointer_is_device(src);
    if (kd < 0 || ks < 0) detail::throw_status(kd < 0 ? kd : ks, "compute");
    int status;
    if (kd && ks) {
      status = inverse ? xntt_inverse(plan_, dst, src, nullptr) : xntt_forward(plan_, dst, src, nullptr);
      if (status == XNTT_OK) status = xntt_stream_synchronize(nullptr);
    } else if (!kd && !ks) {
      status = inverse ? xntt_inverse_host(plan_, dst, src) : xntt_forward_host(plan_, dst, src);
    } else {
      status = XNTT_ERR_INVALID;  // mixing host and device buffers in one call
    }
    detail::check(status, inverse ? "compute_inverse" : "compute_forward");
  }

  xntt_plan* plan_ = nullptr;
  xntt_mgpu* mgpu_ = nullptr;
};

}  // namespace sventt

#endif
