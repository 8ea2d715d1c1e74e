// SPDX-License-Identifier: Apache-2.0
//
// sventt::PageMemory<T> - the page-aligned owner of the reference (include/sventt/vector.hpp:61-168:
// size / data / operator[] / at / begin / end / reset) backed by PINNED host memory, so that the
// host entry points of libxntt can DMA straight out of it.  DeviceMemory<T> is the device-resident
// twin for callers that keep their residues in HBM.
#ifndef XNTT_SVENTT_VECTOR_HPP
#define XNTT_SVENTT_VECTOR_HPP

#include <cstddef>
#include <cstdint>
#include <new>
#include <stdexcept>
#include <utility>

#include "xntt.h"

namespace sventt {

template <class value_type_>
class PageMemory {
 public:
  using value_type = value_type_;
  using size_type = std::uint64_t;

  PageMemory() = default;
  // the second argument (huge pages) is accepted for source compatibility and ignored
  PageMemory(size_type length, bool /*allocate_huge_pages*/ = false) { reset(length); }
  PageMemory(const PageMemory&) = delete;
  PageMemory& operator=(const PageMemory&) = delete;
  PageMemory(PageMemory&& o) noexcept : length_{o.length_}, ptr_{o.ptr_} { o.length_ = 0, o.ptr_ = nullptr; }
  PageMemory& operator=(PageMemory&& o) noexcept {
    if (this != &o) {
      reset();
      std::swap(length_, o.length_);
      std::swap(ptr_, o.ptr_);
    }
    return *this;
  }
  ~PageMemory() { reset(); }

  size_type size() const { return length_; }
  value_type* data() { return ptr_; }
  const value_type* data() const { return ptr_; }

  void reset() {
    if (ptr_) xntt_free_pinned(ptr_);
    ptr_ = nullptr;
    length_ = 0;
  }
  void reset(size_type len, bool /*allocate_huge_pages*/ = false) {
    reset();
    if (len == 0) return;
    void* p = nullptr;
    if (xntt_alloc_pinned(&p, sizeof(value_type) * len) != XNTT_OK) throw std::bad_alloc{};
    ptr_ = static_cast<value_type*>(p);
    length_ = len;
  }

  value_type& operator[](size_type i) { return ptr_[i]; }
  const value_type& operator[](size_type i) const { return ptr_[i]; }
  value_type& at(size_type i) {
    if (i >= length_) throw std::out_of_range{"Index out of range"};
    return ptr_[i];
  }
  const value_type& at(size_type i) const {
    if (i >= length_) throw std::out_of_range{"Index out of range"};
    return ptr_[i];
  }
  value_type* begin() { return ptr_; }
  value_type* end() { return ptr_ + length_; }
  const value_type* begin() const { return ptr_; }
  const value_type* end() const { return ptr_ + length_; }
  const value_type* cbegin() const { return ptr_; }
  const value_type* cend() const { return ptr_ + length_; }

 private:
  size_type length_ = 0;
  value_type* ptr_ = nullptr;
};

template <class value_type_>
class DeviceMemory {
 public:
  using value_type = value_type_;
  using size_type = std::uint64_t;

  DeviceMemory() = default;
  explicit DeviceMemory(size_type length, int device = -1) { reset(length, device); }
  DeviceMemory(const DeviceMemory&) = delete;
  DeviceMemory& operator=(const DeviceMemory&) = delete;
  ~DeviceMemory() { reset(); }

  size_type size() const { return length_; }
  value_type* data() { return ptr_; }
  const value_type* data() const { return ptr_; }
  void reset() {
    if (ptr_) xntt_free_device(ptr_);
    ptr_ = nullptr;
    length_ = 0;
  }
  void reset(size_type len, int device = -1) {
    reset();
    if (len == 0) return;
    void* p = nullptr;
    if (xntt_alloc_device(&p, sizeof(value_type) * len, device) != XNTT_OK) throw std::bad_alloc{};
    ptr_ = static_cast<value_type*>(p);
    length_ = len;
  }
  void copy_from_host(const value_type* src, void* stream = nullptr) {
    if (xntt_memcpy_h2d(ptr_, src, sizeof(value_type) * length_, stream) != XNTT_OK ||
        xntt_stream_synchronize(stream) != XNTT_OK)
      throw std::runtime_error{xntt_last_cuda_error()};
  }
  void copy_to_host(value_type* dst, void* stream = nullptr) const {
    if (xntt_memcpy_d2h(dst, ptr_, sizeof(value_type) * length_, stream) != XNTT_OK ||
        xntt_stream_synchronize(stream) != XNTT_OK)
      throw std::runtime_error{xntt_last_cuda_error()};
  }

 private:
  size_type length_ = 0;
  value_type* ptr_ = nullptr;
};

}  // namespace sventt

#endif
