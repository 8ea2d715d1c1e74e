// SPDX-License-Identifier: Apache-2.0
// TEST INFRASTRUCTURE.  Stand-in for <boost/align/aligned_allocator.hpp>, which the reference's vector.hpp includes
// (include/sventt/vector.hpp:20) and which is not installed in this image.  Only the scalar CPU baseline
// (oracle/refscalar.cpp) is compiled against it; that path uses sventt::AuxiliaryVector / PageMemory, not the allocator,
// so an allocator with the Boost class's name and shape is all that is needed.
#pragma once
#include <cstddef>
#include <cstdlib>
#include <new>

namespace boost {
namespace alignment {
template <class T, std::size_t Alignment = alignof(T)>
class aligned_allocator {
 public:
  using value_type = T;
  template <class U>
  struct rebind {
    using other = aligned_allocator<U, Alignment>;
  };
  aligned_allocator() = default;
  template <class U>
  aligned_allocator(const aligned_allocator<U, Alignment>&) {}
  T* allocate(std::size_t n) {
    const std::size_t a = Alignment < sizeof(void*) ? sizeof(void*) : Alignment;
    void* p = nullptr;
    if (posix_memalign(&p, a, n * sizeof(T) ? n * sizeof(T) : a) != 0) throw std::bad_alloc{};
    return static_cast<T*>(p);
  }
  void deallocate(T* p, std::size_t) { std::free(p); }
  template <class U>
  bool operator==(const aligned_allocator<U, Alignment>&) const { return true; }
  template <class U>
  bool operator!=(const aligned_allocator<U, Alignment>&) const { return false; }
};
}  // namespace alignment
}  // namespace boost
