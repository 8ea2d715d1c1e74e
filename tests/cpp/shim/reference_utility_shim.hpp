// SPDX-License-Identifier: Apache-2.0
// Test infrastructure, force-included (-include) in front of the UNMODIFIED reference driver tests/bench-ntt.cpp:
// the reference's tests/utility.hpp contains Arm SVE intrinsics (reduce_serial) and cannot compile on this host, so its
// include guard is pre-defined here and the two helpers the driver calls are provided with the same meaning
// (tests/utility.hpp:123-154: fill with one byte value; fill with value, value + 1, ...).
#pragma once
#define SVENTT_TESTS_UTILITY_HPP_INCLUDED
#include <cstdint>
#include <cstring>
#include <iterator>
#include <numeric>

[[maybe_unused]] static void memset_parallel(void* const dst, const std::uint8_t value, const std::uint64_t size) {
  std::memset(dst, value, size);
}

template <class iterator_type, class value_type>
[[maybe_unused]] static void iota_parallel(iterator_type begin, iterator_type end, const value_type value) {
  std::iota(begin, end, static_cast<std::iter_value_t<iterator_type>>(value));
}
