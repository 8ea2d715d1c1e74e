// SPDX-License-Identifier: Apache-2.0
//
// Per-element bodies of the small kernels: on-device twiddle generation (north-star item 4; replaces
// the serial host loops of every prepare_forward / prepare_inverse, e.g.
// include/sventt/layer/sve/radix-eight.hpp:34-93 and include/sventt/layer/sve/generic.hpp:72-110)
// and the PAdic64 element-wise helpers.
#pragma once
#include "field.cuh"
#include "params.h"

namespace xntt {

#if defined(XNTT_HOST_EMU)
inline u32 brev32(u32 v) {
  u32 r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
inline int clz32(u32 v) { return v ? __builtin_clz(v) : 32; }
#else
__device__ __forceinline__ u32 brev32(u32 v) { return __brev(v); }
__device__ __forceinline__ int clz32(u32 v) { return __clz(v); }
#endif

template <class F>
__device__ __forceinline__ u64 pow_mont(const F& f, const PowTable& t, u32 e) {
  u64 acc = t.scale;
#pragma unroll 1
  for (int i = 0; e; ++i, e >>= 1)
    if (e & 1) acc = f.mont(acc, t.sq[i], f.companion(t.sq[i]));
  return acc;
}

// exponent of table entry idx
__device__ __forceinline__ u32 table_exponent(u32 idx, int kind, int logn, int shift, u32 col0 = 0, u32 row0 = 0) {
  if (kind == kFwdG) {
    // G[b] = omega_N^bitrev_{logn-1}(b)
    return logn > 1 ? (brev32(idx) >> (32 - (logn - 1))) : 0u;
  } else if (kind == kInvI) {
    // I[l + j] = (omega_N^-1)^(j * N / 2l)
    if (idx == 0) return 0u;
    const int ll = 31 - clz32(idx);  // log2 l
    return (idx - (1u << ll)) << (logn - 1 - ll);
  }
  if (kind == kTwist) {
    // twist matrix entry (k, c), idx = (k << shift) + c: omega_M^(bitrev_logn(k) * (col0 + c)); the table holds 2^shift
    // columns starting at global column col0 and rows starting at row0 (the whole matrix: both 0)
    const u32 k = row0 + (idx >> shift), c = idx & ((1u << shift) - 1u);
    return (logn ? (brev32(k) >> (32 - logn)) : 0u) * (col0 + c);
  }
  return idx << shift;
}

template <class F>
__device__ __forceinline__ Tw table_entry(const F& f, u32 idx, int kind, int logn, int shift, const PowTable& t) {
  return f.make_tw(pow_mont<F>(f, t, table_exponent(idx, kind, logn, shift, t.col0, t.row0)));
}

// PAdic64::to_montgomery (p-adic-64.hpp:19-22): a * 2^64 mod P, r2 = 2^128 mod P
template <class F>
__device__ __forceinline__ u64 ew_to_mont(const F& f, u64 a, u64 r2, u64 r2p) {
  return f.mont(a, r2, r2p);
}
// PAdic64::from_montgomery (p-adic-64.hpp:24-38), canonical
template <class F>
__device__ __forceinline__ u64 ew_from_mont(const F& f, u64 a) {
  return f.mont(a, 1ull, f.pinv());
}
// PAdic64::multiply_normalize (p-adic-64.hpp:101-115) with on-the-fly precompute
template <class F>
__device__ __forceinline__ u64 ew_mulnorm(const F& f, u64 a, u64 b) {
  return f.mont(a, b, f.companion(b));
}

}  // namespace xntt
