// SPDX-License-Identifier: Apache-2.0
//
// Kinnaes' formula for the number of magic series, the modmul-only workload of the reference's
// examples/magic-series-kinnaes (kinnaes.hpp:51-157, MagicSeriesKinnaes::compute_sum):
//
//   S(j_begin, j_end) = sum_{J = j_begin+1}^{j_end}  prod_{l<m} (w^(J (m^2-m+1+l)) - 1)
//                                                    ---------------------------------------
//                                                    w^(J r) prod_{l<m} (w^(J (l+1)) - 1)
//
// w = primitive n-th root of unity (n odd, not a power of two), r = m (m-1)/2 * m.  The reference walks J in
// SVE vectors of cntd lanes and keeps one running fraction per lane; here one thread owns one J at a time (its
// 4 m PAdic64 products are the whole cost), keeps a running fraction over the J it visits, and the CTA folds its
// 256 fractions a/b + c/d = (a d + c b)/(b d) in shared memory.  One (numerator, denominator) pair per CTA goes
// back to the host, which folds those and performs the single division (kinnaes.hpp:143-156).
// Everything is kept in Montgomery form, so PAdic64::multiply maps to FieldOps::mont.
#pragma once
#include "field.cuh"
#include "params.h"

namespace xntt {

// base^e, base and result in Montgomery form
template <class F>
__device__ __forceinline__ u64 kn_pow(const F& f, u64 base, u64 e) {
  u64 acc = f.one();
#pragma unroll 1
  for (; e; e >>= 1) {
    if (e & 1) acc = f.mont(acc, base, f.companion(base));
    base = f.mont(base, base, f.companion(base));
  }
  return acc;
}
// (a - 1) mod p in Montgomery form, a canonical
template <class F>
__device__ __forceinline__ u64 kn_dec(const F& f, u64 a) {
  const u64 one = f.one();
  return a >= one ? a - one : a - one + f.p();
}
template <class F>
__device__ __forceinline__ u64 kn_add(const F& f, u64 a, u64 b) {
  const u64 p = f.p();
  return a < p - b ? a + b : a + b - p;  // a, b canonical (kinnaes.hpp:129: modmul_type::add)
}

// numerator and denominator product of one J (kinnaes.hpp:103-124)
template <class F>
__device__ __forceinline__ void kinnaes_term(const F& f, const KinnaesParams& prm, u64 J, u64& num_prod, u64& den_prod) {
  // w^J from the ladder
  u64 wj = f.one();
  {
    u64 e = J;
#pragma unroll 1
    for (int i = 0; e; ++i, e >>= 1)
      if (e & 1) wj = f.mont(wj, prm.ladder[i], f.companion(prm.ladder[i]));
  }
  const u64 wjp = f.companion(wj);
  u64 num_term = kn_pow(f, wj, prm.exp_num), den_term = wj;
  num_prod = f.one();
  den_prod = kn_pow(f, wj, prm.exp_r);
#pragma unroll 1
  for (u64 l = 0; l < prm.m; ++l) {
    num_prod = f.mont(kn_dec(f, num_term), num_prod, f.companion(num_prod));
    den_prod = f.mont(kn_dec(f, den_term), den_prod, f.companion(den_prod));
    num_term = f.mont(num_term, wj, wjp);
    den_term = f.mont(den_term, wj, wjp);
  }
}

// (ns/ds) += (n/d)   (kinnaes.hpp:126-133)
template <class F>
__device__ __forceinline__ void kinnaes_fold(const F& f, u64& ns, u64& ds, u64 n, u64 d) {
  const u64 dp = f.companion(d);
  const u64 t0 = f.mont(ds, n, f.companion(n)), t1 = f.mont(ns, d, dp);
  ns = kn_add(f, t0, t1);
  ds = f.mont(ds, d, dp);
}

#if !defined(XNTT_HOST_EMU)
template <class F>
__global__ void __launch_bounds__(kKinnaesThreads) kinnaes_kernel(const F f, const __grid_constant__ KinnaesParams prm) {
  __shared__ u64 sn[kKinnaesThreads], sd[kKinnaesThreads];
  u64 ns = 0, ds = f.one();
  for (u64 i = (u64)blockIdx.x * kKinnaesThreads + threadIdx.x; i < prm.count; i += (u64)gridDim.x * kKinnaesThreads) {
    u64 n, d;
    kinnaes_term<F>(f, prm, prm.j_first + i, n, d);
    kinnaes_fold<F>(f, ns, ds, n, d);
  }
  sn[threadIdx.x] = ns;
  sd[threadIdx.x] = ds;
  __syncthreads();
  for (int s = kKinnaesThreads / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      kinnaes_fold<F>(f, ns, ds, sn[threadIdx.x + s], sd[threadIdx.x + s]);
      sn[threadIdx.x] = ns;
      sd[threadIdx.x] = ds;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    prm.partial[2 * blockIdx.x] = ns;
    prm.partial[2 * blockIdx.x + 1] = ds;
  }
}
#endif

}  // namespace xntt
