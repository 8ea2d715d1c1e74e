// SPDX-License-Identifier: Apache-2.0
// Lab: butterflies that hand the multiplications by the 11-bit constant K = 1827 of the production prime
// P = 2^64 - K*2^31 + 1 to the FP64 pipe.  A 32-bit integer x paired with an exponent word IS the double
// 2^e + x*2^(e-52), so a DFMA multiplies x by K exactly with no conversion instruction on the way in, and with
// c = K * 2^-1074 (a denormal) the result is the integer itself in the low mantissa bits on the way out.
//
//   q*P = (q - T_hi) * 2^64 + (q - T_lo * 2^31),  T = q*K = T_hi * 2^33 + T_lo
//   hi64(q*P) = q - T_hi - [q < T_lo*2^31] = q - T_hi - [L > q]      (L = lo64(q*P) = lo64(a*w), known)
//   T_hi = qa*K + floor(qb*K / 2^33),  q = qa * 2^33 + qb
#pragma once
#include "field.cuh"

namespace lab {
using xntt::u32;
using xntt::u64;
using xntt::pack64;
using xntt::unpack64;

constexpr unsigned long long kK = 1827;

// floor(q * K / 2^33) through two DFMA + one DADD (denormal outputs: bits = integer)
__device__ __forceinline__ u64 thi_fp64(u32 q0, u32 q1) {
  const double cK = __longlong_as_double((long long)kK);             // K * 2^-1074
  const double z0 = -__longlong_as_double((long long)(kK << 19));    // -(2^19 * K * 2^-1074)
  const double dqb = __hiloint2double((int)((q1 & 1u) | 0x41200000u), (int)q0);  // 2^19 + qb * 2^-33
  const double f0 = __fma_rd(dqb, cK, z0);                           // floor(qb*K/2^33) * 2^-1074
  const double dqa = __hiloint2double(0x43300000, (int)(q1 >> 1)) - 4503599627370496.0;  // qa, exact
  const double f1 = fma(dqa, cK, f0);                                // (qa*K + g) * 2^-1074, exact
  return (u64)__double_as_longlong(f1);
}
// the same with the classic 2^52 magic (normal numbers only): result carries 0x43300000 in its high word
__device__ __forceinline__ u64 thi_fp64_magic(u32 q0, u32 q1) {
  const double dqb = __hiloint2double((int)((q1 & 1u) | 0x43300000u), (int)q0);  // 2^52 + qb
  const double f0 = __fma_rd(dqb, (double)kK * 0x1p-33, 0x1p52 - (double)kK * 0x1p19);  // 2^52 + g
  const double dqa = __hiloint2double(0x43300000, (int)(q1 >> 1)) - 0x1p52;
  const double f1 = fma(dqa, (double)kK, f0);                        // 2^52 + T_hi
  return (u64)__double_as_longlong(f1) - 0x4330000000000000ull;
}
// integer-pipe version of the same quantity (2 wide): for the emulator / as a cross-check
__device__ __forceinline__ u64 thi_int(u32 q0, u32 q1) {
  const u64 lo = (u64)q0 * kK;
  const u64 hi = (u64)q1 * kK + (lo >> 32);
  return hi >> 1;
}

// h1 = hi64(a*w), L = lo64(a*w)  (4 wide)
__device__ __forceinline__ void mul_full(u64 a, u64 w, u64& h1, u32& l0, u32& l1) {
  u32 a0, a1, w0, w1, h1l, h1h, vl, vh, lh;
  unpack64(a, a0, a1);
  unpack64(w, w0, w1);
  unpack64((u64)a0 * w0, vl, vh);
  asm("{\n\t.reg .u32 xl, xh, xc;\n\t"
      "mul.lo.u32 xl, %3, %6;\n\tmul.hi.u32 xh, %3, %6;\n\t"
      "mad.lo.cc.u32 xl, %4, %5, xl;\n\tmadc.hi.cc.u32 xh, %4, %5, xh;\n\taddc.u32 xc, 0, 0;\n\t"
      "add.cc.u32 %2, xl, %7;\n\t"
      "madc.lo.cc.u32 %0, %4, %6, xh;\n\tmadc.hi.u32 %1, %4, %6, xc;\n\t"
      "}"
      : "=r"(h1l), "=r"(h1h), "=r"(lh)
      : "r"(a0), "r"(a1), "r"(w0), "r"(w1), "r"(vh));
  h1 = pack64(h1l, h1h);
  l0 = vl;
  l1 = lh;
}

// h2 = q - T - [L > q]
__device__ __forceinline__ u64 h2_from(u32 q0, u32 q1, u32 l0, u32 l1, u64 T) {
  u32 t0, t1, r0, r1;
  unpack64(T, t0, t1);
  asm("{\n\t.reg .u32 t;\n\t"
      "sub.cc.u32 t, %2, %4;\n\tsubc.cc.u32 t, %3, %5;\n\t"   // CF = [q < L]
      "subc.cc.u32 %0, %2, %6;\n\tsubc.u32 %1, %3, %7;\n\t"
      "}"
      : "=r"(r0), "=r"(r1)
      : "r"(q0), "r"(q1), "r"(l0), "r"(l1), "r"(t0), "r"(t1));
  return pack64(r0, r1);
}

// q = L * P^-1 mod 2^64 with P^-1 = 1 + K*2^31 + 2^62: one DFMA (L0*K) + shifts; no w' needed
__device__ __forceinline__ void q_from_L(u32 l0, u32 l1, u32& q0, u32& q1) {
  const double cK = __longlong_as_double((long long)kK);
  const double z = -__longlong_as_double((long long)(kK << 52));  // -(2^52 * K * 2^-1074) = -K * 2^-1022
  const double f = fma(__hiloint2double(0x43300000, (int)l0), cK, z);  // L0*K * 2^-1074
  u32 f0, f1;
  unpack64((u64)__double_as_longlong(f), f0, f1);
  const u32 m = __funnelshift_r(f0, f1, 1);  // bits 1..32 of L0*K
  const u32 e = (l1 << 31) + (l0 << 30);
  asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, %5;" : "=r"(q0), "=r"(q1) : "r"(l0), "r"(l0 << 31), "r"(l1 + m), "r"(e));
}

template <int THI, int MODE, bool QFP>
__device__ __forceinline__ void bf_fp64(u64& x0, u64& x1, u64 w, u64 wp) {
  const xntt::F0 f{};
  u64 h1, u, s, d;
  u32 l0, l1, q0, q1, m, d0, d1;
  mul_full(x1, w, h1, l0, l1);
  if constexpr (QFP)
    q_from_L(l0, l1, q0, q1);
  else
    unpack64(x1 * wp, q0, q1);
  const u64 T = THI == 0 ? thi_fp64(q0, q1) : THI == 1 ? thi_fp64_magic(q0, q1) : thi_int(q0, q1);
  const u64 h2 = h2_from(q0, q1, l0, l1, T);
  xntt::sub_borrow_mask(h1, h2, u, m);
  xntt::add_carry_plus(x0, u, m, s, d0);
  xntt::sub_borrow_minus(x0, u, m, d, d1);
  x0 = (MODE & 1) ? fix_alu(s, d0) : f.fix(s, d0);
  x1 = (MODE & 2) ? fix_alu(d, d1) : f.fix(d, d1);
}
}  // namespace lab
