// SPDX-License-Identifier: Apache-2.0
// Lab: candidate butterfly formulations (see bfly_lab.cu).  Results on B200 (cycles per warp-butterfly at 1965 MHz, static
// fma-pipe cycles in brackets): v0 = field.cuh at the start of the round 61.8 [55.2]; v8 = full low product instead of
// IMAD.HI 60.6 [53.2] (adopted); v1 C-level conditionals 71.5; v3 PTX predicates 66.1; v4 canonical u + three fma repairs
// 64.5; v5 plain C 71.3; v10 / v11 repairs on the ALU side: same static fma count as v9 (ptxas rebalances);
// probes: no repairs 45.3-47.6, Montgomery product alone 49.2-50.4.
#pragma once
#include "field.cuh"

namespace lab {
using xntt::u32;
using xntt::u64;
using xntt::pack64;
using xntt::unpack64;

constexpr u64 P = xntt::kP0;
constexpr u64 C = 0 - P;
constexpr u32 P_LO = (u32)P, P_HI = (u32)(P >> 32), C_LO = (u32)C, C_HI = (u32)(C >> 32);

// u = (h1 - h2) mod 2^64, m = -borrow
__device__ __forceinline__ void mont_diff(u64 a, u64 w, u64 wp, u64& u, u32& m) {
  u32 a0, a1, w0, w1, q0, q1, ul, uh;
  unpack64(a, a0, a1);
  unpack64(w, w0, w1);
  unpack64(a * wp, q0, q1);
  asm("{\n\t.reg .u32 xl, xh, xc, vh, lh, dl, dh, yl, yh, yc, el, eh, t, h1l, h1h, h2l, h2h;\n\t"
      "mul.hi.u32 vh, %3, %5;\n\t"
      "mul.lo.u32 xl, %3, %6;\n\tmul.hi.u32 xh, %3, %6;\n\t"
      "mad.lo.cc.u32 xl, %4, %5, xl;\n\tmadc.hi.cc.u32 xh, %4, %5, xh;\n\taddc.u32 xc, 0, 0;\n\t"
      "mul.lo.u32 dl, %4, %6;\n\tmul.hi.u32 dh, %4, %6;\n\t"
      "add.cc.u32 lh, xl, vh;\n\taddc.cc.u32 h1l, dl, xh;\n\taddc.u32 h1h, dh, xc;\n\t"
      "mul.lo.u32 yl, %7, %10;\n\tmul.hi.u32 yh, %7, %10;\n\t"
      "mad.lo.cc.u32 yl, %8, %9, yl;\n\tmadc.hi.cc.u32 yh, %8, %9, yh;\n\taddc.u32 yc, 0, 0;\n\t"
      "mul.lo.u32 el, %8, %10;\n\tmul.hi.u32 eh, %8, %10;\n\t"
      "not.b32 t, lh;\n\tadd.cc.u32 t, yl, t;\n\taddc.cc.u32 h2l, el, yh;\n\taddc.u32 h2h, eh, yc;\n\t"
      "sub.cc.u32 %0, h1l, h2l;\n\tsubc.cc.u32 %1, h1h, h2h;\n\tsubc.u32 %2, 0, 0;\n\t"
      "}"
      : "=r"(ul), "=r"(uh), "=r"(m)
      : "r"(a0), "r"(a1), "r"(w0), "r"(w1), "r"(q0), "r"(q1), "r"(P_LO), "r"(P_HI));
  u = pack64(ul, uh);
}

// v1: canonical u, C-level conditionals
__device__ __forceinline__ void bf_v1(u64& x0, u64& x1, u64 w, u64 wp) {
  u64 u;
  u32 m;
  mont_diff(x1, w, wp, u, m);
  if (m) u -= C;  // + P
  const u64 a = x0;
  u64 s = a + u;
  if (s < a) s += C;
  u64 d = a - u;
  if (a < u) d -= C;
  x0 = s;
  x1 = d;
}

// v2: canonical u, everything on carry flags + masks (LOP3).  NOTE: this sketch returns wrong residues (the run on the
// box flagged it, 65.4 cycles anyway) and is not part of main(); kept only as the record of what was timed.
__device__ __forceinline__ void bf_v2(u64& x0, u64& x1, u64 w, u64 wp) {
  u64 u;
  u32 m;
  mont_diff(x1, w, wp, u, m);
  u32 ul, uh, al, ah, sl, sh, dl, dh;
  unpack64(u, ul, uh);
  unpack64(x0, al, ah);
  asm("{\n\t.reg .u32 t0, t1, k;\n\t"
      // u -= m & C
      "and.b32 t0, %6, %9;\n\tand.b32 t1, %6, %10;\n\t"
      "sub.cc.u32 %4, %4, t0;\n\tsubc.u32 %5, %5, t1;\n\t"
      // s = a + u; k = -carry
      "add.cc.u32 %0, %7, %4;\n\taddc.cc.u32 %1, %8, %5;\n\taddc.u32 k, 0xffffffff, 0;\n\t"  // k = carry - 1
      "lop3.b32 t0, k, %9, 0, 0x44;\n\tlop3.b32 t1, k, %10, 0, 0x44;\n\t"                       // ~k & C
      "add.cc.u32 %0, %0, t0;\n\taddc.u32 %1, %1, t1;\n\t"
      // d = a - u; k = -borrow
      "sub.cc.u32 %2, %7, %4;\n\tsubc.cc.u32 %3, %8, %5;\n\tsubc.u32 k, 0, 0;\n\t"
      "and.b32 t0, k, %9;\n\tand.b32 t1, k, %10;\n\t"
      "sub.cc.u32 %2, %2, t0;\n\tsubc.u32 %3, %3, t1;\n\t"
      "}"
      : "=r"(sl), "=r"(sh), "=r"(dl), "=r"(dh), "+r"(ul), "+r"(uh)
      : "r"(m), "r"(al), "r"(ah), "r"(C_LO), "r"(C_HI));
  x0 = pack64(sl, sh);
  x1 = pack64(dl, dh);
}

// v3: canonical u; PTX predicates taken from materialised carries
__device__ __forceinline__ void bf_v3(u64& x0, u64& x1, u64 w, u64 wp) {
  u64 u;
  u32 m;
  mont_diff(x1, w, wp, u, m);
  u32 ul, uh, al, ah, sl, sh, dl, dh;
  unpack64(u, ul, uh);
  unpack64(x0, al, ah);
  asm("{\n\t.reg .u32 k;\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %6, 0;\n\t"
      "@p sub.cc.u32 %4, %4, %9;\n\t@p subc.u32 %5, %5, %10;\n\t"
      "add.cc.u32 %0, %7, %4;\n\taddc.cc.u32 %1, %8, %5;\n\taddc.u32 k, 0, 0;\n\t"
      "setp.ne.u32 p, k, 0;\n\t"
      "@p add.cc.u32 %0, %0, %9;\n\t@p addc.u32 %1, %1, %10;\n\t"
      "sub.cc.u32 %2, %7, %4;\n\tsubc.cc.u32 %3, %8, %5;\n\tsubc.u32 k, 0, 0;\n\t"
      "setp.ne.u32 p, k, 0;\n\t"
      "@p sub.cc.u32 %2, %2, %9;\n\t@p subc.u32 %3, %3, %10;\n\t"
      "}"
      : "=r"(sl), "=r"(sh), "=r"(dl), "=r"(dh), "+r"(ul), "+r"(uh)
      : "r"(m), "r"(al), "r"(ah), "r"(C_LO), "r"(C_HI));
  x0 = pack64(sl, sh);
  x1 = pack64(dl, dh);
}

// v4: canonical u by mask; outputs repaired with one signed IMAD.WIDE + IMAD each (fma pipe)
__device__ __forceinline__ u64 fixw(u64 v, u32 delta) {
  u64 t = v + (u64)((long long)(int)delta * (long long)(int)C_LO);
  u32 tl, th;
  unpack64(t, tl, th);
  th += delta * C_HI;
  return pack64(tl, th);
}
__device__ __forceinline__ void bf_v4(u64& x0, u64& x1, u64 w, u64 wp) {
  u64 u;
  u32 m;
  mont_diff(x1, w, wp, u, m);
  u = fixw(u, m);
  u64 s, d;
  u32 k0, k1;
  xntt::add_carry_plus(x0, u, 0u, s, k0);
  xntt::sub_borrow_minus(x0, u, 0u, d, k1);
  x0 = fixw(s, k0);
  x1 = fixw(d, k1);
}


// v5: plain C
__device__ __forceinline__ void bf_v5(u64& x0, u64& x1, u64 w, u64 wp) {
  const u64 a = x1;
  const u64 h1 = __umul64hi(a, w), q = a * wp, h2 = __umul64hi(q, P);
  u64 u = h1 - h2;
  if (h1 < h2) u += P;
  const u64 b = x0;
  u64 s = b + u;
  if (s < b) s += C;
  u64 d = b - u;
  if (b < u) d -= C;
  x0 = s;
  x1 = d;
}
}  // namespace lab
namespace lab {
// v6: cost probe (NOT a correct butterfly): products and add/sub, no repairs at all
__device__ __forceinline__ void bf_v6(u64& x0, u64& x1, u64 w, u64 wp) {
  const xntt::F0 f{};
  u64 h1, h2;
  lab::mont_parts(f, x1, w, wp, h1, h2);
  const u64 u = h1 - h2;
  const u64 a = x0;
  x0 = a + u;
  x1 = a - u;
}
// v7: cost probe: canonical Montgomery product only (x0 passes through an xor so that it stays live)
__device__ __forceinline__ void bf_v7(u64& x0, u64& x1, u64 w, u64 wp) {
  const xntt::F0 f{};
  const u64 u = f.mont(x1, w, wp);
  x1 = x0 ^ u;
  x0 = u;
}
// v8: v0 with the full low product instead of IMAD.HI (no addend pair to build)
__device__ __forceinline__ void bf_v8(u64& x0, u64& x1, u64 w, u64 wp) {
  u32 a0, a1, w0, w1, q0, q1, h1l, h1h, h2l, h2h;
  unpack64(x1, a0, a1);
  unpack64(w, w0, w1);
  unpack64(x1 * wp, q0, q1);
  const u64 A = (u64)a0 * w0;
  u32 Al, Ah;
  unpack64(A, Al, Ah);
  asm("{\n\t.reg .u32 xl, xh, xc, lh, yl, yh, yc, t;\n\t"
      "mul.lo.u32 xl, %4, %7;\n\tmul.hi.u32 xh, %4, %7;\n\t"
      "mad.lo.cc.u32 xl, %5, %6, xl;\n\tmadc.hi.cc.u32 xh, %5, %6, xh;\n\taddc.u32 xc, 0, 0;\n\t"
      "add.cc.u32 lh, xl, %12;\n\t"
      "madc.lo.cc.u32 %0, %5, %7, xh;\n\tmadc.hi.u32 %1, %5, %7, xc;\n\t"
      "mul.lo.u32 yl, %8, %11;\n\tmul.hi.u32 yh, %8, %11;\n\t"
      "mad.lo.cc.u32 yl, %9, %10, yl;\n\tmadc.hi.cc.u32 yh, %9, %10, yh;\n\taddc.u32 yc, 0, 0;\n\t"
      "not.b32 t, lh;\n\tadd.cc.u32 t, yl, t;\n\t"
      "madc.lo.cc.u32 %2, %9, %11, yh;\n\tmadc.hi.u32 %3, %9, %11, yc;\n\t"
      "}"
      : "=r"(h1l), "=r"(h1h), "=r"(h2l), "=r"(h2h)
      : "r"(a0), "r"(a1), "r"(w0), "r"(w1), "r"(q0), "r"(q1), "r"(P_LO), "r"(P_HI), "r"(Ah), "r"(Al));
  const xntt::F0 f{};
  u64 u, s, d;
  u32 m, d0, d1;
  xntt::sub_borrow_mask(pack64(h1l, h1h), pack64(h2l, h2h), u, m);
  xntt::add_carry_plus(x0, u, m, s, d0);
  xntt::sub_borrow_minus(x0, u, m, d, d1);
  x0 = f.fix(s, d0);
  x1 = f.fix(d, d1);
}

// ALU-side repair v + delta*C, delta in {-1,0,1}: only the 32x10-bit product stays on the fma pipe
__device__ __forceinline__ u64 fix_alu(u64 v, u32 delta) {
  u32 vl, vh, rl, rh;
  unpack64(v, vl, vh);
  const u32 eps = 0u - delta;
  const u32 lo_add = (eps << 31) + eps;                    // delta * 0x7fffffff mod 2^32
  const u32 t = delta * C_HI + vh;                         // IMAD
  const u32 sgn = (u32)((int)delta >> 31);                 // -[delta < 0]
  asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, %5;" : "=r"(rl), "=r"(rh) : "r"(vl), "r"(lo_add), "r"(t), "r"(sgn));
  return pack64(rl, rh);
}
template <int MODE>
__device__ __forceinline__ void bf_mixed(u64& x0, u64& x1, u64 w, u64 wp) {
  const xntt::F0 f{};
  u64 h1, h2, u, s, d;
  u32 m, d0, d1;
  lab::mont_parts(f, x1, w, wp, h1, h2);
  xntt::sub_borrow_mask(h1, h2, u, m);
  xntt::add_carry_plus(x0, u, m, s, d0);
  xntt::sub_borrow_minus(x0, u, m, d, d1);
  x0 = (MODE & 1) ? fix_alu(s, d0) : f.fix(s, d0);
  x1 = (MODE & 2) ? fix_alu(d, d1) : f.fix(d, d1);
}
}

namespace lab {
// v12: the second word of the low product taken from the q*P side.  The low 64 bits of a*w and q*P coincide, so
// L.hi = lo32(q0*P1 + q1*P0) + hi32(q0*P0) as well, and with P0 = 2^31 + 1 that last term is shifts and one add
// instead of the IMAD.WIDE a0*w0:  hi32(q0*(2^31+1)) = (q0 >> 1) + carry(q0 + (q0 << 31)).
// carry2 is then the plain carry of yl + vh', carry1 = [L.hi < xl] (the roles of v8 swapped).  9 wide + 4 narrow.
__device__ __forceinline__ void mont_parts_v12(u64 a, u64 w, u64 wp, u64& h1, u64& h2) {
  u32 a0, a1, w0, w1, q0, q1, h1l, h1h, h2l, h2h;
  unpack64(a, a0, a1);
  unpack64(w, w0, w1);
  unpack64(a * wp, q0, q1);
  u32 vhp;  // explicit carry chain: written as C, ptxas folds it back into IMAD.HI q0 * 0x80000001
  asm("{\n\t.reg .u32 s, h;\n\tshl.b32 s, %1, 31;\n\tshr.u32 h, %1, 1;\n\tadd.cc.u32 s, s, %1;\n\taddc.u32 %0, h, 0;\n\t}"
      : "=r"(vhp) : "r"(q0));
  asm("{\n\t.reg .u32 xl, xh, xc, lh, yl, yh, yc, t;\n\t"
      "mul.lo.u32 yl, %8, %11;\n\tmul.hi.u32 yh, %8, %11;\n\t"  // q0*P1
      "mad.lo.cc.u32 yl, %9, %10, yl;\n\tmadc.hi.cc.u32 yh, %9, %10, yh;\n\taddc.u32 yc, 0, 0;\n\t"  // + q1*P0
      "add.cc.u32 lh, yl, %12;\n\t"  // L.hi, carry2
      "madc.lo.cc.u32 %2, %9, %11, yh;\n\tmadc.hi.u32 %3, %9, %11, yc;\n\t"  // h2
      "mul.lo.u32 xl, %4, %7;\n\tmul.hi.u32 xh, %4, %7;\n\t"  // a0*w1
      "mad.lo.cc.u32 xl, %5, %6, xl;\n\tmadc.hi.cc.u32 xh, %5, %6, xh;\n\taddc.u32 xc, 0, 0;\n\t"  // + a1*w0
      "not.b32 t, lh;\n\tadd.cc.u32 t, xl, t;\n\t"  // carry1 = [xl > L.hi]
      "madc.lo.cc.u32 %0, %5, %7, xh;\n\tmadc.hi.u32 %1, %5, %7, xc;\n\t"  // h1
      "}"
      : "=r"(h1l), "=r"(h1h), "=r"(h2l), "=r"(h2h)
      : "r"(a0), "r"(a1), "r"(w0), "r"(w1), "r"(q0), "r"(q1), "r"(P_LO), "r"(P_HI), "r"(vhp));
  h1 = pack64(h1l, h1h);
  h2 = pack64(h2l, h2h);
}
__device__ __forceinline__ void bf_v12(u64& x0, u64& x1, u64 w, u64 wp) {
  const xntt::F0 f{};
  u64 h1, h2, u, s, d;
  u32 m, d0, d1;
  mont_parts_v12(x1, w, wp, h1, h2);
  xntt::sub_borrow_mask(h1, h2, u, m);
  xntt::add_carry_plus(x0, u, m, s, d0);
  xntt::sub_borrow_minus(x0, u, m, d, d1);
  x0 = f.fix(s, d0);
  x1 = f.fix(d, d1);
}
// v13: v12 with q1*P0 = q1 + (q1 << 31) by shifts as well (8 wide): y = q0*P1 + q1*P0 as a 65-bit sum
__device__ __forceinline__ void bf_v13(u64& x0, u64& x1, u64 w, u64 wp) {
  const xntt::F0 f{};
  u32 a0, a1, w0, w1, q0, q1, h1l, h1h, h2l, h2h;
  unpack64(x1, a0, a1);
  unpack64(w, w0, w1);
  unpack64(x1 * wp, q0, q1);
  u32 vhp, zl, zh;
  asm("{\n\t.reg .u32 s, h;\n\tshl.b32 s, %1, 31;\n\tshr.u32 h, %1, 1;\n\tadd.cc.u32 s, s, %1;\n\taddc.u32 %0, h, 0;\n\t}"
      : "=r"(vhp) : "r"(q0));
  // z = q1*P0 = q1 + (q1 << 31) < 2^63 + 2^32
  asm("{\n\t.reg .u32 s, h;\n\tshl.b32 s, %2, 31;\n\tshr.u32 h, %2, 1;\n\tadd.cc.u32 %0, s, %2;\n\taddc.u32 %1, h, 0;\n\t}"
      : "=r"(zl), "=r"(zh) : "r"(q1));
  asm("{\n\t.reg .u32 xl, xh, xc, lh, yl, yh, yc, t;\n\t"
      "mad.lo.cc.u32 yl, %8, %11, %13;\n\tmadc.hi.cc.u32 yh, %8, %11, %14;\n\taddc.u32 yc, 0, 0;\n\t"  // q0*P1 + z
      "add.cc.u32 lh, yl, %12;\n\t"
      "madc.lo.cc.u32 %2, %9, %11, yh;\n\tmadc.hi.u32 %3, %9, %11, yc;\n\t"
      "mul.lo.u32 xl, %4, %7;\n\tmul.hi.u32 xh, %4, %7;\n\t"
      "mad.lo.cc.u32 xl, %5, %6, xl;\n\tmadc.hi.cc.u32 xh, %5, %6, xh;\n\taddc.u32 xc, 0, 0;\n\t"
      "not.b32 t, lh;\n\tadd.cc.u32 t, xl, t;\n\t"
      "madc.lo.cc.u32 %0, %5, %7, xh;\n\tmadc.hi.u32 %1, %5, %7, xc;\n\t"
      "}"
      : "=r"(h1l), "=r"(h1h), "=r"(h2l), "=r"(h2h)
      : "r"(a0), "r"(a1), "r"(w0), "r"(w1), "r"(q0), "r"(q1), "r"(P_LO), "r"(P_HI), "r"(vhp), "r"(zl), "r"(zh));
  u64 u, s, d;
  u32 m, d0, d1;
  xntt::sub_borrow_mask(pack64(h1l, h1h), pack64(h2l, h2h), u, m);
  xntt::add_carry_plus(x0, u, m, s, d0);
  xntt::sub_borrow_minus(x0, u, m, d, d1);
  x0 = f.fix(s, d0);
  x1 = f.fix(d, d1);
}
}  // namespace lab
