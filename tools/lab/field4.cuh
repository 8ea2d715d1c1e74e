// SPDX-License-Identifier: Apache-2.0
// Development lab (not part of the product): butterfly variants that ask for three-operand 64-bit additions
// (x0 + h1 - h2 and x0 - h1 + h2 directly, with the net carry count as a third word) instead of u = h1 - h2 followed by
// x0 +- u.  SASS has IADD3 / IADD3.X with two carry-outs / carry-ins, which would do each of the two in 3 instructions
// (6 instead of 9 per butterfly); PTX has no such instruction, so whether ptxas forms them is what is being probed.
#pragma once
#include "field.cuh"

namespace lab {
using xntt::u32;
using xntt::u64;

// v23: 128-bit C arithmetic
__device__ __forceinline__ void bf_v23(u64& x0, u64& x1, u64 w, u64 wp) {
  const xntt::F0 f{};
  u64 h1, h2;
  lab::mont_parts(f, x1, w, wp, h1, h2);
  const __int128 S = (__int128)(unsigned __int128)x0 + (__int128)(unsigned __int128)h1 - (__int128)(unsigned __int128)h2;
  const __int128 D = (__int128)(unsigned __int128)x0 - (__int128)(unsigned __int128)h1 + (__int128)(unsigned __int128)h2;
  x0 = f.fix((u64)S, (u32)(u64)(S >> 64));
  x1 = f.fix((u64)D, (u32)(u64)(D >> 64));
}

// v24: 96-bit arithmetic in 32-bit words, two chained two-operand carry chains per output written so that ptxas may
// merge them
__device__ __forceinline__ void bf_v24(u64& x0, u64& x1, u64 w, u64 wp) {
  const xntt::F0 f{};
  u64 h1, h2;
  lab::mont_parts(f, x1, w, wp, h1, h2);
  u32 al, ah, bl, bh, cl, ch, sl, sh, sk, dl, dh, dk;
  xntt::unpack64(x0, al, ah);
  xntt::unpack64(h1, bl, bh);
  xntt::unpack64(h2, cl, ch);
  asm("{\n\t.reg .u32 tl, th, tk;\n\t"
      "add.cc.u32 tl, %3, %5;\n\taddc.cc.u32 th, %4, %6;\n\taddc.u32 tk, 0, 0;\n\t"
      "sub.cc.u32 %0, tl, %7;\n\tsubc.cc.u32 %1, th, %8;\n\tsubc.u32 %2, tk, 0;\n\t}"
      : "=r"(sl), "=r"(sh), "=r"(sk)
      : "r"(al), "r"(ah), "r"(bl), "r"(bh), "r"(cl), "r"(ch));
  asm("{\n\t.reg .u32 tl, th, tk;\n\t"
      "add.cc.u32 tl, %3, %7;\n\taddc.cc.u32 th, %4, %8;\n\taddc.u32 tk, 0, 0;\n\t"
      "sub.cc.u32 %0, tl, %5;\n\tsubc.cc.u32 %1, th, %6;\n\tsubc.u32 %2, tk, 0;\n\t}"
      : "=r"(dl), "=r"(dh), "=r"(dk)
      : "r"(al), "r"(ah), "r"(bl), "r"(bh), "r"(cl), "r"(ch));
  x0 = f.fix(xntt::pack64(sl, sh), sk);
  x1 = f.fix(xntt::pack64(dl, dh), dk);
}
}  // namespace lab

// v25..: keep the glue of the butterfly on the alu pipe.  ptxas equalises the instruction COUNTS of the fma and alu pipes
// by re-homing moves and two-operand additions as IMAD.MOV / IMAD.IADD / IMAD.X, as if a 32x32->64 product cost the fma
// pipe what an IMAD does; on this part it costs 2.3 times that, so every re-homed instruction lengthens the bound pipe.
// Forms ptxas cannot re-home: PRMT with a selector it does not know (an identity byte permutation read from constant
// memory instead of a MOV), three-operand IADD3, LOP3, SEL.
namespace lab {
__constant__ u32 k_id_perm = 0x3210u;  // not a compile-time constant for ptxas (cudaMemcpyToSymbol could change it)

template <int MODE>
__device__ __forceinline__ void mont_parts_v25(u64 a, u64 w, u64 wp, u64& h1, u64& h2) {
  constexpr u64 P = xntt::kP0;
  constexpr u32 P_LO = (u32)P, P_HI = (u32)(P >> 32);
  u32 a0, a1, w0, w1, wp0, wp1, q0, q1, h1l, h1h, h2l, h2h, vl, vh;
  xntt::unpack64(a, a0, a1);
  xntt::unpack64(w, w0, w1);
  xntt::unpack64(wp, wp0, wp1);
  // q = a * wp mod 2^64 with the two narrow products chained through the addend (no separate add)
  xntt::unpack64((u64)a0 * wp0, q0, q1);
  asm("mad.lo.u32 %0, %1, %2, %0;\n\tmad.lo.u32 %0, %3, %4, %0;" : "+r"(q1) : "r"(a0), "r"(wp1), "r"(a1), "r"(wp0));
  xntt::unpack64((u64)a0 * w0, vl, vh);
  (void)vl;
  const u32 sel = k_id_perm;
  if constexpr (MODE == 0) {
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
        : "r"(a0), "r"(a1), "r"(w0), "r"(w1), "r"(q0), "r"(q1), "r"(P_LO), "r"(P_HI), "r"(vh));
  } else {
    // the high word of each cross sum reaches its addend pair through PRMT instead of MOV
    asm("{\n\t.reg .u32 xl, xh, xc, lh, yl, yh, yc, t, xh2, yh2;\n\t"
        "mul.lo.u32 xl, %4, %7;\n\tmul.hi.u32 xh, %4, %7;\n\t"
        "mad.lo.cc.u32 xl, %5, %6, xl;\n\tmadc.hi.cc.u32 xh, %5, %6, xh;\n\taddc.u32 xc, 0, 0;\n\t"
        "prmt.b32 xh2, xh, xh, %13;\n\t"
        "add.cc.u32 lh, xl, %12;\n\t"
        "madc.lo.cc.u32 %0, %5, %7, xh2;\n\tmadc.hi.u32 %1, %5, %7, xc;\n\t"
        "mul.lo.u32 yl, %8, %11;\n\tmul.hi.u32 yh, %8, %11;\n\t"
        "mad.lo.cc.u32 yl, %9, %10, yl;\n\tmadc.hi.cc.u32 yh, %9, %10, yh;\n\taddc.u32 yc, 0, 0;\n\t"
        "prmt.b32 yh2, yh, yh, %13;\n\t"
        "not.b32 t, lh;\n\tadd.cc.u32 t, yl, t;\n\t"
        "madc.lo.cc.u32 %2, %9, %11, yh2;\n\tmadc.hi.u32 %3, %9, %11, yc;\n\t"
        "}"
        : "=r"(h1l), "=r"(h1h), "=r"(h2l), "=r"(h2h)
        : "r"(a0), "r"(a1), "r"(w0), "r"(w1), "r"(q0), "r"(q1), "r"(P_LO), "r"(P_HI), "r"(vh), "r"(sel));
  }
  h1 = xntt::pack64(h1l, h1h);
  h2 = xntt::pack64(h2l, h2h);
}

template <int MODE>
__device__ __forceinline__ void bf_v25(u64& x0, u64& x1, u64 w, u64 wp) {
  const xntt::F0 f{};
  u64 h1, h2, u, s, d;
  u32 m, d0, d1;
  mont_parts_v25<MODE>(x1, w, wp, h1, h2);
  xntt::sub_borrow_mask(h1, h2, u, m);
  xntt::add_carry_plus(x0, u, m, s, d0);
  xntt::sub_borrow_minus(x0, u, m, d, d1);
  x0 = f.fix(s, d0);
  x1 = f.fix(d, d1);
}
}  // namespace lab

namespace lab {
__constant__ u32 k_zero = 0u;  // a zero ptxas does not know

// v27: v26 plus the carry words delta0 = carry(s) - br, delta1 = br - borrow(d) built from three 0/1 flags (SEL) and two
// three-operand additions (third operand: the unknown zero), which cannot be re-homed as IMAD.X / IMAD.IADD
__device__ __forceinline__ void bf_v27(u64& x0, u64& x1, u64 w, u64 wp) {
  const xntt::F0 f{};
  u64 h1, h2;
  mont_parts_v25<1>(x1, w, wp, h1, h2);
  u32 al, ah, bl, bh, cl, ch, ul, uh, sl, sh, dl, dh, br, cs, bd, d0, d1;
  xntt::unpack64(x0, al, ah);
  xntt::unpack64(h1, bl, bh);
  xntt::unpack64(h2, cl, ch);
  const u32 z = k_zero;
  asm("sub.cc.u32 %0, %3, %5;\n\tsubc.cc.u32 %1, %4, %6;\n\taddc.u32 %2, 0, 0;"
      : "=r"(ul), "=r"(uh), "=r"(br) : "r"(bl), "r"(bh), "r"(cl), "r"(ch));
  asm("add.cc.u32 %0, %3, %5;\n\taddc.cc.u32 %1, %4, %6;\n\taddc.u32 %2, 0, 0;"
      : "=r"(sl), "=r"(sh), "=r"(cs) : "r"(al), "r"(ah), "r"(ul), "r"(uh));
  asm("sub.cc.u32 %0, %3, %5;\n\tsubc.cc.u32 %1, %4, %6;\n\taddc.u32 %2, 0, 0;"
      : "=r"(dl), "=r"(dh), "=r"(bd) : "r"(al), "r"(ah), "r"(ul), "r"(uh));
  d0 = cs - br + z;
  d1 = br - bd + z;
  x0 = f.fix(xntt::pack64(sl, sh), d0);
  x1 = f.fix(xntt::pack64(dl, dh), d1);
}
}  // namespace lab

namespace lab {
// v28: two instructions less per butterfly, both inside what PTX can say.
//  * q = a * w' mod 2^64 with its two narrow products chained through the addend (v25);
//  * carry2 = [L.hi < lo32(q0 P1 + q1 P0)] is the borrow of L.hi - yl, and a borrow is what the subtraction h1 - h2 that
//    follows can take as its borrow-in: h2' = q1 P1 + {yh, yc} without the carry, u = h1 - h2' - carry2 in the same two
//    subc - no NOT, no carry-in on the last product.  (sub.cc feeding subc only: the pairing ptxas 12.9 gets wrong is
//    sub.cc feeding madc / addc.)
__device__ __forceinline__ void mont_diff_v28(u64 a, u64 w, u64 wp, u64& u, u32& m) {
  constexpr u64 P = xntt::kP0;
  constexpr u32 P_LO = (u32)P, P_HI = (u32)(P >> 32);
  u32 a0, a1, w0, w1, wp0, wp1, q0, q1, ul, uh, vl, vh;
  xntt::unpack64(a, a0, a1);
  xntt::unpack64(w, w0, w1);
  xntt::unpack64(wp, wp0, wp1);
  xntt::unpack64((u64)a0 * wp0, q0, q1);
  asm("mad.lo.u32 %0, %1, %2, %0;\n\tmad.lo.u32 %0, %3, %4, %0;" : "+r"(q1) : "r"(a0), "r"(wp1), "r"(a1), "r"(wp0));
  xntt::unpack64((u64)a0 * w0, vl, vh);
  (void)vl;
  asm("{\n\t.reg .u32 xl, xh, xc, lh, yl, yh, yc, t, h1l, h1h, h2l, h2h;\n\t"
      "mul.lo.u32 xl, %3, %6;\n\tmul.hi.u32 xh, %3, %6;\n\t"  // a0*w1
      "mad.lo.cc.u32 xl, %4, %5, xl;\n\tmadc.hi.cc.u32 xh, %4, %5, xh;\n\taddc.u32 xc, 0, 0;\n\t"  // + a1*w0
      "add.cc.u32 lh, xl, %11;\n\t"    // L.hi, carry1
      "madc.lo.cc.u32 h1l, %4, %6, xh;\n\tmadc.hi.u32 h1h, %4, %6, xc;\n\t"  // h1 = a1*w1 + {xh, xc} + carry1
      "mul.lo.u32 yl, %7, %10;\n\tmul.hi.u32 yh, %7, %10;\n\t"  // q0*P1
      "mad.lo.cc.u32 yl, %8, %9, yl;\n\tmadc.hi.cc.u32 yh, %8, %9, yh;\n\taddc.u32 yc, 0, 0;\n\t"  // + q1*P0
      "mad.lo.cc.u32 h2l, %8, %10, yh;\n\tmadc.hi.u32 h2h, %8, %10, yc;\n\t"  // h2' = q1*P1 + {yh, yc}
      "sub.cc.u32 t, lh, yl;\n\t"  // borrow = carry2
      "subc.cc.u32 %0, h1l, h2l;\n\tsubc.cc.u32 %1, h1h, h2h;\n\tsubc.u32 %2, 0, 0;\n\t"
      "}"
      : "=r"(ul), "=r"(uh), "=r"(m)
      : "r"(a0), "r"(a1), "r"(w0), "r"(w1), "r"(q0), "r"(q1), "r"(P_LO), "r"(P_HI), "r"(vh));
  u = xntt::pack64(ul, uh);
}
__device__ __forceinline__ void bf_v28(u64& x0, u64& x1, u64 w, u64 wp) {
  const xntt::F0 f{};
  u64 u, s, d;
  u32 m, d0, d1;
  mont_diff_v28(x1, w, wp, u, m);
  xntt::add_carry_plus(x0, u, m, s, d0);
  xntt::sub_borrow_minus(x0, u, m, d, d1);
  x0 = f.fix(s, d0);
  x1 = f.fix(d, d1);
}
}  // namespace lab

namespace lab {
// v29: v28 with the borrow of h1 - h2 kept as nb = 1 - br (one subc, like the mask before), so that both carry words come
// out of one instruction each without a NOT:  delta0 = carry(s) - br = nb - 1 + carry(s)  (addc nb, -1),
// delta1 = br - borrow(d) = 1 - nb - borrow(d)  (subc 1, nb).
__device__ __forceinline__ void mont_diff_v29(u64 a, u64 w, u64 wp, u64& u, u32& nb) {
  constexpr u64 P = xntt::kP0;
  constexpr u32 P_LO = (u32)P, P_HI = (u32)(P >> 32);
  u32 a0, a1, w0, w1, wp0, wp1, q0, q1, ul, uh, vl, vh;
  xntt::unpack64(a, a0, a1);
  xntt::unpack64(w, w0, w1);
  xntt::unpack64(wp, wp0, wp1);
  xntt::unpack64((u64)a0 * wp0, q0, q1);
  asm("mad.lo.u32 %0, %1, %2, %0;\n\tmad.lo.u32 %0, %3, %4, %0;" : "+r"(q1) : "r"(a0), "r"(wp1), "r"(a1), "r"(wp0));
  xntt::unpack64((u64)a0 * w0, vl, vh);
  (void)vl;
  asm("{\n\t.reg .u32 xl, xh, xc, lh, yl, yh, yc, t, h1l, h1h, h2l, h2h;\n\t"
      "mul.lo.u32 xl, %3, %6;\n\tmul.hi.u32 xh, %3, %6;\n\t"
      "mad.lo.cc.u32 xl, %4, %5, xl;\n\tmadc.hi.cc.u32 xh, %4, %5, xh;\n\taddc.u32 xc, 0, 0;\n\t"
      "add.cc.u32 lh, xl, %11;\n\t"
      "madc.lo.cc.u32 h1l, %4, %6, xh;\n\tmadc.hi.u32 h1h, %4, %6, xc;\n\t"
      "mul.lo.u32 yl, %7, %10;\n\tmul.hi.u32 yh, %7, %10;\n\t"
      "mad.lo.cc.u32 yl, %8, %9, yl;\n\tmadc.hi.cc.u32 yh, %8, %9, yh;\n\taddc.u32 yc, 0, 0;\n\t"
      "mad.lo.cc.u32 h2l, %8, %10, yh;\n\tmadc.hi.u32 h2h, %8, %10, yc;\n\t"
      "sub.cc.u32 t, lh, yl;\n\t"
      "subc.cc.u32 %0, h1l, h2l;\n\tsubc.cc.u32 %1, h1h, h2h;\n\tsubc.u32 %2, 0, 0xffffffff;\n\t"
      "}"
      : "=r"(ul), "=r"(uh), "=r"(nb)
      : "r"(a0), "r"(a1), "r"(w0), "r"(w1), "r"(q0), "r"(q1), "r"(P_LO), "r"(P_HI), "r"(vh));
  u = xntt::pack64(ul, uh);
}
__device__ __forceinline__ void bf_v29(u64& x0, u64& x1, u64 w, u64 wp) {
  const xntt::F0 f{};
  u64 u;
  u32 nb, al, ah, bl, bh, sl, sh, dl, dh, d0, d1;
  mont_diff_v29(x1, w, wp, u, nb);
  xntt::unpack64(x0, al, ah);
  xntt::unpack64(u, bl, bh);
  asm("add.cc.u32 %0, %3, %5;\n\taddc.cc.u32 %1, %4, %6;\n\taddc.u32 %2, %7, 0xffffffff;"
      : "=r"(sl), "=r"(sh), "=r"(d0) : "r"(al), "r"(ah), "r"(bl), "r"(bh), "r"(nb));
  asm("sub.cc.u32 %0, %3, %5;\n\tsubc.cc.u32 %1, %4, %6;\n\tsubc.u32 %2, 1, %7;"
      : "=r"(dl), "=r"(dh), "=r"(d1) : "r"(al), "r"(ah), "r"(bl), "r"(bh), "r"(nb));
  x0 = f.fix(xntt::pack64(sl, sh), d0);
  x1 = f.fix(xntt::pack64(dl, dh), d1);
}
}  // namespace lab
