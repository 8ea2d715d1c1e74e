// SPDX-License-Identifier: Apache-2.0
// Development lab (not part of the product): the two-sided Montgomery product the kernels used until round 3 -
// h1 = hi64(a*w) and h2 = hi64(q*P) completed separately, then subtracted (sub_borrow_mask) - which the older lab
// variants are written against.  The product uses FieldOps::mont_diff (field.cuh).
#pragma once
#include "field.cuh"

namespace xntt {
// (a - b) mod 2^64 and m = -borrow (IADD3, IADD3.X, IADD3.X)
__device__ __forceinline__ void sub_borrow_mask(u64 a, u64 b, u64& d, u32& m) {
  u32 al, ah, bl, bh, dl, dh;
  unpack64(a, al, ah);
  unpack64(b, bl, bh);
  asm("sub.cc.u32 %0, %3, %5;\n\tsubc.cc.u32 %1, %4, %6;\n\tsubc.u32 %2, 0, 0;"
      : "=r"(dl), "=r"(dh), "=r"(m)
      : "r"(al), "r"(ah), "r"(bl), "r"(bh));
  d = pack64(dl, dh);
}
}  // namespace xntt

namespace lab {
using xntt::u32;
using xntt::u64;
using xntt::pack64;
using xntt::unpack64;

// returns h1 = hi64(a*w), h2 = hi64(q*P) with q = a*w' mod 2^64; a*omega == h1 - h2 (mod P)
template <class F>
__device__ __forceinline__ void mont_parts(const F& f, u64 a, u64 w, u64 wp, u64& h1, u64& h2) {
  const u64 P = f.p();
  const u32 P_LO = (u32)P, P_HI = (u32)(P >> 32);
    u32 a0, a1, w0, w1, q0, q1, h1l, h1h, h2l, h2h, vl, vh;
    unpack64(a, a0, a1);
    unpack64(w, w0, w1);
    unpack64(a * wp, q0, q1);
    // a0*w0 as a full IMAD.WIDE: costs the fma pipe what IMAD.HI does, but spares the (0 : xl) addend pair
    // ptxas builds for the IMAD.HI form (butterfly loop: 53.2 instead of 55.2 fma-pipe cycles, measured 3 % faster)
    unpack64((u64)a0 * w0, vl, vh);
    (void)vl;
    asm("{\n\t.reg .u32 xl, xh, xc, lh, yl, yh, yc, t;\n\t"
        "mul.lo.u32 xl, %4, %7;\n\tmul.hi.u32 xh, %4, %7;\n\t"  // a0*w1
        "mad.lo.cc.u32 xl, %5, %6, xl;\n\tmadc.hi.cc.u32 xh, %5, %6, xh;\n\taddc.u32 xc, 0, 0;\n\t"  // + a1*w0
        "add.cc.u32 lh, xl, %12;\n\t"    // L.hi, carry1
        "madc.lo.cc.u32 %0, %5, %7, xh;\n\tmadc.hi.u32 %1, %5, %7, xc;\n\t"  // h1 = a1*w1 + {xh, xc} + carry1
        "mul.lo.u32 yl, %8, %11;\n\tmul.hi.u32 yh, %8, %11;\n\t"  // q0*P1
        "mad.lo.cc.u32 yl, %9, %10, yl;\n\tmadc.hi.cc.u32 yh, %9, %10, yh;\n\taddc.u32 yc, 0, 0;\n\t"  // + q1*P0
        // carry2 = [lh < yl] as the carry of yl + ~lh  (do NOT use sub.cc -> madc here: ptxas 12.9
        // feeds the IADD3 carry-out, i.e. NOT borrow, straight into IMAD.WIDE.X)
        "not.b32 t, lh;\n\tadd.cc.u32 t, yl, t;\n\t"
        "madc.lo.cc.u32 %2, %9, %11, yh;\n\tmadc.hi.u32 %3, %9, %11, yc;\n\t"  // h2 = q1*P1 + {yh, yc} + carry2
        "}"
        : "=r"(h1l), "=r"(h1h), "=r"(h2l), "=r"(h2h)
        : "r"(a0), "r"(a1), "r"(w0), "r"(w1), "r"(q0), "r"(q1), "r"(P_LO), "r"(P_HI), "r"(vh));
    h1 = pack64(h1l, h1h);
    h2 = pack64(h2l, h2h);
}
}  // namespace lab
