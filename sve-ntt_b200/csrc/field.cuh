// SPDX-License-Identifier: Apache-2.0
//
// PAdic64 device arithmetic for sm_100a: Montgomery multiplication with R = 2^64 built from
// 32-bit IMAD / IMAD.WIDE chains, plus the "lazy" add/sub used by the butterflies.
//
// Values ("residues") live in one of two domains:
//   canonical : v in [0, P)
//   lazy      : any v in [0, 2^64), standing for v mod P.  Because P + C = 2^64 (C = 2^64 - P),
//               a single carry/borrow out of a 64-bit add/sub is repaired by adding/subtracting C.
//
// What this replaces in the reference (value-for-value, not instruction-for-instruction):
//   sventt::PAdic64SVE::multiply            include/sventt/modmul/sve/p-adic-64.hpp:80-95
//   sventt::PAdic64SVE::multiply_normalize  include/sventt/modmul/sve/p-adic-64.hpp:101-115
//   sventt::PAdic64SVE::butterfly_inverse   include/sventt/modmul/sve/p-adic-64.hpp:229-246
//   sventt::PAdic64SVE::precompute          include/sventt/modmul/sve/p-adic-64.hpp:64-74
// Twiddles are kept, like the reference does, as a pair (w, w') with w = omega * 2^64 mod P
// (Montgomery form) and w' = w * P^-1 mod 2^64, so that mont(a, w, w') = a * omega mod P with the
// data staying in the normal domain.
#pragma once
#include <cstdint>

#include "params.h"

#if defined(XNTT_HOST_EMU)
#define __device__
#define __host__
#define __forceinline__ inline
#define __align__(n) alignas(n)
#endif

namespace xntt {

__host__ __device__ constexpr u64 montgomery_inverse(u64 p) {
  // Newton iteration for p^-1 mod 2^64 (p odd); same value as
  // sventt::Modulus::get_montgomery_inverse (include/sventt/modulus.hpp:36-68).
  u64 x = p;  // correct to 3 bits
  for (int i = 0; i < 6; ++i) x *= 2 - p * x;
  return x;
}

#if defined(XNTT_HOST_EMU)
// Host emulation of the PTX primitives (tests/emu only: lets the CPU test-suite execute the very
// same kernel templates thread by thread; never part of the product path).
inline u64 pack64(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }
inline void unpack64(u64 v, u32& lo, u32& hi) {
  lo = (u32)v;
  hi = (u32)(v >> 32);
}
inline u64 mulhi64(u64 a, u64 b) { return (u64)(((unsigned __int128)a * b) >> 64); }
// (a + b) mod 2^64 and carry + addend
inline void add_carry_plus(u64 a, u64 b, u32 addend, u64& s, u32& k) {
  s = a + b;
  k = addend + (s < a ? 1u : 0u);
}
// (a - b) mod 2^64 and -sub - borrow
inline void sub_borrow_minus(u64 a, u64 b, u32 sub, u64& d, u32& k) {
  d = a - b;
  k = 0u - sub - (a < b ? 1u : 0u);
}
#else
__device__ __forceinline__ u64 pack64(u32 lo, u32 hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void unpack64(u64 v, u32& lo, u32& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ u64 mulhi64(u64 a, u64 b) { return __umul64hi(a, b); }
// (a + b) mod 2^64 and k = addend + carry
__device__ __forceinline__ void add_carry_plus(u64 a, u64 b, u32 addend, u64& s, u32& k) {
  u32 al, ah, bl, bh, sl, sh;
  unpack64(a, al, ah);
  unpack64(b, bl, bh);
  asm("add.cc.u32 %0, %3, %5;\n\taddc.cc.u32 %1, %4, %6;\n\taddc.u32 %2, %7, 0;"
      : "=r"(sl), "=r"(sh), "=r"(k)
      : "r"(al), "r"(ah), "r"(bl), "r"(bh), "r"(addend));
  s = pack64(sl, sh);
}
// (a - b) mod 2^64 and k = -sub - borrow
__device__ __forceinline__ void sub_borrow_minus(u64 a, u64 b, u32 sub, u64& d, u32& k) {
  u32 al, ah, bl, bh, dl, dh;
  unpack64(a, al, ah);
  unpack64(b, bl, bh);
  asm("sub.cc.u32 %0, %3, %5;\n\tsubc.cc.u32 %1, %4, %6;\n\tsubc.u32 %2, 0, %7;"
      : "=r"(dl), "=r"(dh), "=r"(k)
      : "r"(al), "r"(ah), "r"(bl), "r"(bh), "r"(sub));
  d = pack64(dl, dh);
}
#endif

// ------------------------------------------------------------------------------------------------
// Constant providers.  StaticModulus bakes the modulus into the instruction stream (immediates);
// RuntimeModulus carries it in registers / the kernel parameter bank so that one set of kernels
// serves every odd prime below 2^64 (sventt::Modulus<p, g> is a template, i.e. any p).
template <u64 P_>
struct StaticModulus {
  static constexpr bool kStatic = true;
  static constexpr u64 P = P_;
  __host__ __device__ constexpr u64 p() const { return P_; }
  __host__ __device__ constexpr u64 c() const { return 0 - P_; }  // 2^64 - P
  __host__ __device__ constexpr u64 pinv() const { return montgomery_inverse(P_); }
  // 2^64 mod P (Montgomery form of 1)
  __host__ __device__ constexpr u64 one() const { return (u64)((((unsigned __int128)1) << 64) % P_); }
};

struct RuntimeModulus {
  static constexpr bool kStatic = false;
  FieldConsts k;
  __host__ __device__ u64 p() const { return k.p; }
  __host__ __device__ u64 c() const { return 0 - k.p; }
  __host__ __device__ u64 pinv() const { return k.pinv; }
  __host__ __device__ u64 one() const { return k.one; }
};

// All the lazy-domain identities below only need P < 2^64 odd (P + C = 2^64).
template <class K>
struct FieldOps : K {
  // v + delta * C (mod 2^64) for delta in {-1, 0, +1} held as a 32-bit two's complement value.
  __device__ __forceinline__ u64 fix(u64 v, u32 delta) const {
    const u64 C = this->c();
    const u32 C_LO = (u32)C, C_HI = (u32)(C >> 32);
    if constexpr (K::kStatic) {
      if (C_LO < 0x80000000u) {
        // one signed IMAD.WIDE plus one IMAD
        u64 t = v + (u64)((long long)(int)delta * (long long)(int)C_LO);
        u32 tl, th;
        unpack64(t, tl, th);
        th += delta * C_HI;
        return pack64(tl, th);
      }
    }
    // unsigned form: (2^32-1)*C_LO overshoots -C_LO by C_LO*2^32, taken back out of the high word
    u64 t = v + (u64)delta * (u64)C_LO;
    u32 tl, th;
    unpack64(t, tl, th);
    th += delta * C_HI - (delta >> 31) * C_LO;
    return pack64(tl, th);
  }

  // The Montgomery product as the kernels use it: u = (h1 - h2) mod 2^64 and m = -borrow (0 or 0xffffffff), so that
  // a*omega == u - br * 2^64 (mod P), with h1 = hi64(a*w), h2 = hi64(q*P), q = a*w' mod 2^64: the true difference
  // h1 - h2 lies in (-P, P).  `a` may be lazy.
  //
  // On sm_100a IMAD.WIDE issues at well under half the IMAD rate (measured: 8.0 vs 18.5 Tinstr/s), so the 32x32->64
  // products are what the kernel is bound by.  Seven of them are needed here (four for a*w, three for q*P; the eighth
  // is the low product a*w') instead of the eight of a pair of mul.hi.u64: the low 64 bits of a*w and q*P are equal by
  // construction, so the carry out of the low half of q*P follows from the low half of a*w:
  //   carry2 = [L.hi < lo32(q0*P1 + q1*P0)]  with  L.hi = bits 32..63 of a*w.
  // Two instructions shorter than completing h1 and h2 separately and subtracting them (the form until round 3, kept in
  // tools/lab/field1.cuh; butterfly loop on B200 60.7 -> 57.9 cycles per warp-butterfly, lab v28):
  //   * q = a*w' mod 2^64 with its two narrow products chained through the addend (no separate add);
  //   * carry2 = [L.hi < yl] is the BORROW of L.hi - yl, and a borrow is what the subtraction h1 - h2 that follows takes
  //     as its borrow-in: h2' = q1*P1 + {yh, yc} without the carry, u = h1 - h2' - carry2 in the same two subc - no
  //     NOT, no carry-in on the last product.  sub.cc feeds subc only (the pairing ptxas 12.9 gets wrong is sub.cc
  //     feeding madc).
  __device__ __forceinline__ void mont_diff(u64 a, u64 w, u64 wp, u64& u, u32& m) const {
    const u64 P = this->p();
    const u32 P_LO = (u32)P, P_HI = (u32)(P >> 32);
    (void)P_LO;
    (void)P_HI;
#if defined(XNTT_HOST_EMU)
    // the same partial-product algebra, word by word, in plain C
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), w0 = (u32)w, w1 = (u32)(w >> 32);
    const u32 wp0 = (u32)wp, wp1 = (u32)(wp >> 32);
    const u64 ql = (u64)a0 * wp0;
    const u32 q0 = (u32)ql, q1 = (u32)(ql >> 32) + a0 * wp1 + a1 * wp0;
    const u64 x_a = (u64)a0 * w1, x_b = (u64)a1 * w0;
    const u64 x = x_a + x_b;                       // 65-bit cross sum: carry bit xc
    const u32 xc = x < x_a ? 1u : 0u;
    const u32 vh = (u32)(((u64)a0 * w0) >> 32);
    const u32 lh = (u32)x + vh;                    // L.hi
    const u32 carry1 = lh < vh ? 1u : 0u;
    const u64 h1 = (u64)a1 * w1 + ((x >> 32) | ((u64)xc << 32)) + carry1;
    const u64 y_a = (u64)q0 * P_HI, y_b = (u64)q1 * P_LO;
    const u64 y = y_a + y_b;
    const u32 yc = y < y_a ? 1u : 0u;
    const u64 h2p = (u64)q1 * P_HI + ((y >> 32) | ((u64)yc << 32));  // h2 without carry2
    const u32 carry2 = lh < (u32)y ? 1u : 0u;                         // the borrow of L.hi - yl
    const unsigned __int128 dd = (unsigned __int128)h1 - h2p - carry2;
    u = (u64)dd;
    m = (u64)(dd >> 64) != 0 ? 0xffffffffu : 0u;
#else
    u32 a0, a1, w0, w1, wp0, wp1, q0, q1, ul, uh, vl, vh;
    unpack64(a, a0, a1);
    unpack64(w, w0, w1);
    unpack64(wp, wp0, wp1);
    unpack64((u64)a0 * wp0, q0, q1);
    asm("mad.lo.u32 %0, %1, %2, %0;\n\tmad.lo.u32 %0, %3, %4, %0;" : "+r"(q1) : "r"(a0), "r"(wp1), "r"(a1), "r"(wp0));
    // a0*w0 as a full IMAD.WIDE: costs the fma pipe what IMAD.HI does, but spares the (0 : xl) addend pair ptxas
    // builds for the IMAD.HI form (butterfly loop: 53.2 instead of 55.2 fma-pipe cycles, measured 3 % faster)
    unpack64((u64)a0 * w0, vl, vh);
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
        "subc.cc.u32 %0, h1l, h2l;\n\tsubc.cc.u32 %1, h1h, h2h;\n\tsubc.u32 %2, 0, 0;\n\t"  // u, m = -borrow
        "}"
        : "=r"(ul), "=r"(uh), "=r"(m)
        : "r"(a0), "r"(a1), "r"(w0), "r"(w1), "r"(q0), "r"(q1), "r"(P_LO), "r"(P_HI), "r"(vh));
    u = pack64(ul, uh);
#endif
  }

  // Canonical Montgomery product (reference: multiply_normalize).
  __device__ __forceinline__ u64 mont(u64 a, u64 w, u64 wp) const {
    u64 u;
    u32 m;
    mont_diff(a, w, wp, u, m);
    return fix(u, m);  // borrow -> subtract C, i.e. add P
  }
  __device__ __forceinline__ u64 mont(u64 a, Tw t) const { return mont(a, t.w, t.wp); }

  // w' for a Montgomery-form w that was not precomputed (reference: precompute).
  __device__ __forceinline__ u64 companion(u64 w) const { return w * this->pinv(); }
  // table entry of the twiddle whose Montgomery form is wm
  __device__ __forceinline__ Tw make_tw(u64 wm) const {
    Tw t;
    t.w = wm;
    t.wp = companion(wm);
    return t;
  }

  // lazy -> canonical
  __device__ __forceinline__ u64 canon(u64 v) const {
    const u64 P = this->p();
    if ((P >> 63) != 0) {
      // v >= P  <=>  v + C carries; then v - P == v + C (mod 2^64)
      u64 t;
      u32 c;
      add_carry_plus(v, this->c(), 0u, t, c);
      return c ? t : v;
    }
    // small moduli: v may exceed P several times over; v * 1 through the Montgomery reducer
    const u64 o = this->one();
    return mont(v, o, companion(o));
  }

  // canon(v) when `on`, v itself otherwise - branch-free for the 64-bit moduli (one more predicate input of the select)
  __device__ __forceinline__ u64 canon_if(u64 v, bool on) const {
    const u64 P = this->p();
    if ((P >> 63) != 0) {
      u64 t;
      u32 c;
      add_carry_plus(v, this->c(), 0u, t, c);
      return (c != 0 && on) ? t : v;
    }
    return on ? canon(v) : v;
  }

  // Cooley-Tukey butterfly on lazy values: (x0, x1) <- (x0 + x1*omega, x0 - x1*omega).
  // x0, x1 lazy in, lazy out.  The Montgomery correction is folded into the add/sub repair:
  // with u = (h1 - h2) mod 2^64 and br its borrow, x1*omega = u - br*2^64, hence
  //   x0 + x1*omega = s + (carry(s)  - br) * 2^64,  s = (x0 + u) mod 2^64
  //   x0 - x1*omega = d + (br - borrow(d)) * 2^64,  d = (x0 - u) mod 2^64
  // and 2^64 == C (mod P).  Range: both true values lie in (-P, 2^64 + P), so the repaired value
  // stays inside [0, 2^64) and the final 64-bit add cannot wrap.
  __device__ __forceinline__ void ct_butterfly(u64& x0, u64& x1, u64 w, u64 wp) const {
    u64 u, s, d;
    u32 m, d0, d1;
    mont_diff(x1, w, wp, u, m);         // m  = -br
    add_carry_plus(x0, u, m, s, d0);    // d0 = carry - br
    sub_borrow_minus(x0, u, m, d, d1);  // d1 = br - borrow
    x0 = fix(s, d0);
    x1 = fix(d, d1);
  }
  __device__ __forceinline__ void ct_butterfly(u64& x0, u64& x1, Tw t) const {
    ct_butterfly(x0, x1, t.w, t.wp);
  }

  // Butterfly with omega = 1: (x0, x1) <- (x0 + x1, x0 - x1).  x0 lazy, x1 MUST be canonical.
  __device__ __forceinline__ void ct_butterfly_one(u64& x0, u64& x1) const {
    u64 s, d;
    u32 d0, d1;
    add_carry_plus(x0, x1, 0u, s, d0);    // d0 = carry
    sub_borrow_minus(x0, x1, 0u, d, d1);  // d1 = -borrow
    x0 = fix(s, d0);
    x1 = fix(d, d1);
  }
};

template <u64 P_>
using Field = FieldOps<StaticModulus<P_>>;
typedef FieldOps<RuntimeModulus> FieldRT;

// ------------------------------------------------------------------------------------------------
// Shoup ("fixed point") arithmetic for runtime moduli below 2^62 - the device twin of the reference's alternative
// modmul type sventt::FixedPoint64SVE (include/sventt/modmul/sve/fixed-point-64.hpp:13-69):
//   multiply(a, b, bp) : q = hi64(a * bp), c = a * b - q * N (low 64 bits), with bp = floor(b * 2^64 / N)
// c lies in [0, 2N) for ANY 64-bit a.  With 4N < 2^64 the butterflies keep every intermediate in [0, 4N) (Harvey's
// lazy butterflies: one conditional subtraction of 2N on the way in, none on the way out) - no carries to repair, and
// the product needs 6 32x32->64 multiplies (4 for hi64(a * bp), 1 each for the two low products) instead of the 10 of
// the lazy Montgomery butterfly above.  Twiddles are stored as (omega, floor(omega * 2^64 / N)) in the Tw slots.
// The element-wise PAdic64 entry points of the C ABI (to_montgomery, multiply_normalize, the fused point-wise
// product) keep their Montgomery meaning: those overloads are inherited from FieldRT.
struct FieldShoup : FieldOps<RuntimeModulus> {
  typedef FieldOps<RuntimeModulus> M;
  using M::mont;  // mont(a, w, wp): Montgomery semantics (element-wise helpers)

  // v in [0, 2m) -> [0, m)
  __device__ __forceinline__ static u64 csub(u64 v, u64 m) {
    const u64 t = v - m;
    return v < m ? v : t;
  }
  // a * omega mod N, lazy: result in [0, 2N); a is any 64-bit value
  __device__ __forceinline__ u64 mul_lazy(u64 a, u64 w, u64 wp) const { return a * w - mulhi64(a, wp) * this->p(); }
  // canonical product with a table twiddle
  __device__ __forceinline__ u64 mont(u64 a, Tw t) const { return csub(mul_lazy(a, t.w, t.wp), this->p()); }
  // (omega, floor(omega * 2^64 / N)) from the Montgomery form wm = omega * 2^64 mod N: omega = wm * 2^-64, and the
  // quotient Q of omega * 2^64 = Q * N + wm is exact, hence Q = -wm * N^-1 mod 2^64
  __device__ __forceinline__ Tw make_tw(u64 wm) const {
    Tw t;
    t.w = M::mont(wm, 1ull, this->pinv());
    t.wp = (0 - wm) * this->pinv();
    return t;
  }
  // [0, 4N) -> [0, N)
  __device__ __forceinline__ u64 canon(u64 v) const {
    const u64 P = this->p();
    return csub(csub(v, 2 * P), P);
  }
  __device__ __forceinline__ u64 canon_if(u64 v, bool on) const { return on ? canon(v) : v; }
  // (x0, x1) <- (x0 + x1 * omega, x0 - x1 * omega), all values in [0, 4N)
  __device__ __forceinline__ void ct_butterfly(u64& x0, u64& x1, u64 w, u64 wp) const {
    const u64 P2 = 2 * this->p();
    const u64 a = csub(x0, P2), t = mul_lazy(x1, w, wp);
    x0 = a + t;
    x1 = a - t + P2;
  }
  __device__ __forceinline__ void ct_butterfly(u64& x0, u64& x1, Tw t) const { ct_butterfly(x0, x1, t.w, t.wp); }
  // omega = 1; x1 canonical (the networks canonicalise it first)
  __device__ __forceinline__ void ct_butterfly_one(u64& x0, u64& x1) const {
    const u64 P2 = 2 * this->p();
    const u64 a = csub(x0, P2), t = x1;
    x0 = a + t;
    x1 = a - t + P2;
  }
};

// The production prime of the reference README (README.md:19), C = 0x3917fffffff.
typedef Field<kP0> F0;
typedef Field<kPGold> FGold;

template <class F>
__host__ __device__ __forceinline__ F make_field(const FieldConsts& k) {
  if constexpr (F::kStatic) {
    return F{};
  } else {
    F f;
    f.k = k;
    return f;
  }
}

}  // namespace xntt
