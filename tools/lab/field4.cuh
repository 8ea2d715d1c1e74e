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
  f.mont_parts(x1, w, wp, h1, h2);
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
  f.mont_parts(x1, w, wp, h1, h2);
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
