// SPDX-License-Identifier: Apache-2.0
// Host-side field arithmetic with the semantics of sventt::Modulus
// (include/sventt/modulus.hpp:78-132): constants and roots only, never the data path.
#pragma once
#include <initializer_list>

#include "params.h"

namespace xntt {

typedef unsigned __int128 u128;

inline u64 h_mul(u64 a, u64 b, u64 p) { return (u64)((u128)a * b % p); }
inline u64 h_pow(u64 a, u64 e, u64 p) {
  u64 r = 1 % p;
  for (; e; e >>= 1) {
    if (e & 1) r = h_mul(r, a, p);
    a = h_mul(a, a, p);
  }
  return r;
}
inline u64 h_inv(u64 a, u64 p) { return h_pow(a, p - 2, p); }
inline u64 h_to_mont(u64 a, u64 p) { return (u64)(((u128)a << 64) % p); }
inline u64 h_montgomery_inverse(u64 p) {
  u64 x = p;
  for (int i = 0; i < 6; ++i) x *= 2 - p * x;
  return x;
}

inline bool h_is_prime(u64 n) {
  if (n < 2) return false;
  for (u64 q : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
    if (n % q == 0) return n == q;
  }
  u64 d = n - 1;
  int s = 0;
  while ((d & 1) == 0) d >>= 1, ++s;
  for (u64 a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
    u64 x = h_pow(a % n, d, n);
    if (x == 1 || x == n - 1) continue;
    bool comp = true;
    for (int i = 1; i < s && comp; ++i) {
      x = h_mul(x, x, n);
      if (x == n - 1) comp = false;
    }
    if (comp) return false;
  }
  return true;
}

}  // namespace xntt
