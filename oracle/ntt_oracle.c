/* SPDX-License-Identifier: Apache-2.0
 *
 * TEST INFRASTRUCTURE - the CPU oracle for the 64-bit NTT hot path.
 *
 * A plain-C restatement of the algorithm the reference uses as ITS oracle and of the PAdic64
 * arithmetic its kernels are built from.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the checker - the
 * product (sve-ntt_b200/) never links, imports or executes anything in this directory.
 *
 * Parity status: PINNED.  The reference stores no golden NTT vectors; what pins this path are
 * (i) the algebraic spot values and round trip asserted by tests/test-ntt-reference.cpp:45-85,
 * (ii) the root sums of tests/test-modulus.cpp:22-46 and (iii) equality with NTTReference itself,
 * which compiles here: oracle/_ref/libnttref.so is built from /root/reference/tests/ntt-reference.hpp
 * (see refdriver.cpp, Makefile) and tests/test_oracle.py checks this file against it word for word;
 * tests/golden/ holds vectors generated from that library so the check also travels to the GPU box.
 *
 * Each function cites the reference lines it follows.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* x*y mod N  (tests/ntt-reference.hpp:16-18; sventt::Modulus::multiply, include/sventt/modulus.hpp:90-93) */
uint64_t oracle_mulmod(uint64_t x, uint64_t y, uint64_t N) { return (uint64_t)((u128)x * y % N); }

/* x^e mod N by square and multiply  (ntt-reference.hpp:20-29; modulus.hpp:100-109) */
uint64_t oracle_powmod(uint64_t x, uint64_t e, uint64_t N) {
  uint64_t acc = 1 % N;
  while (e) {
    if (e & 1) acc = oracle_mulmod(acc, x, N);
    x = oracle_mulmod(x, x, N);
    e >>= 1;
  }
  return acc;
}

/* canonical add / subtract  (ntt-reference.hpp:52-54; modulus.hpp:78-88) */
static inline uint64_t addmod(uint64_t a, uint64_t b, uint64_t N) { return a < N - b ? a + b : a + b - N; }
static inline uint64_t submod(uint64_t a, uint64_t b, uint64_t N) { return a >= b ? a - b : a - b + N; }
uint64_t oracle_addmod(uint64_t a, uint64_t b, uint64_t N) { return addmod(a % N, b % N, N); }
uint64_t oracle_submod(uint64_t a, uint64_t b, uint64_t N) { return submod(a % N, b % N, N); }

/* sventt::Modulus::get_root_forward / get_root_inverse  (modulus.hpp:115-132).
 * Returns 0 (not a root of anything) when order does not divide N-1 - the reference throws
 * std::invalid_argument there. */
uint64_t oracle_root_forward(uint64_t N, uint64_t g, uint64_t order) {
  if (order == 0 || (N - 1) % order != 0) return 0;
  return oracle_powmod(g, (N - 1) / order, N);
}
uint64_t oracle_root_inverse(uint64_t N, uint64_t g, uint64_t order) {
  if (order == 0 || (N - 1) % order != 0) return 0;
  /* g^(((N-1)/order) * (N-2) mod (N-1)) */
  uint64_t e = (uint64_t)((u128)((N - 1) / order) * (N - 2) % (N - 1));
  return oracle_powmod(g, e, N);
}

/* N^-1 mod 2^64  (modulus.hpp:36-68) */
uint64_t oracle_montgomery_inverse(uint64_t N) {
  uint64_t x = (N * 3) ^ 2;
  for (int i = 0; i < 5; ++i) x *= 2 - N * x;
  return x;
}

/* PAdic64: to_montgomery / from_montgomery / precompute / multiply_normalize
 * (include/sventt/modmul/sve/p-adic-64.hpp:19-38, 64-74, 101-115), scalar and canonical. */
uint64_t oracle_to_montgomery(uint64_t b, uint64_t N) { return oracle_mulmod(b % N, (uint64_t)(0 - N) % N, N); }
uint64_t oracle_from_montgomery(uint64_t b, uint64_t N) {
  uint64_t r = (uint64_t)(0 - N) % N; /* 2^64 mod N */
  return oracle_mulmod(b % N, oracle_powmod(r, N - 2, N), N);
}
uint64_t oracle_precompute(uint64_t b, uint64_t N) { return b * oracle_montgomery_inverse(N); }
uint64_t oracle_multiply_normalize(uint64_t a, uint64_t b, uint64_t bp, uint64_t N) {
  uint64_t q = a * bp;
  uint64_t ab1 = (uint64_t)(((u128)a * b) >> 64);
  uint64_t qn1 = (uint64_t)(((u128)q * N) >> 64);
  uint64_t c = ab1 - qn1;
  if (ab1 < qn1) c += N;
  return c;
}

/* Forward transform: natural order in, bit-reversed order out  (ntt-reference.hpp:43-61).
 * Gentleman-Sande sweeps with half-length l = m/2, m/4, ..., 1; the twiddle of position j inside
 * a block is omega_{2l}^j. */
void oracle_ntt_forward(uint64_t* dst, const uint64_t* src, uint64_t m, uint64_t N, uint64_t g) {
  if (m == 0) return;
  int log2m = 0;
  while ((1ull << log2m) < m) ++log2m;
  uint64_t w_len = oracle_powmod(g, (N - 1) >> log2m, N); /* omega_m */
  if (dst != src) memcpy(dst, src, m * sizeof(uint64_t));
  for (uint64_t l = m >> 1; l >= 1; l >>= 1) {
    uint64_t w = 1;
    for (uint64_t j = 0; j < l; ++j) {
      for (uint64_t k = j; k < m; k += 2 * l) {
        uint64_t a = dst[k], b = dst[k + l];
        dst[k] = addmod(a, b, N);
        dst[k + l] = oracle_mulmod(submod(a, b, N), w, N);
      }
      w = oracle_mulmod(w, w_len, N);
    }
    w_len = oracle_mulmod(w_len, w_len, N);
  }
}

/* Inverse transform: bit-reversed in, natural out, scaled by 1/m  (ntt-reference.hpp:63-83). */
void oracle_ntt_inverse(uint64_t* dst, const uint64_t* src, uint64_t m, uint64_t N, uint64_t g) {
  if (m == 0) return;
  int log2m = 0;
  while ((1ull << log2m) < m) ++log2m;
  const uint64_t w_m = oracle_powmod(g, (N - 1) >> log2m, N);
  const uint64_t winv_m = oracle_powmod(w_m, N - 2, N);
  const uint64_t m_inv = oracle_powmod(m % N, N - 2, N);
  for (uint64_t i = 0; i < m; ++i) dst[i] = oracle_mulmod(src[i], m_inv, N);
  for (int s = 0; s < log2m; ++s) {
    const uint64_t l = 1ull << s;
    const uint64_t step = oracle_powmod(winv_m, 1ull << (log2m - s - 1), N); /* omega_{2l}^-1 */
    uint64_t w = 1;
    for (uint64_t j = 0; j < l; ++j) {
      for (uint64_t k = j; k < m; k += 2 * l) {
        uint64_t a = dst[k], b = oracle_mulmod(dst[k + l], w, N);
        dst[k] = addmod(a, b, N);
        dst[k + l] = submod(a, b, N);
      }
      w = oracle_mulmod(w, step, N);
    }
  }
}

/* Point-wise product between the transforms of a polynomial multiply
 * (examples/magic-series/gaussian-polynomial.hpp:201-212): c[j] = c[j] * b[j] mod N. */
void oracle_pointwise_mul(uint64_t* dst, const uint64_t* a, const uint64_t* b, uint64_t count, uint64_t N) {
  for (uint64_t i = 0; i < count; ++i) dst[i] = oracle_mulmod(a[i], b[i], N);
}

/* Direct evaluation of one output word of the forward transform (size-independent spot check):
 * dst[pos] = sum_i src[i] * omega_m^(i * bitrev(pos))  (the definition behind
 * tests/test-ntt-reference.cpp:45-80). */
uint64_t oracle_dft_point(const uint64_t* src, uint64_t m, uint64_t N, uint64_t g, uint64_t pos) {
  int log2m = 0;
  while ((1ull << log2m) < m) ++log2m;
  uint64_t k = 0;
  for (int b = 0; b < log2m; ++b) k |= ((pos >> b) & 1) << (log2m - 1 - b);
  const uint64_t w = oracle_powmod(oracle_powmod(g, (N - 1) >> log2m, N), k, N);
  uint64_t acc = 0, wi = 1;
  for (uint64_t i = 0; i < m; ++i) {
    acc = addmod(acc, oracle_mulmod(src[i] % N, wi, N), N);
    wi = oracle_mulmod(wi, w, N);
  }
  return acc;
}

/* The synthetic input stream of SURVEY.md section 8(c)/(d): xorshift64, reduced mod N. */
void oracle_fill_xorshift(uint64_t* dst, uint64_t count, uint64_t seed, uint64_t N) {
  uint64_t s = seed;
  for (uint64_t i = 0; i < count; ++i) {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    dst[i] = s % N;
  }
}

/* FNV-1a over 64-bit words (SURVEY.md section 8(c) fingerprint). */
uint64_t oracle_fnv64(const uint64_t* v, uint64_t count) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (uint64_t i = 0; i < count; ++i) h = (h ^ v[i]) * 0x100000001b3ull;
  return h;
}

/* ---- Kinnaes' formula for the number of magic series (examples/magic-series-kinnaes/kinnaes.hpp) ----
 *
 * x / y mod N  (sventt::Modulus::divide, include/sventt/modulus.hpp; N prime: Fermat inverse) */
static uint64_t divmod(uint64_t x, uint64_t y, uint64_t N) { return oracle_mulmod(x % N, oracle_powmod(y % N, N - 2, N), N); }

/* MagicSeriesKinnaes::compute_comb(a, b) = binomial(a, b) mod N  (kinnaes.hpp:36-47) */
uint64_t oracle_kinnaes_comb(uint64_t a, uint64_t b, uint64_t N) {
  uint64_t num = a % N, den = b % N;
  for (uint64_t i = 1; i < b; ++i) num = oracle_mulmod(num, (a - i) % N, N);
  for (uint64_t i = 2; i < b; ++i) den = oracle_mulmod(den, i % N, N);
  return divmod(num, den, N);
}

/* MagicSeriesKinnaes::compute_sum(j_begin, j_end)  (kinnaes.hpp:51-157), lane by lane:
 *   sum over J = j_begin+1 .. j_end of  prod_{l<m} (w^(J (m^2-m+1+l)) - 1)  /  ( w^(J r) prod_{l<m} (w^(J (l+1)) - 1) )
 * with w a primitive n-th root of unity and r = m (m-1)/2 * m; accumulated as one fraction like the
 * reference does (kinnaes.hpp:126-133) and divided at the end (:156).  Returns 0 when n does not divide N-1. */
uint64_t oracle_kinnaes_sum(uint64_t N, uint64_t g, uint64_t m, uint64_t n, uint64_t j_begin, uint64_t j_end) {
  const uint64_t w = oracle_root_forward(N, g, n);
  if (w == 0) return 0;
  const uint64_t r = m * (m - 1) / 2 * m;
  uint64_t num_sum = 0, den_sum = 1;
  for (uint64_t J = j_begin + 1; J <= j_end; ++J) {
    const uint64_t wj = oracle_powmod(w, J, N);
    uint64_t num_term = oracle_powmod(wj, m * m - m + 1, N), den_term = wj;
    uint64_t num_prod = 1, den_prod = oracle_powmod(wj, r, N);
    for (uint64_t l = 0; l < m; ++l) {
      num_prod = oracle_mulmod(submod(num_term, 1, N), num_prod, N);
      den_prod = oracle_mulmod(submod(den_term, 1, N), den_prod, N);
      num_term = oracle_mulmod(num_term, wj, N);
      den_term = oracle_mulmod(den_term, wj, N);
    }
    num_sum = addmod(oracle_mulmod(den_sum, num_prod, N), oracle_mulmod(num_sum, den_prod, N), N);
    den_sum = oracle_mulmod(den_sum, den_prod, N);
  }
  return divmod(num_sum, den_sum, N);
}

/* MagicSeriesKinnaes::compute()  (kinnaes.hpp:27-34): (2 * compute_sum(0, n/2) + binomial(m^2, m)) / n */
uint64_t oracle_kinnaes_compute(uint64_t N, uint64_t g, uint64_t m, uint64_t n) {
  uint64_t sum = oracle_kinnaes_sum(N, g, m, n, 0, n / 2);
  sum = addmod(sum, sum, N);
  sum = addmod(sum, oracle_kinnaes_comb(m * m, m, N), N);
  return divmod(sum, n, N);
}
