// SPDX-License-Identifier: Apache-2.0
// Development lab (not part of the product): register-resident butterfly loops for candidate PAdic64
// butterfly formulations.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I../../sve-ntt_b200/csrc
// bfly_lab.cu -o bfly_lab ; static pipe estimate: cuobjdump -sass + sasscount.py.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "field.cuh"
#include "field1.cuh"
#include "field2.cuh"
#include "field3.cuh"
#include "field4.cuh"

using namespace xntt;

template <int V>
__device__ __forceinline__ void bfly(u64& x0, u64& x1, u64 w, u64 wp) {
  if constexpr (V == 0) {
    const F0 f{};
    f.ct_butterfly(x0, x1, w, wp);
  } else if constexpr (V == 1) {
    lab::bf_v1(x0, x1, w, wp);
  } else if constexpr (V == 2) {
    lab::bf_v2(x0, x1, w, wp);
  } else if constexpr (V == 3) {
    lab::bf_v3(x0, x1, w, wp);
  } else if constexpr (V == 4) {
    lab::bf_v4(x0, x1, w, wp);
  } else if constexpr (V == 5) {
    lab::bf_v5(x0, x1, w, wp);
  } else if constexpr (V == 6) {
    lab::bf_v6(x0, x1, w, wp);
  } else if constexpr (V == 7) {
    lab::bf_v7(x0, x1, w, wp);
  } else if constexpr (V == 8) {
    lab::bf_v8(x0, x1, w, wp);
  } else if constexpr (V == 9) {
    lab::bf_mixed<0>(x0, x1, w, wp);
  } else if constexpr (V == 10) {
    lab::bf_mixed<1>(x0, x1, w, wp);
  } else if constexpr (V == 11) {
    lab::bf_mixed<3>(x0, x1, w, wp);
  } else if constexpr (V == 12) {
    lab::bf_v12(x0, x1, w, wp);
  } else if constexpr (V == 13) {
    lab::bf_v13(x0, x1, w, wp);
  } else if constexpr (V == 15) {
    lab::bf_fp64<0, 0, false>(x0, x1, w, wp);
  } else if constexpr (V == 16) {
    lab::bf_fp64<0, 1, false>(x0, x1, w, wp);
  } else if constexpr (V == 17) {
    lab::bf_fp64<0, 3, false>(x0, x1, w, wp);
  } else if constexpr (V == 18) {
    lab::bf_fp64<0, 0, true>(x0, x1, w, wp);
  } else if constexpr (V == 19) {
    lab::bf_fp64<1, 0, false>(x0, x1, w, wp);
  } else if constexpr (V == 20) {
    lab::bf_fp64<2, 0, false>(x0, x1, w, wp);
  } else if constexpr (V == 21) {
    lab::bf_fp64<0, 1, true>(x0, x1, w, wp);
  } else if constexpr (V == 22) {
    lab::bf_fp64<1, 1, false>(x0, x1, w, wp);
  } else if constexpr (V == 23) {
    lab::bf_v23(x0, x1, w, wp);
  } else if constexpr (V == 24) {
    lab::bf_v24(x0, x1, w, wp);
  } else if constexpr (V == 25) {
    lab::bf_v25<0>(x0, x1, w, wp);
  } else if constexpr (V == 26) {
    lab::bf_v25<1>(x0, x1, w, wp);
  } else if constexpr (V == 27) {
    lab::bf_v27(x0, x1, w, wp);
  } else if constexpr (V == 28) {
    lab::bf_v28(x0, x1, w, wp);
  } else if constexpr (V == 29) {
    lab::bf_v29(x0, x1, w, wp);
  }
}

// 16 values per thread = 8 independent butterflies per level, 4 levels of a radix-16 network per iteration
// (indices permuted between levels so that values really mix, as in the kernel)
template <int V, int MINB>
__global__ void __launch_bounds__(256, MINB) loop_kernel(u64* out, const u64* in, const Tw* tw, int iters) {
  const u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
  u64 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = in[(tid * 16 + i) & 0xffff];
  Tw t[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = tw[(tid + i) & 255];
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int lam = 0; lam < 4; ++lam) {
      const int h = 8 >> lam;
#pragma unroll
      for (int g = 0; g < (1 << lam); ++g)
#pragma unroll
        for (int r = 0; r < h; ++r) bfly<V>(x[g * 2 * h + r], x[g * 2 * h + r + h], t[lam].w, t[lam].wp);
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) out[(size_t)tid * 16 + i] = x[i];
}

// v14: deferred repairs.  A value that is the ADDEND (x0) of its next butterfly keeps its pending carry count e (true
// value = x + e * 2^64) instead of being repaired; the count enters the next carry word for free (third operand of the
// addc / subc that materialises it) and the repair by delta * C happens once, with |delta| <= 2.  Only values about to
// be multiplied (x1) and the outputs of the last level are repaired: 16 instead of 24 repairs per radix-8 network,
// 40 instead of 64 in this radix-16 loop.
__device__ __forceinline__ void bfly_deferred(u64& x0, u32& e0, u64& x1, u32& e1, u64 w, u64 wp) {
  const F0 f{};
  u64 h1, h2, u, s, d;
  u32 m;
  lab::mont_parts(f, x1, w, wp, h1, h2);  // x1 arrives repaired (e1 == 0)
  sub_borrow_mask(h1, h2, u, m);
  u32 al, ah, bl, bh, sl, sh, dl, dh, ks, kd;
  unpack64(x0, al, ah);
  unpack64(u, bl, bh);
  asm("add.cc.u32 %0, %3, %5;\n\taddc.cc.u32 %1, %4, %6;\n\taddc.u32 %2, %7, %8;"
      : "=r"(sl), "=r"(sh), "=r"(ks)
      : "r"(al), "r"(ah), "r"(bl), "r"(bh), "r"(m), "r"(e0));
  asm("sub.cc.u32 %0, %3, %5;\n\tsubc.cc.u32 %1, %4, %6;\n\tsubc.u32 %2, %8, %7;"
      : "=r"(dl), "=r"(dh), "=r"(kd)
      : "r"(al), "r"(ah), "r"(bl), "r"(bh), "r"(m), "r"(e0));
  s = pack64(sl, sh);
  d = pack64(dl, dh);
  x0 = s, e0 = ks;
  x1 = d, e1 = kd;
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) loop_kernel_deferred(u64* out, const u64* in, const Tw* tw, int iters) {
  const F0 f{};
  const u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
  u64 x[16];
  u32 e[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = in[(tid * 16 + i) & 0xffff], e[i] = 0;
  Tw t[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = tw[(tid + i) & 255];
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int lam = 0; lam < 4; ++lam) {
      const int h = 8 >> lam;
#pragma unroll
      for (int g = 0; g < (1 << lam); ++g)
#pragma unroll
        for (int r = 0; r < h; ++r) {
          const int i0 = g * 2 * h + r, i1 = i0 + h;
          bfly_deferred(x[i0], e[i0], x[i1], e[i1], t[lam].w, t[lam].wp);
        }
      // repair what is multiplied next (or leaves the registers)
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const bool next_is_x1 = lam == 3 || ((q >> (2 - lam)) & 1);
        if (next_is_x1) x[q] = f.fix(x[q], e[q]), e[q] = 0;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) out[(size_t)tid * 16 + i] = x[i];
}

static u64 mulmod(u64 a, u64 b, u64 p) { return (u64)((unsigned __int128)a * b % p); }

template <int V, int MINB = 2>
static int run(const char* name, int iters_time, bool check = true) {
  const u64 P = kP0;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned blocks = sms * 8, threads = 256;
  const size_t nthreads = (size_t)blocks * threads;
  std::vector<u64> hin(65536);
  u64 s = 0x9e3779b97f4a7c15ull;
  for (auto& v : hin) {
    s ^= s << 13, s ^= s >> 7, s ^= s << 17;
    v = s;  // lazy inputs: any u64
  }
  // some extreme values
  hin[0] = 0, hin[1] = ~0ull, hin[2] = P, hin[3] = P - 1, hin[4] = P + 1, hin[5] = 1, hin[16] = ~0ull, hin[17] = ~0ull;
  std::vector<Tw> htw(256);
  const u64 pinv = montgomery_inverse(P);
  for (int i = 0; i < 256; ++i) {
    s ^= s << 13, s ^= s >> 7, s ^= s << 17;
    u64 w = s % P;
    if (i == 0) w = P - 1;
    if (i == 1) w = 1;
    if (i == 2) w = 0;
    htw[i].w = w;
    htw[i].wp = w * pinv;
  }
  u64 *din, *dout;
  Tw* dtw;
  cudaMalloc(&din, 65536 * 8);
  cudaMalloc(&dout, nthreads * 16 * 8);
  cudaMalloc(&dtw, 256 * sizeof(Tw));
  cudaMemcpy(din, hin.data(), 65536 * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dtw, htw.data(), 256 * sizeof(Tw), cudaMemcpyHostToDevice);
  // correctness: 2 iterations, compare residues mod P for the first 4096 threads
  if constexpr (V == 14)
    loop_kernel_deferred<MINB><<<blocks, threads>>>(dout, din, dtw, 2);
  else
    loop_kernel<V, MINB><<<blocks, threads>>>(dout, din, dtw, 2);
  cudaDeviceSynchronize();
  std::vector<u64> got(4096 * 16);
  cudaMemcpy(got.data(), dout, got.size() * 8, cudaMemcpyDeviceToHost);
  // 2^-64 mod P
  u64 rinv = 1;
  {
    // R^-1 = (2^64 mod P)^(P-2)
    u64 base = (u64)((((unsigned __int128)1) << 64) % P), e = P - 2;
    while (e) {
      if (e & 1) rinv = mulmod(rinv, base, P);
      base = mulmod(base, base, P);
      e >>= 1;
    }
  }
  size_t bad = 0;
  for (u32 tid = 0; tid < 4096; ++tid) {
    u64 x[16];
    for (int i = 0; i < 16; ++i) x[i] = hin[(tid * 16 + i) & 0xffff] % P;
    for (int it = 0; it < 2; ++it)
      for (int lam = 0; lam < 4; ++lam) {
        const int h = 8 >> lam;
        const u64 om = mulmod(htw[(tid + lam) & 255].w, rinv, P);
        for (int g = 0; g < (1 << lam); ++g)
          for (int r = 0; r < h; ++r) {
            u64 &a = x[g * 2 * h + r], &b = x[g * 2 * h + r + h];
            const u64 u = mulmod(b, om, P);
            const u64 s0 = (u64)(((unsigned __int128)a + u) % P), d0 = (u64)(((unsigned __int128)a + P - u) % P);
            a = s0, b = d0;
          }
      }
    for (int i = 0; i < 16; ++i)
      if (check && got[(size_t)tid * 16 + i] % P != x[i]) ++bad;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    if constexpr (V == 14)
      loop_kernel_deferred<MINB><<<blocks, threads>>>(dout, din, dtw, iters_time);
    else
      loop_kernel<V, MINB><<<blocks, threads>>>(dout, din, dtw, iters_time);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double bf = 32.0 * iters_time * nthreads;
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  const double cyc = best * 1e-3 * clk * 1e3 * sms * 4 / (bf / 32);
  printf("{\"variant\": \"%s\", \"minb\": %d, \"bad\": %zu, \"ms\": %.3f, \"gbfly_s\": %.1f, \"cyc_per_warp_bfly_at_maxclk\": %.1f}\n", name, MINB, bad,
         best, bf / (best * 1e-3) / 1e9, cyc);
  fflush(stdout);
  cudaFree(din), cudaFree(dout), cudaFree(dtw);
  return bad != 0;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 500;
  int rc = 0;
  const int only = argc > 2 ? atoi(argv[2]) : -1;
  if (only < 0 || only == 0) rc |= run<0, 2>("v0_field_cuh_product", iters);  // FieldOps::ct_butterfly as shipped
  if (only < 0 || only == 9) rc |= run<9, 2>("v9_field_cuh_two_sided", iters);  // mont_parts + sub_borrow_mask (until r3)
  if (only < 0 || only == 14) rc |= run<14, 2>("v14_deferred_repairs", iters);
  if (only < 0 || only == 12) rc |= run<12, 2>("v12_lhi_from_qP", iters);
  if (only < 0 || only == 13) rc |= run<13, 2>("v13_lhi_from_qP_q1P0_shifts", iters);
  if (only < 0 || only == 10) rc |= run<10, 2>("v10_fix_alu_sum", iters);
  if (only < 0 || only == 11) rc |= run<11, 2>("v11_fix_alu_both", iters);
  if (only < 0 || only == 15) rc |= run<15, 2>("v15_h2_fp64_denormal", iters);
  if (only < 0 || only == 16) rc |= run<16, 2>("v16_h2_fp64_fix_alu_sum", iters);
  if (only < 0 || only == 17) rc |= run<17, 2>("v17_h2_fp64_fix_alu_both", iters);
  if (only < 0 || only == 18) rc |= run<18, 2>("v18_h2_and_q_fp64", iters);
  if (only < 0 || only == 21) rc |= run<21, 2>("v21_h2_and_q_fp64_fix_alu_sum", iters);
  if (only < 0 || only == 19) rc |= run<19, 2>("v19_h2_fp64_magic", iters);
  if (only < 0 || only == 22) rc |= run<22, 2>("v22_h2_fp64_magic_fix_alu_sum", iters);
  if (only < 0 || only == 20) rc |= run<20, 2>("v20_h2_int_2wide", iters);
  if (only < 0 || only == 23) rc |= run<23, 2>("v23_three_operand_int128", iters);
  if (only < 0 || only == 24) rc |= run<24, 2>("v24_three_operand_chains", iters);
  if (only < 0 || only == 25) rc |= run<25, 2>("v25_q1_chained", iters);
  if (only < 0 || only == 26) rc |= run<26, 2>("v26_q1_chained_prmt_pairs", iters);
  if (only < 0 || only == 27) rc |= run<27, 2>("v27_glue_pinned_to_alu", iters);
  if (only < 0 || only == 28) rc |= run<28, 2>("v28_q1_chained_carry2_as_borrow", iters);
  if (only < 0 || only == 29) rc |= run<29, 2>("v29_v28_carry_words_without_not", iters);
  if (only < 0 || only == 6) rc |= run<6, 2>("v6_probe_nofix", iters, false);
  if (only < 0 || only == 7) rc |= run<7, 2>("v7_probe_mont_only", iters, false);
  return rc;
}
