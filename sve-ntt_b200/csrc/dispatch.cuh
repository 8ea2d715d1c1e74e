// SPDX-License-Identifier: Apache-2.0
// Launch-side view of the pass kernels: the dispatchers (one translation unit per direction / mode
// / field flavour so the heavy template instantiations compile in parallel).
#pragma once
#include <atomic>
#include <mutex>

#include "pass_kernel.cuh"

namespace xntt {

// static-modulus kernels (p = kP0) and runtime-modulus kernels (any odd prime < 2^64)
cudaError_t launch_fwd_row(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_row(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_fwd_col(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_col(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
// kernels with generalised src/dst address maps (passes next to the all-to-all of a sharded plan): production modulus
cudaError_t launch_fwd_row_map(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_row_map(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_fwd_col_map(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_col_map(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
// ... and for runtime moduli (Montgomery / Shoup)
cudaError_t launch_fwd_row_map_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_row_map_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_fwd_col_map_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_col_map_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_fwd_row_map_sh(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_row_map_sh(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_fwd_col_map_sh(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_col_map_sh(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_fwd_row_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_row_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_fwd_col_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_col_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
// narrow tiles (params.h: pass_logw(logn, true)): production modulus and runtime Montgomery, plain addressing
cudaError_t launch_fwd_row_narrow(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_row_narrow(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_fwd_col_narrow(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_col_narrow(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_fwd_row_narrow_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_row_narrow_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_fwd_col_narrow_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_col_narrow_rt(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
// Goldilocks with the modulus baked in (plain addressing only; sharded plans take the runtime-modulus kernels)
cudaError_t launch_fwd_row_gold(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_row_gold(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_fwd_col_gold(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_col_gold(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
// runtime moduli below 2^62 with Shoup / FixedPoint64 arithmetic (field.cuh: FieldShoup)
cudaError_t launch_fwd_row_sh(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_row_sh(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_fwd_col_sh(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);
cudaError_t launch_inv_col_sh(int logn, const PassParams& prm, unsigned grid, cudaStream_t st);

#if XNTT_TMA_ROWS
// Measured variant (XNTT_TMA_ROWS): tensor map over the source of a 2^13 row pass viewed as rows of 16 residues
// (128 bytes, 128-byte swizzle), box = 16 x 256.  The encoder comes from the driver through the runtime
// (cudaGetDriverEntryPoint), so the library still does not link libcuda.
inline cudaError_t make_row_tensor_map(CUtensorMap* map, const u64* base, u64 rows) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess) return e;
    if (q != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
    encode = (encode_fn)fn;
  }
  const cuuint64_t dims[2] = {16, rows * 512};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {16, 256}, estr[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void*)base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <class F, bool INV, int TWIST>
cudaError_t launch_row_tma(const PassParams& prm, unsigned grid, cudaStream_t st) {
  auto kern = row_kernel_tma<F, INV, TWIST>;
  constexpr int kSmem = 65536 + 1024 + 16;  // tile, alignment slack, mbarrier
  static std::atomic<bool> attr_done_on[64];
  static std::mutex attr_mu;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (!attr_done_on[dev & 63].load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lock(attr_mu);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    attr_done_on[dev & 63].store(true, std::memory_order_release);
  }
  CUtensorMap map;
  e = make_row_tensor_map(&map, prm.src, prm.rows);
  if (e != cudaSuccess) return e;
  kern<<<grid, kThreads, kSmem, st>>>(prm, map);
  return cudaGetLastError();
}
#endif

inline unsigned sm_count(int dev) {
  static std::atomic<unsigned> cached[64];
  unsigned n = cached[dev & 63].load(std::memory_order_relaxed);
  if (n == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n = (unsigned)v;
    cached[dev & 63].store(n, std::memory_order_relaxed);
  }
  return n;
}
// XNTT_PDL=0 launches the passes fully serialised (measurement knob)
#ifndef XNTT_PDL
#define XNTT_PDL 1
#endif
template <class F, int LOGN, bool COL, bool INV, bool MAP, int TWIST, bool NARROW = false>
cudaError_t launch_kernel(const PassParams& prm, unsigned grid, cudaStream_t st) {
#if XNTT_TMA_ROWS
  if constexpr (!COL && !MAP && LOGN == 13 && F::kStatic) return launch_row_tma<F, INV, TWIST>(prm, grid, st);
#endif
  static_assert(!NARROW || (!MAP && has_narrow_tile(LOGN, COL)), "no narrow tile for this pass");
  constexpr int LOGW = pass_logw(LOGN, NARROW), C = pass_c(LOGN, NARROW);
  typedef PassCfg<LOGN, LOGW, C, COL> Cfg;
  auto kern = pass_kernel<F, LOGN, LOGW, C, COL, INV, TWIST, MAP>;
  // function attributes live in the context of the device they were set on: once per kernel and device, and
  // safe against two host threads launching the same kernel for the first time (the reference's compute_* are
  // const and re-entrant, wrapper.hpp:50-82)
  static std::atomic<bool> attr_done_on[64];
  static std::mutex attr_mu;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::atomic<bool>& attr_done = attr_done_on[dev & 63];
  if (!attr_done.load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lock(attr_mu);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
#ifndef XNTT_CARVE_TILES
#define XNTT_CARVE_TILES pass_minb(COL, C, TWIST == kNoTwist)
#endif
    if constexpr (XNTT_CARVE_TILES > 2) {
      // room for that many resident tiles and not more: what is left of the 228 KiB stays L1, which the twiddle
      // tables live in (measured: 2^24 forward 433 us with the default split, 506 us with L1 squeezed to 32 KiB)
      e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                               (int)((XNTT_CARVE_TILES * (Cfg::kSmemBytes + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)));
      if (e != cudaSuccess) return e;
    }
    attr_done.store(true, std::memory_order_release);
  }
  // programmatic stream serialisation: this grid may become resident while the kernel before it in the stream is
  // still running and waits for it in griddepcontrol.wait (pass_kernel.cuh) - the launch latency between the
  // dependent passes of a plan disappears behind the pass before
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  // ... except for grids of whole tiles with fewer CTAs than SMs: placed while the kernel before them still occupies
  // its SMs, their CTAs end up two to an SM on the idle ones instead of one per SM (measured: one 2^19 transform on
  // whole tiles, 64 CTAs per pass, 31 -> 46 us; 2^21, 256 CTAs, gains: 55.8 -> 52.0 us); narrow tiles do not care
  attr[0].val.programmaticStreamSerializationAllowed = (XNTT_PDL && (NARROW || grid >= sm_count(dev))) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, prm);
}

// one instantiation per kind a pass can take (pass_kernel.cuh: pass_kind).  The generalised-addressing kernels of
// sharded plans serve the pass next to the exchange: a column pass there has the compact form, plus what the planner
// uses for an inner column pass - the whole matrix in the inverse, none at all in the forward direction (the row pass
// applies it) - and a row pass is always plain.
template <class F, int LOGN, bool COL, bool INV, bool MAP = false, bool NARROW = false>
cudaError_t launch_one(const PassParams& prm, unsigned grid, cudaStream_t st) {
  const int kind = pass_kind(COL, INV, MAP, prm);
  if constexpr (COL) {
    if (kind == kColPre) {
      // production modulus, forward only (the planner asks for it nowhere else)
      if constexpr (!INV && F::kStatic) {
        if constexpr (F::P == kP0) return launch_kernel<F, LOGN, COL, INV, MAP, kColPre, NARROW>(prm, grid, st);
      }
      return cudaErrorInvalidValue;
    }
    if (kind == kCompactTwist) return launch_kernel<F, LOGN, COL, INV, MAP, kCompactTwist, NARROW>(prm, grid, st);
    if constexpr (INV || !MAP) {
      if (kind == kFullTwist) return launch_kernel<F, LOGN, COL, INV, MAP, kFullTwist, NARROW>(prm, grid, st);
    }
    // twist-free: the row pass next to it applies the matrix
    if (kind == kNoTwist) return launch_kernel<F, LOGN, COL, INV, MAP, kNoTwist, NARROW>(prm, grid, st);
    return cudaErrorInvalidValue;
  } else {
    if constexpr (!INV && !MAP) {
      if (kind == kPointwise) return launch_kernel<F, LOGN, COL, INV, MAP, kPointwise, NARROW>(prm, grid, st);
      if (kind == kPreTwist) return launch_kernel<F, LOGN, COL, INV, MAP, kPreTwist, NARROW>(prm, grid, st);
      if (kind == kPrePointwise) return launch_kernel<F, LOGN, COL, INV, MAP, kPrePointwise, NARROW>(prm, grid, st);
    }
    if constexpr (INV && !MAP) {
      if (kind == kPostTwist) return launch_kernel<F, LOGN, COL, INV, MAP, kPostTwist, NARROW>(prm, grid, st);
    }
    return launch_kernel<F, LOGN, COL, INV, MAP, kNoTwist, NARROW>(prm, grid, st);
  }
}

#define XNTT_CASE(F, L, COL, INV) \
  case L:                         \
    return launch_one<F, L, COL, INV>(prm, grid, st);
#define XNTT_CASE_MAP(F, L, COL, INV) \
  case L:                             \
    return launch_one<F, L, COL, INV, true>(prm, grid, st);
#define XNTT_CASE_NARROW(F, L, COL, INV) \
  case L:                                \
    return launch_one<F, L, COL, INV, false, true>(prm, grid, st);

}  // namespace xntt
