// SPDX-License-Identifier: Apache-2.0
//
// CUDA implementation of backend.h: memory, streams, the small kernels (twiddle generation,
// element-wise PAdic64 helpers, register micro-benchmarks) and the pass-kernel dispatcher.
#include <cuda_runtime.h>

#include <string>

#include "backend.h"
#include "dispatch.cuh"
#include "kinnaes_kernel.cuh"
#include "misc_kernels.cuh"
#include "transpose_kernel.cuh"

namespace xntt {

static thread_local std::string g_err = "no error";
static int fail(cudaError_t e) {
  g_err = cudaGetErrorString(e);
  return e == cudaErrorMemoryAllocation ? 2 : 1;
}
#define CU(call)                              \
  do {                                        \
    cudaError_t e_ = (call);                  \
    if (e_ != cudaSuccess) return fail(e_);   \
  } while (0)

template <class F>
__global__ void gen_table_kernel(const F f, Tw* out, u32 count, int kind, int logn, int shift,
                                 const __grid_constant__ PowTable t) {
  const u32 idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < count) out[idx] = table_entry<F>(f, idx, kind, logn, shift, t);
}
template <class F>
__global__ void to_mont_kernel(const F f, u64* dst, const u64* src, size_t n, u64 r2) {
  const u64 r2p = f.companion(r2);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = ew_to_mont<F>(f, src[i], r2, r2p);
}
template <class F>
__global__ void from_mont_kernel(const F f, u64* dst, const u64* src, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = ew_from_mont<F>(f, src[i]);
}
template <class F>
__global__ void mulnorm_kernel(const F f, u64* dst, const u64* a, const u64* b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = ew_mulnorm<F>(f, a[i], b[i]);
}

// ------------------------------------------------------------------------------------------------
// Register-resident micro-benchmarks (calibrate the integer roofline on the box).
template <int KIND>
__global__ void __launch_bounds__(256) microbench_kernel(u64* out, int iters, u32 a, u32 b) {
  const u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
  if constexpr (KIND == 0) {
    u32 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = tid + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = x[i] * a + b;
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i];
    if (s == 0x12345u) out[tid] = s;
  } else if constexpr (KIND == 1) {
    u64 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = ((u64)tid << 32) + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = (u64)(u32)x[i] * a + x[i];
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i];
    if (s == 0x12345u) out[tid] = s;
  } else if constexpr (KIND == 2) {
    u32 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = tid + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x[i]) : "r"(a), "r"(b));
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i];
    if (s == 0x12345u) out[tid] = s;
  } else if constexpr (KIND == 3) {
    u64 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = ((u64)tid << 32) + i * 0x9e3779b97f4a7c15ull;
    const F0 f{};
    const u64 w = (((u64)a << 32) | b) % kP0, wp = f.companion(w);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) f.ct_butterfly(x[i], x[i + 1], w, wp);
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i];
    if (s == 0x12345u) out[tid] = s;
  } else if constexpr (KIND == 5) {
    u32 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = tid * 0x9e3779b1u + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __umulhi(x[i], a) + b;
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i];
    if (s == 0x12345u) out[tid] = s;
  } else {
    // IMAD and LOP3 interleaved 1:1 (dual-pipe issue test)
    u32 x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = tid + i, y[i] = tid * 3 + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          x[i] = x[i] * a + b;
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(y[i]) : "r"(a), "r"(b));
        }
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i] ^ y[i];
    if (s == 0x12345u) out[tid] = s;
  }
}


namespace be {

int device_count(int* n) {
  CU(cudaGetDeviceCount(n));
  return 0;
}
int get_device(int* dev) {
  CU(cudaGetDevice(dev));
  return 0;
}
int set_device(int dev) {
  CU(cudaSetDevice(dev));
  return 0;
}
int dev_malloc(void** p, size_t bytes) {
  CU(cudaMalloc(p, bytes ? bytes : 16));
  return 0;
}
int dev_free(void* p) {
  if (p) CU(cudaFree(p));
  return 0;
}
int mem_info(size_t* free_bytes, size_t* total_bytes) {
  CU(cudaMemGetInfo(free_bytes, total_bytes));
  return 0;
}
int host_malloc_pinned(void** p, size_t bytes) {
  CU(cudaMallocHost(p, bytes ? bytes : 16));
  return 0;
}
int host_free_pinned(void* p) {
  if (p) CU(cudaFreeHost(p));
  return 0;
}
int memcpy_h2d(void* dst, const void* src, size_t bytes, void* st) {
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)st));
  return 0;
}
int memcpy_d2h(void* dst, const void* src, size_t bytes, void* st) {
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)st));
  return 0;
}
int memcpy2d_h2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, void* st) {
  CU(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice, (cudaStream_t)st));
  return 0;
}
int memcpy2d_d2h(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, void* st) {
  CU(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost, (cudaStream_t)st));
  return 0;
}
int enable_peer_access(int dev, int peer) {
  if (dev == peer) return 0;
  int can = 0;
  CU(cudaDeviceCanAccessPeer(&can, dev, peer));
  if (!can) {
    g_err = "no peer access between devices " + std::to_string(dev) + " and " + std::to_string(peer);
    return 1;
  }
  int prev = 0;
  CU(cudaGetDevice(&prev));
  CU(cudaSetDevice(dev));
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    e = cudaSuccess;
  }
  cudaSetDevice(prev);
  if (e != cudaSuccess) return fail(e);
  return 0;
}
int stream_sync(void* st) {
  CU(cudaStreamSynchronize((cudaStream_t)st));
  return 0;
}
int stream_create(void** st) {
  cudaStream_t s;
  CU(cudaStreamCreate(&s));  // blocking stream: ordered against the legacy default stream the plain path uses
  *st = s;
  return 0;
}
int stream_destroy(void* st) {
  CU(cudaStreamDestroy((cudaStream_t)st));
  return 0;
}
int event_create(void** ev) {
  cudaEvent_t e;
  CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *ev = e;
  return 0;
}
int event_destroy(void* ev) {
  CU(cudaEventDestroy((cudaEvent_t)ev));
  return 0;
}
int event_record(void* ev, void* st) {
  CU(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)st));
  return 0;
}
int stream_wait_event(void* st, void* ev) {
  CU(cudaStreamWaitEvent((cudaStream_t)st, (cudaEvent_t)ev, 0));
  return 0;
}
int pointer_is_device(const void* p, int* is_device) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();  // unregistered host memory on old drivers: not an error for us
    *is_device = 0;
    return 0;
  }
  *is_device = (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? 1 : 0;
  return 0;
}
int host_device_pointer(const void* p, void** dev) {
  *dev = nullptr;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (a.type == cudaMemoryTypeHost && a.devicePointer != nullptr) *dev = a.devicePointer;
  return 0;
}
const char* last_error() { return g_err.c_str(); }

int launch_pass(int logn, bool col, bool inverse, bool map, const PassParams& prm, unsigned grid, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (prm.narrow) {
    // narrow tiles: plain addressing, production modulus or runtime Montgomery (the planner asks for nothing else)
    if (map || !field_has_narrow(prm.field) || !has_narrow_tile(logn, col)) return fail(cudaErrorInvalidValue);
    if (prm.field.p == kP0) {
      if (col)
        e = inverse ? launch_inv_col_narrow(logn, prm, grid, st) : launch_fwd_col_narrow(logn, prm, grid, st);
      else
        e = inverse ? launch_inv_row_narrow(logn, prm, grid, st) : launch_fwd_row_narrow(logn, prm, grid, st);
    } else {
      if (col)
        e = inverse ? launch_inv_col_narrow_rt(logn, prm, grid, st) : launch_fwd_col_narrow_rt(logn, prm, grid, st);
      else
        e = inverse ? launch_inv_row_narrow_rt(logn, prm, grid, st) : launch_fwd_row_narrow_rt(logn, prm, grid, st);
    }
  } else if (map) {
    if (prm.field.p == kP0) {
      if (col)
        e = inverse ? launch_inv_col_map(logn, prm, grid, st) : launch_fwd_col_map(logn, prm, grid, st);
      else
        e = inverse ? launch_inv_row_map(logn, prm, grid, st) : launch_fwd_row_map(logn, prm, grid, st);
    } else if (prm.field.kind == kFieldShoup) {
      if (col)
        e = inverse ? launch_inv_col_map_sh(logn, prm, grid, st) : launch_fwd_col_map_sh(logn, prm, grid, st);
      else
        e = inverse ? launch_inv_row_map_sh(logn, prm, grid, st) : launch_fwd_row_map_sh(logn, prm, grid, st);
    } else {
      if (col)
        e = inverse ? launch_inv_col_map_rt(logn, prm, grid, st) : launch_fwd_col_map_rt(logn, prm, grid, st);
      else
        e = inverse ? launch_inv_row_map_rt(logn, prm, grid, st) : launch_fwd_row_map_rt(logn, prm, grid, st);
    }
  } else if (prm.field.p == kP0) {
    if (col)
      e = inverse ? launch_inv_col(logn, prm, grid, st) : launch_fwd_col(logn, prm, grid, st);
    else
      e = inverse ? launch_inv_row(logn, prm, grid, st) : launch_fwd_row(logn, prm, grid, st);
  } else if (prm.field.p == kPGold) {
    if (col)
      e = inverse ? launch_inv_col_gold(logn, prm, grid, st) : launch_fwd_col_gold(logn, prm, grid, st);
    else
      e = inverse ? launch_inv_row_gold(logn, prm, grid, st) : launch_fwd_row_gold(logn, prm, grid, st);
  } else if (prm.field.kind == kFieldShoup) {
    if (col)
      e = inverse ? launch_inv_col_sh(logn, prm, grid, st) : launch_fwd_col_sh(logn, prm, grid, st);
    else
      e = inverse ? launch_inv_row_sh(logn, prm, grid, st) : launch_fwd_row_sh(logn, prm, grid, st);
  } else {
    if (col)
      e = inverse ? launch_inv_col_rt(logn, prm, grid, st) : launch_fwd_col_rt(logn, prm, grid, st);
    else
      e = inverse ? launch_inv_row_rt(logn, prm, grid, st) : launch_fwd_row_rt(logn, prm, grid, st);
  }
  if (e != cudaSuccess) return fail(e);
  return 0;
}

// run `call` with the field object matching fc
#define XNTT_WITH_FIELD(fc, call)                \
  do {                                           \
    if ((fc).p == kP0) {                         \
      typedef F0 F;                              \
      const F f = make_field<F>(fc);             \
      call;                                      \
    } else if ((fc).kind == kFieldShoup) {       \
      typedef FieldShoup F;                      \
      const F f = make_field<F>(fc);             \
      call;                                      \
    } else {                                     \
      typedef FieldRT F;                         \
      const F f = make_field<F>(fc);             \
      call;                                      \
    }                                            \
  } while (0)

int launch_gen_table(const FieldConsts& fc, Tw* out, u32 count, int kind, int logn, int shift, const PowTable& t,
                     void* stream) {
  const u32 threads = 128, blocks = (count + threads - 1) / threads;
  XNTT_WITH_FIELD(fc, (gen_table_kernel<F><<<blocks, threads, 0, (cudaStream_t)stream>>>(f, out, count, kind, logn,
                                                                                       shift, t)));
  CU(cudaGetLastError());
  return 0;
}

int launch_kinnaes(const KinnaesParams& prm, unsigned blocks, void* stream) {
  XNTT_WITH_FIELD(prm.field, (kinnaes_kernel<F><<<blocks, kKinnaesThreads, 0, (cudaStream_t)stream>>>(f, prm)));
  CU(cudaGetLastError());
  return 0;
}

static unsigned ew_grid(size_t n) {
  size_t b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  return b ? (unsigned)b : 1u;
}
int launch_to_mont(const FieldConsts& fc, u64* dst, const u64* src, size_t n, u64 r2, void* st) {
  XNTT_WITH_FIELD(fc, (to_mont_kernel<F><<<ew_grid(n), 256, 0, (cudaStream_t)st>>>(f, dst, src, n, r2)));
  CU(cudaGetLastError());
  return 0;
}
int launch_from_mont(const FieldConsts& fc, u64* dst, const u64* src, size_t n, void* st) {
  XNTT_WITH_FIELD(fc, (from_mont_kernel<F><<<ew_grid(n), 256, 0, (cudaStream_t)st>>>(f, dst, src, n)));
  CU(cudaGetLastError());
  return 0;
}
int launch_mulnorm(const FieldConsts& fc, u64* dst, const u64* a, const u64* b, size_t n, void* st) {
  XNTT_WITH_FIELD(fc, (mulnorm_kernel<F><<<ew_grid(n), 256, 0, (cudaStream_t)st>>>(f, dst, a, b, n)));
  CU(cudaGetLastError());
  return 0;
}

int launch_transpose(u64* dst, const u64* src, u64 rows, u64 cols, u64 ld_dst, u64 ld_src, void* stream) {
  TransposeParams p{dst, src, rows, cols, ld_dst, ld_src, (u32)((cols + kTrTile - 1) / kTrTile)};
  const u64 tiles_r = (rows + kTrTile - 1) / kTrTile;
  static std::atomic<bool> attr_done_on[64];  // per device, like the pass kernels (dispatch.cuh)
  static std::mutex attr_mu;
  int dev = 0;
  CU(cudaGetDevice(&dev));
  std::atomic<bool>& attr_done = attr_done_on[dev & 63];
  if (!attr_done.load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lock(attr_mu);
    CU(cudaFuncSetAttribute(transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(kTrSmemWords * sizeof(u64))));
    CU(cudaFuncSetAttribute(transpose_inplace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(2 * kTrSmemWords * sizeof(u64))));
    attr_done.store(true, std::memory_order_release);
  }
  if (dst == src) {
    const u64 pairs = (u64)p.tiles_c * (p.tiles_c + 1) / 2;  // upper triangle of the tile grid
    if (pairs > 0x7fffffffull) {
      g_err = "matrix too large for the in-place transposition";
      return 1;
    }
    transpose_inplace_kernel<<<(unsigned)pairs, kTrThreads, 2 * kTrSmemWords * sizeof(u64), (cudaStream_t)stream>>>(p);
  } else {
    transpose_kernel<<<(unsigned)(tiles_r * p.tiles_c), kTrThreads, kTrSmemWords * sizeof(u64),
                       (cudaStream_t)stream>>>(p);
  }
  CU(cudaGetLastError());
  return 0;
}

int microbench(int kind, int iters, double* gops, double* ms_out) {
  int dev = 0, sms = 0;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned blocks = (unsigned)sms * 8, threads = 256;
  u64* out = nullptr;
  CU(cudaMalloc(&out, (size_t)blocks * threads * sizeof(u64)));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CU(cudaEventRecord(e0, 0));
    switch (kind) {
      case 0:
        microbench_kernel<0><<<blocks, threads>>>(out, iters, 0x9e3779b1u, 0x7f4a7c15u);
        break;
      case 1:
        microbench_kernel<1><<<blocks, threads>>>(out, iters, 0x9e3779b1u, 0x7f4a7c15u);
        break;
      case 2:
        microbench_kernel<2><<<blocks, threads>>>(out, iters, 0x9e3779b1u, 0x7f4a7c15u);
        break;
      case 3:
        microbench_kernel<3><<<blocks, threads>>>(out, iters, 0x9e3779b1u, 0x7f4a7c15u);
        break;
      case 5:
        microbench_kernel<5><<<blocks, threads>>>(out, iters, 0x9e3779b1u, 0x7f4a7c15u);
        break;
      default:
        microbench_kernel<4><<<blocks, threads>>>(out, iters, 0x9e3779b1u, 0x7f4a7c15u);
        break;
    }
    CU(cudaEventRecord(e1, 0));
    CU(cudaEventSynchronize(e1));
    float t = 0;
    CU(cudaEventElapsedTime(&t, e0, e1));
    if (rep > 0 && t < best) best = t;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  const double per_thread = kind == 3 ? 4.0 * iters : (kind == 4 ? 128.0 * iters : 64.0 * iters);
  *gops = per_thread * blocks * threads / (best * 1e-3) / 1e9;
  if (ms_out) *ms_out = best;
  return 0;
}


}  // namespace be
}  // namespace xntt
