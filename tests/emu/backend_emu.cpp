// SPDX-License-Identifier: Apache-2.0
//
// TEST INFRASTRUCTURE ONLY.  A host implementation of sve-ntt_b200/csrc/backend.h that executes
// the *same* kernel templates (pass_kernel.cuh, misc_kernels.cuh) one emulated CUDA thread at a
// time.  Linked with the unchanged planner (plan.cpp) it yields libxntt_emu.so, which the
// `-m "not gpu"` tests compare against the oracle: this covers the planner, the table
// generation, the stage/twiddle index algebra and the lazy arithmetic identities without a GPU.
// It is never loaded by the product, by bench.py's timed path or by the -m gpu tests.
#define XNTT_HOST_EMU 1
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

// --- minimal CUDA vocabulary for the kernel templates ---------------------------------------
struct ulonglong2 {
  unsigned long long x, y;
};
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return {x, y}; }
struct EmuIdx {
  unsigned x = 0, y = 0, z = 0;
};
static thread_local EmuIdx threadIdx, blockIdx;
static inline void __syncthreads() {}
template <class T>
static inline T __ldg(const T* p) {
  return *p;
}
static inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
#define __restrict__
#define __global__
#define __grid_constant__

#include "../../sve-ntt_b200/csrc/backend.h"
#include "../../sve-ntt_b200/csrc/kinnaes_kernel.cuh"
#include "../../sve-ntt_b200/csrc/misc_kernels.cuh"
#include "../../sve-ntt_b200/csrc/pass_kernel.cuh"
#include "../../sve-ntt_b200/csrc/transpose_kernel.cuh"

namespace xntt {

// Emulated schedule.  A stage ends with a barrier over barrier_group<>() consecutive threads (pass_kernel.cuh:
// stage_barrier), not necessarily the whole CTA.  The emulator runs the most eager schedule those barriers allow:
// a group of threads goes as deep into the stages as its barriers permit before the next group has run anything -
// so a barrier group that is too small (a stage reading a slot another group has not written yet, or overwriting one
// another group still has to read) produces wrong words, the tile having been poisoned beforehand.
template <class F, class Cfg, bool INV, int TWIST, int J>
static void emu_stage_threads(const PassParams& prm, typename Slot<Cfg::C>::type* sm, const u64* gsrc, u64* gdst,
                              u32 col0, u32 row0, unsigned t0, unsigned t1) {
  for (unsigned t = t0; t < t1; ++t) {
    threadIdx.x = t;
    if constexpr (INV)
      inv_stage<F, Cfg, TWIST, J>(prm, sm, gsrc, gdst, col0, row0);
    else
      fwd_stage<F, Cfg, TWIST, J>(prm, sm, gsrc, gdst, col0, row0);
  }
}

// forward: groups shrink from stage to stage.  [t0, t1) may start stage J; each barrier group of the barrier
// after stage J runs stage J and then goes on alone
template <class F, class Cfg, int TWIST, int J>
static void emu_fwd_from(const PassParams& prm, typename Slot<Cfg::C>::type* sm, const u64* gsrc, u64* gdst, u32 col0,
                         u32 row0, unsigned t0, unsigned t1) {
  unsigned step = 1;  // after the last stage nobody waits for anybody
  if constexpr (J + 1 < Cfg::NS) step = barrier_group<Cfg, false, J>();
  if (step > t1 - t0) step = t1 - t0;
  for (unsigned g = t0; g < t1; g += step) {
    emu_stage_threads<F, Cfg, false, TWIST, J>(prm, sm, gsrc, gdst, col0, row0, g, g + step);
    if constexpr (J + 1 < Cfg::NS) emu_fwd_from<F, Cfg, TWIST, J + 1>(prm, sm, gsrc, gdst, col0, row0, g, g + step);
  }
}
// inverse: groups grow.  [t0, t1) is one barrier group of the barrier after stage J - 1 (a single thread for
// J = 0): bring each of its sub-groups through stage J - 1, then run stage J on it
template <class F, class Cfg, int TWIST, int J>
static void emu_inv_upto(const PassParams& prm, typename Slot<Cfg::C>::type* sm, const u64* gsrc, u64* gdst, u32 col0,
                         u32 row0, unsigned t0, unsigned t1) {
  if constexpr (J > 0) {
    unsigned step = 1;
    if constexpr (J > 1) step = barrier_group<Cfg, true, J - 2>();
    if (step > t1 - t0) step = t1 - t0;
    for (unsigned g = t0; g < t1; g += step) emu_inv_upto<F, Cfg, TWIST, J - 1>(prm, sm, gsrc, gdst, col0, row0, g, g + step);
  }
  emu_stage_threads<F, Cfg, true, TWIST, J>(prm, sm, gsrc, gdst, col0, row0, t0, t1);
}

template <class F, class Cfg, bool INV, int TWIST, int... Js>
static void emu_stages(const PassParams& prm, typename Slot<Cfg::C>::type* sm, const u64* gsrc, u64* gdst,
                       u32 col0, u32 row0, std::integer_sequence<int, Js...>) {
  if constexpr (INV) {
    unsigned step = 1;
    if constexpr (Cfg::NS > 1) step = barrier_group<Cfg, true, Cfg::NS - 2>();
    for (unsigned g = 0; g < (unsigned)kThreads; g += step)
      emu_inv_upto<F, Cfg, TWIST, Cfg::NS - 1>(prm, sm, gsrc, gdst, col0, row0, g, g + step);
  } else {
    emu_fwd_from<F, Cfg, TWIST, 0>(prm, sm, gsrc, gdst, col0, row0, 0, (unsigned)kThreads);
  }
}

template <class F, int LOGN, bool COL, bool INV, bool MAP, bool NARROW = false>
static int emu_launch2(const PassParams& prm, unsigned grid) {
  constexpr int LOGW = pass_logw(LOGN, NARROW), C = pass_c(LOGN, NARROW);
  typedef PassCfg<LOGN, LOGW, C, COL, MAP> Cfg;
  std::vector<typename Slot<C>::type> sm((size_t)Cfg::N * Cfg::NP + 1);
  for (unsigned tile = 0; tile < grid; ++tile) {
    // body of pass_kernel()
    u64 sbase, dbase;
    u32 col0 = 0, row0 = 0;
    tile_origin<Cfg>(prm, tile, sbase, dbase, col0, row0);
    // poison shared memory so that a missing write shows up
    memset(sm.data(), 0xcd, sm.size() * sizeof(sm[0]));
    // same choice of the fused extras as dispatch.cuh: launch_one
    const auto seq = std::make_integer_sequence<int, Cfg::NS>{};
    const u64 *gs = prm.src + sbase;
    u64* gd = prm.dst + dbase;
    switch (pass_kind(COL, INV, MAP, prm)) {
      case kCompactTwist: emu_stages<F, Cfg, INV, kCompactTwist>(prm, sm.data(), gs, gd, col0, row0, seq); break;
      case kFullTwist: emu_stages<F, Cfg, INV, kFullTwist>(prm, sm.data(), gs, gd, col0, row0, seq); break;
      case kPointwise: emu_stages<F, Cfg, INV, kPointwise>(prm, sm.data(), gs, gd, col0, row0, seq); break;
      case kPreTwist: emu_stages<F, Cfg, INV, kPreTwist>(prm, sm.data(), gs, gd, col0, row0, seq); break;
      case kPrePointwise: emu_stages<F, Cfg, INV, kPrePointwise>(prm, sm.data(), gs, gd, col0, row0, seq); break;
      case kPostTwist: emu_stages<F, Cfg, INV, kPostTwist>(prm, sm.data(), gs, gd, col0, row0, seq); break;
      case kColPre: emu_stages<F, Cfg, INV, kColPre>(prm, sm.data(), gs, gd, col0, row0, seq); break;
      default: emu_stages<F, Cfg, INV, kNoTwist>(prm, sm.data(), gs, gd, col0, row0, seq); break;
    }
  }
  return 0;
}

static bool g_map = false;
template <int LOGN, bool COL, bool INV>
static int emu_launch(const PassParams& prm, unsigned grid) {
  // same choice as backend_cuda.cu: baked-in modulus for kP0, runtime modulus otherwise
  if (g_map) {
    if (prm.field.p == kP0) return emu_launch2<F0, LOGN, COL, INV, true>(prm, grid);
    if (prm.field.kind == kFieldShoup) return emu_launch2<FieldShoup, LOGN, COL, INV, true>(prm, grid);
    return emu_launch2<FieldRT, LOGN, COL, INV, true>(prm, grid);
  }
  if (prm.narrow) {
    // same rule as backend_cuda.cu: production modulus or runtime Montgomery, pass lengths with a narrow tile
    if constexpr (has_narrow_tile(LOGN, COL)) {
      if (field_has_narrow(prm.field)) {
        if (prm.field.p == kP0) return emu_launch2<F0, LOGN, COL, INV, false, true>(prm, grid);
        return emu_launch2<FieldRT, LOGN, COL, INV, false, true>(prm, grid);
      }
    }
    return 1;
  }
  if (prm.field.p == kP0) return emu_launch2<F0, LOGN, COL, INV, false>(prm, grid);
  if (prm.field.kind == kFieldShoup) return emu_launch2<FieldShoup, LOGN, COL, INV, false>(prm, grid);
  return emu_launch2<FieldRT, LOGN, COL, INV, false>(prm, grid);
}

#define EMU_CASE(L)                                                       \
  case L:                                                                 \
    if (col)                                                              \
      return inverse ? emu_launch<L, true, true>(prm, grid) : emu_launch<L, true, false>(prm, grid); \
    else                                                                  \
      return inverse ? emu_launch<L, false, true>(prm, grid) : emu_launch<L, false, false>(prm, grid);


// The pass-kernel instantiations, split over four translation units by logn mod 4: this file is compiled once per
// part with -DEMU_PART=k (only emu_launch_part<k> is emitted) and once without (everything else).
int emu_launch_part0(int logn, bool col, bool inverse, bool map, const PassParams& prm, unsigned grid);
int emu_launch_part1(int logn, bool col, bool inverse, bool map, const PassParams& prm, unsigned grid);
int emu_launch_part2(int logn, bool col, bool inverse, bool map, const PassParams& prm, unsigned grid);
int emu_launch_part3(int logn, bool col, bool inverse, bool map, const PassParams& prm, unsigned grid);

#ifdef EMU_PART
#define EMU_PART_NAME2(k) emu_launch_part##k
#define EMU_PART_NAME(k) EMU_PART_NAME2(k)
int EMU_PART_NAME(EMU_PART)(int logn, bool col, bool inverse, bool map, const PassParams& prm, unsigned grid) {
  g_map = map;
  switch (logn) {
#if EMU_PART == 0
    EMU_CASE(4) EMU_CASE(8) EMU_CASE(12)
#elif EMU_PART == 1
    EMU_CASE(1) EMU_CASE(5) EMU_CASE(9)
    case 13:
      if (col) break;
      return inverse ? emu_launch<13, false, true>(prm, grid) : emu_launch<13, false, false>(prm, grid);
#elif EMU_PART == 2
    EMU_CASE(2) EMU_CASE(6) EMU_CASE(10)
#else
    EMU_CASE(3) EMU_CASE(7) EMU_CASE(11)
#endif
    default:
      break;
  }
  return -1;
}
#endif  // EMU_PART

#ifndef EMU_PART
namespace be {

static std::string g_err = "no error";

int device_count(int* n) {
  *n = 1;
  return 0;
}
int get_device(int* dev) {
  *dev = 0;
  return 0;
}
int set_device(int) { return 0; }
int dev_malloc(void** p, size_t bytes) {
  *p = aligned_alloc(64, (bytes + 63) / 64 * 64 + 64);
  if (!*p) {
    g_err = "out of memory";
    return 2;
  }
  memset(*p, 0xab, bytes);
  return 0;
}
int dev_free(void* p) {
  free(p);
  return 0;
}
int mem_info(size_t* free_bytes, size_t* total_bytes) {
  *free_bytes = *total_bytes = (size_t)8 << 30;
  return 0;
}
int host_malloc_pinned(void** p, size_t bytes) { return dev_malloc(p, bytes); }
int host_free_pinned(void* p) { return dev_free(p); }
int memcpy_h2d(void* dst, const void* src, size_t bytes, void*) {
  memmove(dst, src, bytes);
  return 0;
}
int memcpy_d2h(void* dst, const void* src, size_t bytes, void*) {
  memmove(dst, src, bytes);
  return 0;
}
int memcpy2d_h2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, void*) {
  for (size_t r = 0; r < height; ++r) memmove((char*)dst + r * dpitch, (const char*)src + r * spitch, width);
  return 0;
}
int memcpy2d_d2h(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, void* st) {
  return memcpy2d_h2d(dst, dpitch, src, spitch, width, height, st);
}
int enable_peer_access(int, int) { return 0; }
int stream_sync(void*) { return 0; }
// the emulator executes every "launch" and copy at once, in program order: streams and events are tokens
int stream_create(void** st) {
  *st = nullptr;
  return 0;
}
int stream_destroy(void*) { return 0; }
int event_create(void** ev) {
  *ev = nullptr;
  return 0;
}
int event_destroy(void*) { return 0; }
int event_record(void*, void*) { return 0; }
int stream_wait_event(void*, void*) { return 0; }
int pointer_is_device(const void*, int* is_device) {
  *is_device = 0;
  return 0;
}
// the emulator's "device" is host memory: every host buffer counts as mapped (exercises the zero-copy host paths)
int host_device_pointer(const void* p, void** dev) {
  *dev = const_cast<void*>(p);
  return 0;
}
const char* last_error() { return g_err.c_str(); }

int launch_pass(int logn, bool col, bool inverse, bool map, const PassParams& prm, unsigned grid, void*) {
  // the pass kernels are instantiated in four translation units (compiled in parallel), by logn mod 4
  int rc = -1;
  if (logn >= 1 && logn <= 13 && !(col && logn == 13)) {
    switch (logn & 3) {
      case 0: rc = emu_launch_part0(logn, col, inverse, map, prm, grid); break;
      case 1: rc = emu_launch_part1(logn, col, inverse, map, prm, grid); break;
      case 2: rc = emu_launch_part2(logn, col, inverse, map, prm, grid); break;
      default: rc = emu_launch_part3(logn, col, inverse, map, prm, grid); break;
    }
  }
  if (rc == 0) return 0;
  g_err = rc < 0 ? "invalid pass length" : "pass kernel not available in this form";
  return 1;
}

#define EMU_WITH_FIELD(fc, call)     \
  do {                               \
    if ((fc).p == kP0) {             \
      typedef F0 F;                  \
      const F f = make_field<F>(fc); \
      call;                          \
    } else if ((fc).kind == kFieldShoup) { \
      typedef FieldShoup F;          \
      const F f = make_field<F>(fc); \
      call;                          \
    } else {                         \
      typedef FieldRT F;             \
      const F f = make_field<F>(fc); \
      call;                          \
    }                                \
  } while (0)

int launch_gen_table(const FieldConsts& fc, Tw* out, u32 count, int kind, int logn, int shift, const PowTable& t,
                     void*) {
  EMU_WITH_FIELD(fc, for (u32 i = 0; i < count; ++i) out[i] = table_entry<F>(f, i, kind, logn, shift, t));
  return 0;
}
template <class F>
static void emu_kinnaes(const F& f, const KinnaesParams& prm, unsigned blocks) {
  // same device functions, CTA by CTA; the fractions of a CTA are folded one after the other instead of in a tree
  for (unsigned b = 0; b < blocks; ++b) {
    u64 ns = 0, ds = f.one();
    for (unsigned t = 0; t < (unsigned)kKinnaesThreads; ++t)
      for (u64 i = (u64)b * kKinnaesThreads + t; i < prm.count; i += (u64)blocks * kKinnaesThreads) {
        u64 n, d;
        kinnaes_term<F>(f, prm, prm.j_first + i, n, d);
        kinnaes_fold<F>(f, ns, ds, n, d);
      }
    prm.partial[2 * b] = ns;
    prm.partial[2 * b + 1] = ds;
  }
}
int launch_kinnaes(const KinnaesParams& prm, unsigned blocks, void*) {
  EMU_WITH_FIELD(prm.field, emu_kinnaes<F>(f, prm, blocks));
  return 0;
}
int launch_to_mont(const FieldConsts& fc, u64* dst, const u64* src, size_t n, u64 r2, void*) {
  EMU_WITH_FIELD(fc, {
    const u64 r2p = f.companion(r2);
    for (size_t i = 0; i < n; ++i) dst[i] = ew_to_mont<F>(f, src[i], r2, r2p);
  });
  return 0;
}
int launch_from_mont(const FieldConsts& fc, u64* dst, const u64* src, size_t n, void*) {
  EMU_WITH_FIELD(fc, for (size_t i = 0; i < n; ++i) dst[i] = ew_from_mont<F>(f, src[i]));
  return 0;
}
int launch_mulnorm(const FieldConsts& fc, u64* dst, const u64* a, const u64* b, size_t n, void*) {
  EMU_WITH_FIELD(fc, for (size_t i = 0; i < n; ++i) dst[i] = ew_mulnorm<F>(f, a[i], b[i]));
  return 0;
}
int launch_transpose(u64* dst, const u64* src, u64 rows, u64 cols, u64 ld_dst, u64 ld_src, void*) {
  const u64 tr = (rows + kTrTile - 1) / kTrTile, tc = (cols + kTrTile - 1) / kTrTile;
  std::vector<u64> sa(kTrSmemWords), sb(kTrSmemWords);
  if (dst == src) {
    for (u64 i = 0; i < tc; ++i)
      for (u64 j = i; j < tc; ++j) {
        const u64 r0 = i * kTrTile, c0 = j * kTrTile;
        for (int t = 0; t < kTrThreads; ++t) {
          tr_load(dst, ld_dst, rows, cols, r0, c0, sa.data(), t);
          if (i != j) tr_load(dst, ld_dst, rows, cols, c0, r0, sb.data(), t);
        }
        for (int t = 0; t < kTrThreads; ++t) {
          tr_store(dst, ld_dst, rows, cols, r0, c0, sa.data(), t);
          if (i != j) tr_store(dst, ld_dst, rows, cols, c0, r0, sb.data(), t);
        }
      }
    return 0;
  }
  for (u64 a = 0; a < tr; ++a)
    for (u64 b = 0; b < tc; ++b) {
      for (int t = 0; t < kTrThreads; ++t) tr_load(src, ld_src, rows, cols, a * kTrTile, b * kTrTile, sa.data(), t);
      for (int t = 0; t < kTrThreads; ++t) tr_store(dst, ld_dst, rows, cols, a * kTrTile, b * kTrTile, sa.data(), t);
    }
  return 0;
}
int microbench(int, int, double*, double*) {
  g_err = "no device";
  return 1;
}

}  // namespace be
#endif  // !EMU_PART

}  // namespace xntt
