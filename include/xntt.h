/* SPDX-License-Identifier: Apache-2.0
 *
 * xntt - C ABI of the B200-native 64-bit NTT that drops in behind the sve-ntt ("sventt") API.
 *
 * The reference (Terminus-IMRC/sve-ntt) is a header-only C++20 template library with no FFI of its
 * own; its boundary for this path is sventt::NTT<kernel_type> (include/sventt/wrapper.hpp:13-83).
 * Each entry point below names the reference interface it stands in for.  The C++ front-end in
 * sve-ntt_b200/host/sventt/ re-creates the reference's template surface on top of these calls.
 *
 * Conventions
 *   - residues are uint64_t in [0, p); forward = natural order in, bit-reversed order out;
 *     inverse = bit-reversed in, natural out, multiplied by inverse_factor^-1 (pass 1 for the
 *     reference's "unscaled" inverse, m for the scaled one) - word for word the function computed by
 *     NTTReference (tests/ntt-reference.hpp:43-83).
 *   - "device" entry points take device pointers and a cudaStream_t passed as void*; they are
 *     asynchronous with respect to the host.  "host" entry points take ordinary host pointers and
 *     include both PCIe copies; they return after the result is in dst.
 *   - a plan owns its twiddle tables and scratch; one plan may be used from one stream at a time.
 *   - every function returns XNTT_OK (0) or a negative xntt_status; nothing throws across the ABI.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     XNTT_ERR_CUDA.
 */
#ifndef XNTT_H_INCLUDED
#define XNTT_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum xntt_status {
  XNTT_OK = 0,
  XNTT_ERR_INVALID = -1,     /* bad argument / shape (reference: std::invalid_argument)            */
  XNTT_ERR_UNSUPPORTED = -2, /* valid in the reference, not implemented here                       */
  XNTT_ERR_ALLOC = -3,       /* device or host allocation failed (reference: std::bad_alloc)       */
  XNTT_ERR_CUDA = -4,        /* CUDA runtime error, see xntt_last_cuda_error()                     */
  XNTT_ERR_STATE = -5        /* direction not enabled in this plan (reference: std::logic_error)   */
} xntt_status;

enum {
  XNTT_ENABLE_FORWARD = 1,
  XNTT_ENABLE_INVERSE = 2,
  /* keep the six-step twiddles as two sqrt(M)-entry tables (two modular products per residue) even where
   * the planner would store the whole twiddle matrix (16 bytes per residue and direction, one product) */
  XNTT_COMPACT_TABLES = 4,
  /* run the transform with Shoup ("fixed point") modular multiplication instead of Montgomery: the reference's
   * alternative modmul type FixedPoint64SVE (include/sventt/modmul/sve/fixed-point-64.hpp:13-69).  Same results; legal
   * for moduli below 2^62 (lazy values in [0, 4p)), where it needs 6 instead of 10 wide multiplies per butterfly.  Ignored
   * (Montgomery kernels run) for larger moduli and for the production modulus, whose kernels have it baked in. */
  XNTT_MODMUL_FIXED_POINT = 8,
  /* Tile shape of the passes.  A pass normally works on tiles of 2^13 residues (64 KiB of shared memory per CTA).  Plans
   * of at most 2^20 residues (m * batch) would leave most SMs without a tile, so their passes run on narrow tiles - a
   * quarter of the residues, four times the CTAs - where the pass length has one (rows up to 2^11, columns up to 2^9;
   * production modulus and runtime Montgomery kernels, unsharded plans), and the planner's own decomposition of such a
   * plan keeps its passes that short.  Same results either way.  These two flags override the size rule (tests,
   * measurements): never narrow / narrow wherever a pass has the variant. */
  XNTT_TILES_WIDE = 16,
  XNTT_TILES_NARROW = 32
};

#define XNTT_MAX_SPLITS 4

/* Transform descriptor: what the reference spells as template arguments
 * (Modulus<p, g>, the transform length m, the layer decomposition, inverse_factor) plus the CUDA
 * device.  Zero-initialise, then fill in. */
typedef struct xntt_desc {
  uint64_t modulus;        /* p, prime; Modulus<p, g> (include/sventt/modulus.hpp:14)               */
  uint64_t generator;      /* g, primitive root of p                                                */
  uint32_t log2_m;         /* transform length m = 2^log2_m, 1 <= log2_m <= 31                      */
  uint32_t batch;          /* number of back-to-back transforms in one buffer (0 means 1)           */
  uint64_t inverse_factor; /* inverse output is divided by this (0 or 1: unscaled); the reference's */
                           /* inverse_factor layer argument (layer/sve/radix-eight.hpp:19)          */
  uint32_t flags;          /* XNTT_ENABLE_* (neither bit = both, wrapper.hpp:34-35) | XNTT_COMPACT_TABLES | XNTT_MODMUL_FIXED_POINT | XNTT_TILES_* */
  int32_t device;          /* CUDA device ordinal, -1 = current                                     */
  uint32_t n_splits;       /* 0 = let the planner decompose m; else the six-step decomposition      */
  uint32_t split_log2[XNTT_MAX_SPLITS]; /* m = prod 2^split_log2[i], outermost (column) first       */
  /* sharded plans (one process per GPU): this rank holds columns                                   */
  /* [shard_rank * n1 / shard_count, ...) of the first split's n0 x n1 matrix. 0/0 = not sharded.   */
  uint32_t shard_count;
  uint32_t shard_rank;
  /* budget for whole twiddle matrices (16 bytes per residue and direction), MiB per plan; 0 = default: 512, plus the */
  /* m-entry matrix of the outermost pass of a three-pass plan while it takes at most a quarter of the free memory.  */
  /* A matrix that does not fit keeps the compact two-table form; XNTT_COMPACT_TABLES forces that everywhere. */
  uint32_t twist_table_max_mb;
  uint32_t reserved_;
} xntt_desc;

typedef struct xntt_plan xntt_plan;

/* sventt::NTT<kernel>::NTT(enable_forward, enable_inverse, huge_pages)  (wrapper.hpp:34-46):
 * builds every twiddle table on the device. */
int xntt_plan_create(xntt_plan** plan, const xntt_desc* desc);
int xntt_plan_destroy(xntt_plan* plan);

/* sventt::NTT<kernel>::get_m()  (wrapper.hpp:48) */
uint64_t xntt_plan_m(const xntt_plan* plan);
uint32_t xntt_plan_batch(const xntt_plan* plan);
/* number of kernel launches one forward / inverse call makes (for bench accounting) */
uint32_t xntt_plan_launches(const xntt_plan* plan, int inverse);
/* arithmetic the pass kernels of this plan run: 0 = Montgomery (PAdic64), 1 = Shoup (FixedPoint64, XNTT_MODMUL_FIXED_POINT) */
uint32_t xntt_plan_modmul(const xntt_plan* plan);
/* how the six-step twiddle of column pass `pass` is applied in the given direction: 0 = not a column pass, 1 = two
 * sqrt(M)-entry tables (two modular products per residue), 2 = whole matrix, by the pass itself (one product),
 * 3 = whole matrix, by the pass next to it while that loads / before it stores (one product; this pass is twist-free) */
uint32_t xntt_plan_twiddle_form(const xntt_plan* plan, uint32_t pass, int inverse);
/* log2 of the residues one CTA of pass `pass` works on (13 for whole tiles and less for short passes, 11 or less for
 * narrow tiles); 0 for an invalid pass index */
uint32_t xntt_plan_tile_log2(const xntt_plan* plan, uint32_t pass);
/* fills out[0..n) with the log2 sizes of the passes, returns the pass count */
uint32_t xntt_plan_splits(const xntt_plan* plan, uint32_t* out, uint32_t n);

/* sventt::NTT<kernel>::compute_forward(dst, src) / compute_forward(dst)  (wrapper.hpp:50-65).
 * dst == src is the in-place form.  Device pointers, 16-byte aligned, m * batch words. */
int xntt_forward(const xntt_plan* plan, uint64_t* dst, const uint64_t* src, void* stream);
/* sventt::NTT<kernel>::compute_inverse(dst, src) / compute_inverse(dst)  (wrapper.hpp:67-82) */
int xntt_inverse(const xntt_plan* plan, uint64_t* dst, const uint64_t* src, void* stream);

/* Forward transform with the point-wise product of a polynomial multiply fused into its last pass:
 * dst = multiply_normalize(forward(src), b_mont) word by word, b_mont being a to_montgomery'd
 * spectrum in the same bit-reversed order (examples/magic-series/gaussian-polynomial.hpp:196-212:
 * compute_forward followed by the multiply_normalize loop).  xntt_inverse of dst then yields the
 * cyclic product. */
int xntt_forward_multiply(const xntt_plan* plan, uint64_t* dst, const uint64_t* src, const uint64_t* b_mont,
                          void* stream);

/* One pass of a plan on its own (profiling / per-kernel timing): pass index in forward order,
 * inverse != 0 runs the inverse kernel of that pass.  In place or src -> dst like the full calls.
 * What a pass leaves between two passes is an internal format: a column pass whose successor begins with a modular
 * product stores residues as any 64-bit value congruent to them ("lazy"); only complete transforms promise [0, p). */
int xntt_run_pass(const xntt_plan* plan, uint32_t pass, int inverse, uint64_t* dst, const uint64_t* src,
                  void* stream);

/* The same two calls on host buffers: H2D copy, transform, D2H copy, synchronise. */
int xntt_forward_host(const xntt_plan* plan, uint64_t* dst, const uint64_t* src);
int xntt_inverse_host(const xntt_plan* plan, uint64_t* dst, const uint64_t* src);

/* Sharded plans only: the two local halves of a forward (A: column passes on the local column
 * block, then the caller's all-to-all, then B: row transforms on the local row block) and of an
 * inverse (B^-1, all-to-all, A^-1).  Layouts are described in DESIGN.md. */
int xntt_shard_forward_cols(const xntt_plan* plan, uint64_t* dst, const uint64_t* src, void* stream);
int xntt_shard_forward_rows(const xntt_plan* plan, uint64_t* dst, const uint64_t* src, void* stream);
int xntt_shard_inverse_rows(const xntt_plan* plan, uint64_t* dst, const uint64_t* src, void* stream);
int xntt_shard_inverse_cols(const xntt_plan* plan, uint64_t* dst, const uint64_t* src, void* stream);

/* Exchange-friendly variants.  The local column block is cut into nchunks
 * column sub-blocks of w = n1 / (shard_count * nchunks) columns.
 *   forward_cols_chunk : column pass + twiddle of chunk c, written compactly at
 *                        tiles + c * n0 * w as [n0][w] - shard_count contiguous messages, so the
 *                        caller can start that chunk's all-to-all while the next chunk computes
 *   forward_rows_tiled : after all chunks arrived (message from rank s of chunk c at
 *                        tiles + c * n0 * w + s * (n0 / shard_count) * w): the row half, reading that
 *                        tiled layout in place of a separate unpacking pass; dst in natural order
 *   inverse_rows_tiled : row half of the inverse, leaving its result in the tiled layout (work: scratch
 *                        of m / shard_count words, needed for plans of three passes, else may be null)
 *   inverse_cols_chunk : column pass of chunk c from the compact chunk layout into the column block */
int xntt_shard_forward_cols_chunk(const xntt_plan* plan, uint64_t* tiles, const uint64_t* src, uint32_t chunk,
                                  uint32_t nchunks, void* stream);
int xntt_shard_forward_rows_tiled(const xntt_plan* plan, uint64_t* dst, const uint64_t* tiles, uint32_t nchunks,
                                  void* stream);
int xntt_shard_inverse_rows_tiled(const xntt_plan* plan, uint64_t* tiles, const uint64_t* src, uint64_t* work,
                                  uint32_t nchunks, void* stream);
int xntt_shard_inverse_cols_chunk(const xntt_plan* plan, uint64_t* dst, const uint64_t* tiles, uint32_t chunk,
                                  uint32_t nchunks, void* stream);

/* Fused compute + exchange over peer memory (NVLink / NVSwitch): the pass next to the all-to-all
 * stores its output directly into every rank's buffer, so no collective is issued at all.
 * peers[s] = device address, valid on THIS GPU, of rank s's exchange buffer (m / shard_count words,
 * e.g. torch symmetric memory or cudaIpc mappings), s = 0 .. shard_count-1 (at most 8).
 *   forward_cols_peer : column pass + twiddle; rank s receives my rows of its row block at
 *                       peers[s] + rank * (n0/G) * (n1/G), i.e. the tiled layout with one chunk -
 *                       after a cross-rank barrier, xntt_shard_forward_rows_tiled(dst, own buffer, 1)
 *   inverse_rows_peer : row half of the inverse; rank s receives its columns of my rows in chunk
 *                       layout - after a barrier, xntt_shard_inverse_cols_chunk(dst, own buffer, 0, 1)
 * The caller orders the barriers (all writers done before anyone reads; readers done before the
 * buffer is written again - two alternating buffers make the second barrier unnecessary). */
int xntt_shard_forward_cols_peer(const xntt_plan* plan, uint64_t* const* peers, const uint64_t* src, void* stream);
int xntt_shard_inverse_rows_peer(const xntt_plan* plan, uint64_t* const* peers, const uint64_t* src, uint64_t* work,
                                 void* stream);

/* One transform over several GPUs of ONE process (C++-hosted; no collective library, no torch): the library owns one
 * sharded sub-plan, one stream and two alternating exchange buffers per device, enables peer access between the devices
 * and orders the single exchange of the six-step split with events.  It stands in for the one call the reference makes
 * for the whole composition, RecursiveNTT<..., (Blocked)GenericSVELayer, inner, true>::compute_forward / compute_inverse
 * (include/sventt/kernel/recursive.hpp:48-84, 103-140), whose global transposition (layer/sve/generic.hpp:112-161) is the
 * exchange.  desc is an ordinary transform descriptor (batch <= 1, shard fields unused, any modulus); devices[]
 * lists n_devices (2, 4 or 8) CUDA ordinals, rank r = devices[r].  The same ordinal may appear more than once: all ranks
 * then share that GPU (how the sharded kernels are parity-tested on a single-GPU box).
 *   xntt_mgpu_forward / _inverse : device-resident shards.  Time domain: rank r holds the column block
 *       A[:, r*n1/G .. (r+1)*n1/G) of the n0 x n1 row-major matrix as [n0][n1/G] (n0 = xntt_mgpu_n0()); frequency domain:
 *       rank r holds the r-th contiguous 1/G of the bit-reversed output.  forward: src = time, dst = frequency; inverse
 *       the other way round.  Work is enqueued on the library's per-rank streams (xntt_mgpu_stream()) and is asynchronous;
 *       xntt_mgpu_synchronize() waits for all ranks.  dst[r] / src[r] live on devices[r].
 *   xntt_mgpu_forward_host / _inverse_host : the whole transform on host buffers of m words (natural order in,
 *       bit-reversed out, and back), scattered to / gathered from the GPUs; returns when dst is complete. */
typedef struct xntt_mgpu xntt_mgpu;
int xntt_mgpu_create(xntt_mgpu** mgpu, const xntt_desc* desc, const int32_t* devices, uint32_t n_devices);
int xntt_mgpu_destroy(xntt_mgpu* mgpu);
uint32_t xntt_mgpu_devices(const xntt_mgpu* mgpu);
uint64_t xntt_mgpu_m(const xntt_mgpu* mgpu);
uint64_t xntt_mgpu_n0(const xntt_mgpu* mgpu);
void* xntt_mgpu_stream(const xntt_mgpu* mgpu, uint32_t rank);
int xntt_mgpu_forward(xntt_mgpu* mgpu, uint64_t* const* dst, const uint64_t* const* src);
int xntt_mgpu_inverse(xntt_mgpu* mgpu, uint64_t* const* dst, const uint64_t* const* src);
int xntt_mgpu_synchronize(xntt_mgpu* mgpu);
int xntt_mgpu_forward_host(xntt_mgpu* mgpu, uint64_t* dst, const uint64_t* src);
int xntt_mgpu_inverse_host(xntt_mgpu* mgpu, uint64_t* dst, const uint64_t* src);

/* PAdic64 element-wise helpers on device buffers (count words each):
 *   to_montgomery      : dst[i] = src[i] * 2^64 mod p     (modmul/sve/p-adic-64.hpp:19-22)
 *   from_montgomery    : dst[i] = src[i] * 2^-64 mod p    (p-adic-64.hpp:24-38, canonical result)
 *   multiply_normalize : dst[i] = a[i] * b_mont[i] * 2^-64 mod p, canonical  (p-adic-64.hpp:101-115);
 *                        with b_mont = to_montgomery(b) this is the point-wise product a[i]*b[i]
 *                        used between the transforms of a polynomial multiply
 *                        (examples/magic-series/gaussian-polynomial.hpp:201-212). */
int xntt_to_montgomery(const xntt_plan* plan, uint64_t* dst, const uint64_t* src, size_t count, void* stream);
int xntt_from_montgomery(const xntt_plan* plan, uint64_t* dst, const uint64_t* src, size_t count, void* stream);
int xntt_multiply_normalize(const xntt_plan* plan, uint64_t* dst, const uint64_t* a, const uint64_t* b_mont,
                            size_t count, void* stream);

/* Transpose*::transpose(dst, src, src_rows, src_cols, ld_dst, ld_src)
 * (every class under include/sventt/transposition/sve/): dst[ld_dst * c + r] = src[ld_src * r + c] on device buffers;
 * dst == src with rows == cols and ld_dst == ld_src is the in-place square form transpose(dst, dim). */
int xntt_transpose(uint64_t* dst, const uint64_t* src, uint64_t rows, uint64_t cols, uint64_t ld_dst, uint64_t ld_src,
                   void* stream);

/* Kinnaes' formula for the number of magic series modulo `modulus` - the reference's second example
 * (examples/magic-series-kinnaes/kinnaes.hpp), a pure PAdic64 multiply workload at a root of unity of odd order:
 *   xntt_kinnaes_sum     : MagicSeriesKinnaes<m, PAdic64<Modulus<modulus, generator>>, n>::compute_sum(j_begin, j_end)
 *                          (kinnaes.hpp:51-157), 0 <= j_begin <= j_end <= n / 2
 *   xntt_kinnaes_compute : ...::compute() (kinnaes.hpp:27-34) = (2 compute_sum(0, n / 2) + binomial(m^2, m)) / n
 * n must divide modulus - 1 (else XNTT_ERR_INVALID, like Modulus::get_root_forward), 2 <= m < 2^20.
 * device = CUDA ordinal, -1 = current.  The result is a canonical residue. */
int xntt_kinnaes_sum(uint64_t modulus, uint64_t generator, uint64_t m, uint64_t n, uint64_t j_begin, uint64_t j_end,
                     int device, uint64_t* result);
int xntt_kinnaes_compute(uint64_t modulus, uint64_t generator, uint64_t m, uint64_t n, int device, uint64_t* result);

/* Frees the per-device result buffers the Kinnaes entry points keep between calls (a cudaMalloc / cudaFree pair per
 * call would cost thirty times the kernel).  Call before cudaDeviceReset / at unload; later calls re-create them. */
int xntt_release_scratch(void);

/* Device-resident PageMemory twin (include/sventt/vector.hpp:61-168): pinned host memory for the
 * host entry points / device memory for the device ones. */
int xntt_alloc_device(void** ptr, size_t bytes, int device);
int xntt_free_device(void* ptr);
int xntt_alloc_pinned(void** ptr, size_t bytes);
int xntt_free_pinned(void* ptr);
int xntt_memcpy_h2d(void* dst_device, const void* src_host, size_t bytes, void* stream);
int xntt_memcpy_d2h(void* dst_host, const void* src_device, size_t bytes, void* stream);
int xntt_stream_synchronize(void* stream);
/* 1 if ptr is device (or managed) memory, 0 if it is ordinary/pinned host memory, < 0 on error */
int xntt_pointer_is_device(const void* ptr);

const char* xntt_strerror(int status);
/* last cudaError_t seen by this thread inside the library, as text */
const char* xntt_last_cuda_error(void);
/* "xntt <version> sm_100a" */
const char* xntt_version(void);
/* number of CUDA devices visible, or a negative xntt_status */
int xntt_device_count(void);

/* Micro-benchmarks that calibrate the integer roofline (tools/ and bench.py): run `iters`
 * dependent-free rounds of the named instruction mix in registers on every SM and return the
 * achieved rate in giga-operations per second in *gops.  kind: 0 = IMAD (32-bit mad.lo),
 * 1 = IMAD.WIDE, 2 = IADD3, 3 = Montgomery butterflies (gops = butterflies/s / 1e9), 4 = IMAD and LOP3 interleaved,
 * 5 = IMAD.HI. */
int xntt_microbench(int kind, int iters, double* gops, double* ms);

#ifdef __cplusplus
}
#endif

#endif /* XNTT_H_INCLUDED */
