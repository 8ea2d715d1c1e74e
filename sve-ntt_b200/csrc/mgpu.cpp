// SPDX-License-Identifier: Apache-2.0
//
// One transform over several GPUs of a node, hosted entirely in C++ (include/xntt.h: xntt_mgpu_*): one process
// owns G sharded sub-plans, one stream per device, the two alternating exchange buffers of every rank, and orders
// the single cross-device dependency of the six-step split with events - no collective library, no torch.
//
// The composition a user of the reference calls once is RecursiveNTT<..., GenericSVELayer / BlockedGenericSVELayer,
// inner, true>::compute_forward (include/sventt/kernel/recursive.hpp:48-84): column phase on the whole n0 x n1
// matrix, barrier, row phase.  Here the matrix is cut into G column blocks (time domain) / G row blocks (frequency
// domain = contiguous 1/G slices of the bit-reversed output); the barrier between the two phases is the one exchange:
//   forward : rank r: column pass + twiddle, every output word stored straight into its owner's buffer over NVLink
//             (xntt_shard_forward_cols_peer)      -> event C[r]
//             rank r waits for C[0..G)            -> row half on the received tiles (xntt_shard_forward_rows_tiled)
//   inverse : row half with peer stores (xntt_shard_inverse_rows_peer) -> events -> column pass
//             (xntt_shard_inverse_cols_chunk)
// A rank may write into buffer b of its peers only after every rank has finished READING buffer b in the transform
// before last (two buffers alternate): every stream first waits for the "read done" events of that buffer.
//
// Pure C++ over backend.h like plan.cpp, so the same file runs in the host emulator of the CPU test-suite; with every
// entry of the device list equal (e.g. {0, 0, 0, 0}) all ranks share one GPU - that is how the sharded kernels get
// parity-tested on a single-GPU box.
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/xntt.h"
#include "backend.h"

using namespace xntt;

struct xntt_mgpu {
  u32 G = 0;
  int log2_m = 0;
  uint64_t n0 = 0, n1 = 0, local_words = 0;
  bool three_pass = false;
  std::vector<int> dev;
  std::vector<xntt_plan*> plan;
  std::vector<void*> stream;
  std::vector<uint64_t*> exch[2];     // [parity][rank]
  std::vector<uint64_t*> work;        // scratch of the inverse row half of three-pass plans
  std::vector<void*> ev_written; // [rank]: this rank's stores into everybody's current buffer are complete
  std::vector<void*> ev_read[2]; // [parity][rank]: this rank has finished reading its buffer `parity`
  std::vector<char> read_valid[2];
  std::vector<void*> ev_done;    // [rank]: last enqueued work of the rank (host entry points, synchronize)
  // lazily allocated staging of the host entry points
  std::vector<uint64_t*> h_in, h_out;
  mutable std::mutex mu;  // one transform at a time per object (exchange buffers and events are shared state)
  int parity = 0;
};

namespace {

thread_local std::string g_merr;

struct Dev {
  int prev = -1;
  bool ok = true;
  explicit Dev(int d) {
    if (be::get_device(&prev) != 0) {
      ok = false;
      prev = -1;
      return;
    }
    if (prev != d && be::set_device(d) != 0) ok = false;
  }
  ~Dev() {
    if (prev >= 0) be::set_device(prev);
  }
};

#define MBE(call)                                   \
  do {                                              \
    int rc_ = (call);                               \
    if (rc_ != 0) return rc_ == 2 ? XNTT_ERR_ALLOC : XNTT_ERR_CUDA; \
  } while (0)
#define MX(call)                     \
  do {                               \
    int rc_ = (call);                \
    if (rc_ != XNTT_OK) return rc_;  \
  } while (0)

void destroy(xntt_mgpu* g) {
  if (!g) return;
  for (u32 r = 0; r < g->G; ++r) {
    Dev d(g->dev[r]);
    if (r < g->stream.size() && g->stream[r]) be::stream_sync(g->stream[r]);
  }
  for (u32 r = 0; r < g->G; ++r) {
    Dev d(g->dev[r]);
    if (r < g->plan.size()) xntt_plan_destroy(g->plan[r]);
    for (int b = 0; b < 2; ++b) {
      if (r < g->exch[b].size()) be::dev_free(g->exch[b][r]);
      if (r < g->ev_read[b].size() && g->ev_read[b][r]) be::event_destroy(g->ev_read[b][r]);
    }
    if (r < g->work.size()) be::dev_free(g->work[r]);
    if (r < g->h_in.size()) be::dev_free(g->h_in[r]);
    if (r < g->h_out.size()) be::dev_free(g->h_out[r]);
    if (r < g->ev_written.size() && g->ev_written[r]) be::event_destroy(g->ev_written[r]);
    if (r < g->ev_done.size() && g->ev_done[r]) be::event_destroy(g->ev_done[r]);
    if (r < g->stream.size() && g->stream[r]) be::stream_destroy(g->stream[r]);
  }
  delete g;
}

int create(xntt_mgpu* g, const xntt_desc* d, const int32_t* devices, u32 n) {
  g->G = n;
  g->log2_m = (int)d->log2_m;
  g->dev.assign(devices, devices + n);
  g->plan.assign(n, nullptr);
  g->stream.assign(n, nullptr);
  g->work.assign(n, nullptr);
  g->ev_written.assign(n, nullptr);
  g->ev_done.assign(n, nullptr);
  g->h_in.assign(n, nullptr);
  g->h_out.assign(n, nullptr);
  for (int b = 0; b < 2; ++b) {
    g->exch[b].assign(n, nullptr);
    g->ev_read[b].assign(n, nullptr);
    g->read_valid[b].assign(n, 0);
  }
  // every device must be able to store into every other one
  for (u32 a = 0; a < n; ++a)
    for (u32 b = 0; b < n; ++b)
      if (g->dev[a] != g->dev[b]) MBE(be::enable_peer_access(g->dev[a], g->dev[b]));
  for (u32 r = 0; r < n; ++r) {
    xntt_desc dr = *d;
    dr.shard_count = n;
    dr.shard_rank = r;
    dr.device = g->dev[r];
    dr.batch = 1;
    MX(xntt_plan_create(&g->plan[r], &dr));
  }
  uint32_t sp[XNTT_MAX_SPLITS] = {};
  const u32 q = xntt_plan_splits(g->plan[0], sp, XNTT_MAX_SPLITS);
  g->n0 = 1ull << sp[0];
  g->n1 = (1ull << g->log2_m) >> sp[0];
  g->three_pass = q > 2;
  g->local_words = (1ull << g->log2_m) / n;
  for (u32 r = 0; r < n; ++r) {
    Dev dv(g->dev[r]);
    if (!dv.ok) return XNTT_ERR_CUDA;
    MBE(be::stream_create(&g->stream[r]));
    MBE(be::event_create(&g->ev_written[r]));
    MBE(be::event_create(&g->ev_done[r]));
    for (int b = 0; b < 2; ++b) {
      MBE(be::dev_malloc((void**)&g->exch[b][r], g->local_words * sizeof(uint64_t)));
      MBE(be::event_create(&g->ev_read[b][r]));
    }
    if (g->three_pass) MBE(be::dev_malloc((void**)&g->work[r], g->local_words * sizeof(uint64_t)));
  }
  return XNTT_OK;
}

// device-resident shards; enqueues on the ranks' own streams
int run(xntt_mgpu* g, uint64_t* const* dst, const uint64_t* const* src, bool inverse) {
  const u32 G = g->G;
  const int b = g->parity;
  g->parity ^= 1;
  std::vector<uint64_t*> peers(g->exch[b]);
  // writers of buffer b must wait until every rank has finished reading it (transform before last)
  for (u32 r = 0; r < G; ++r) {
    Dev dv(g->dev[r]);
    if (!dv.ok) return XNTT_ERR_CUDA;
    for (u32 s = 0; s < G; ++s)
      if (s != r && g->read_valid[b][s]) MBE(be::stream_wait_event(g->stream[r], g->ev_read[b][s]));
    if (!inverse)
      MX(xntt_shard_forward_cols_peer(g->plan[r], peers.data(), src[r], g->stream[r]));
    else
      MX(xntt_shard_inverse_rows_peer(g->plan[r], peers.data(), src[r], g->work[r], g->stream[r]));
    MBE(be::event_record(g->ev_written[r], g->stream[r]));
  }
  for (u32 r = 0; r < G; ++r) {
    Dev dv(g->dev[r]);
    if (!dv.ok) return XNTT_ERR_CUDA;
    for (u32 s = 0; s < G; ++s)
      if (s != r) MBE(be::stream_wait_event(g->stream[r], g->ev_written[s]));
    if (!inverse)
      MX(xntt_shard_forward_rows_tiled(g->plan[r], dst[r], g->exch[b][r], 1, g->stream[r]));
    else
      MX(xntt_shard_inverse_cols_chunk(g->plan[r], dst[r], g->exch[b][r], 0, 1, g->stream[r]));
    MBE(be::event_record(g->ev_read[b][r], g->stream[r]));
    g->read_valid[b][r] = 1;
    MBE(be::event_record(g->ev_done[r], g->stream[r]));
  }
  return XNTT_OK;
}

int sync_all(xntt_mgpu* g) {
  for (u32 r = 0; r < g->G; ++r) {
    Dev dv(g->dev[r]);
    if (!dv.ok) return XNTT_ERR_CUDA;
    MBE(be::stream_sync(g->stream[r]));
  }
  return XNTT_OK;
}

// Whole transform on host buffers.  forward: src natural order, dst bit-reversed; inverse the other way round.
// Time-domain side = column blocks of the n0 x n1 row-major matrix (strided 2-D copies), frequency-domain side =
// contiguous slices.
int run_host(xntt_mgpu* g, uint64_t* dst, const uint64_t* src, bool inverse) {
  const u32 G = g->G;
  const uint64_t block = g->n1 / G;  // columns per rank
  for (u32 r = 0; r < G; ++r) {
    Dev dv(g->dev[r]);
    if (!dv.ok) return XNTT_ERR_CUDA;
    if (!g->h_in[r]) MBE(be::dev_malloc((void**)&g->h_in[r], g->local_words * sizeof(uint64_t)));
    if (!g->h_out[r]) MBE(be::dev_malloc((void**)&g->h_out[r], g->local_words * sizeof(uint64_t)));
    if (!inverse)
      MBE(be::memcpy2d_h2d(g->h_in[r], block * sizeof(uint64_t), src + r * block, g->n1 * sizeof(uint64_t), block * sizeof(uint64_t),
                           g->n0, g->stream[r]));
    else
      MBE(be::memcpy_h2d(g->h_in[r], src + r * g->local_words, g->local_words * sizeof(uint64_t), g->stream[r]));
  }
  MX(run(g, g->h_out.data(), (const uint64_t* const*)g->h_in.data(), inverse));
  for (u32 r = 0; r < G; ++r) {
    Dev dv(g->dev[r]);
    if (!dv.ok) return XNTT_ERR_CUDA;
    if (!inverse)
      MBE(be::memcpy_d2h(dst + r * g->local_words, g->h_out[r], g->local_words * sizeof(uint64_t), g->stream[r]));
    else
      MBE(be::memcpy2d_d2h(dst + r * block, g->n1 * sizeof(uint64_t), g->h_out[r], block * sizeof(uint64_t), block * sizeof(uint64_t),
                           g->n0, g->stream[r]));
  }
  return sync_all(g);
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int xntt_mgpu_create(xntt_mgpu** out, const xntt_desc* desc, const int32_t* devices, uint32_t n_devices) {
  if (!out || !desc || !devices) return XNTT_ERR_INVALID;
  *out = nullptr;
  if (n_devices < 2 || n_devices > 8 || (n_devices & (n_devices - 1))) return XNTT_ERR_INVALID;
  if (desc->batch > 1 || desc->shard_count > 1) return XNTT_ERR_INVALID;
  for (u32 i = 0; i < n_devices; ++i)
    if (devices[i] < 0) return XNTT_ERR_INVALID;
  xntt_mgpu* g = new (std::nothrow) xntt_mgpu;
  if (!g) return XNTT_ERR_ALLOC;
  xntt_desc d = *desc;
  if (d.n_splits == 0) {
    // the planner's own decomposition, with a first split every rank count divides
    if (d.log2_m < 14) {
      delete g;
      return XNTT_ERR_UNSUPPORTED;
    }
  }
  const int rc = create(g, &d, devices, n_devices);
  if (rc != XNTT_OK) {
    destroy(g);
    return rc;
  }
  *out = g;
  return XNTT_OK;
}

int xntt_mgpu_destroy(xntt_mgpu* g) {
  destroy(g);
  return XNTT_OK;
}

uint32_t xntt_mgpu_devices(const xntt_mgpu* g) { return g ? g->G : 0; }
uint64_t xntt_mgpu_m(const xntt_mgpu* g) { return g ? (1ull << g->log2_m) : 0; }
uint64_t xntt_mgpu_n0(const xntt_mgpu* g) { return g ? g->n0 : 0; }
void* xntt_mgpu_stream(const xntt_mgpu* g, uint32_t rank) { return (g && rank < g->G) ? g->stream[rank] : nullptr; }

int xntt_mgpu_forward(xntt_mgpu* g, uint64_t* const* dst, const uint64_t* const* src) {
  if (!g || !dst || !src) return XNTT_ERR_INVALID;
  std::lock_guard<std::mutex> lock(g->mu);
  return run(g, (uint64_t* const*)dst, (const uint64_t* const*)src, false);
}
int xntt_mgpu_inverse(xntt_mgpu* g, uint64_t* const* dst, const uint64_t* const* src) {
  if (!g || !dst || !src) return XNTT_ERR_INVALID;
  std::lock_guard<std::mutex> lock(g->mu);
  return run(g, (uint64_t* const*)dst, (const uint64_t* const*)src, true);
}
int xntt_mgpu_synchronize(xntt_mgpu* g) {
  if (!g) return XNTT_ERR_INVALID;
  std::lock_guard<std::mutex> lock(g->mu);
  return sync_all(g);
}
int xntt_mgpu_forward_host(xntt_mgpu* g, uint64_t* dst, const uint64_t* src) {
  if (!g || !dst || !src) return XNTT_ERR_INVALID;
  std::lock_guard<std::mutex> lock(g->mu);
  return run_host(g, dst, src, false);
}
int xntt_mgpu_inverse_host(xntt_mgpu* g, uint64_t* dst, const uint64_t* src) {
  if (!g || !dst || !src) return XNTT_ERR_INVALID;
  std::lock_guard<std::mutex> lock(g->mu);
  return run_host(g, dst, src, true);
}

#pragma GCC visibility pop
}  // extern "C"
