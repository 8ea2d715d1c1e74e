// SPDX-License-Identifier: Apache-2.0
// The device back-end as the planner sees it: memory, streams and kernel launches, with no CUDA
// types in the signatures.  backend_cuda.cu is the product implementation (sm_100a kernels);
// tests/emu/backend_emu.cpp executes the very same kernel templates on the host, one emulated
// thread at a time, so that the planner and the index algebra are covered by the CPU test-suite.
#pragma once
#include "params.h"

namespace xntt {
namespace be {

// every function returns 0 on success; on failure last_error() describes it
int device_count(int* n);
int get_device(int* dev);
int set_device(int dev);
int dev_malloc(void** p, size_t bytes);   // returns 2 for out-of-memory, 1 for other errors
int dev_free(void* p);
int mem_info(size_t* free_bytes, size_t* total_bytes);  // of the current device
int host_malloc_pinned(void** p, size_t bytes);
int host_free_pinned(void* p);
int memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream);
int memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream);
// strided copies (column blocks of a row-major host matrix <-> compact device blocks); pitches and width in bytes
int memcpy2d_h2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, void* stream);
int memcpy2d_d2h(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, void* stream);
// let kernels running on `dev` store into memory of `peer` (idempotent)
int enable_peer_access(int dev, int peer);
int stream_sync(void* stream);
// streams and events for the chunk pipeline behind the host entry points of batched plans
int stream_create(void** stream);
int stream_destroy(void* stream);
int event_create(void** event);
int event_destroy(void* event);
int event_record(void* event, void* stream);
int stream_wait_event(void* stream, void* event);
int pointer_is_device(const void* p, int* is_device);
// *dev = the address kernels can use for page-locked, mapped host memory at p (with unified addressing: p itself),
// nullptr for pageable host memory (or anything else a kernel must not touch)
int host_device_pointer(const void* p, void** dev);
const char* last_error();

// map = true: use prm.smap / prm.dmap (generalised addressing; built for kP0 only)
int launch_pass(int logn, bool col, bool inverse, bool map, const PassParams& prm, unsigned grid, void* stream);
// the field is chosen by fc.p: kP0 runs the kernels with the modulus baked in, anything else the
// runtime-modulus kernels (PassParams carries its own copy in prm.field)
int launch_gen_table(const FieldConsts& fc, Tw* out, u32 count, int kind, int logn, int shift, const PowTable& t,
                     void* stream);
int launch_to_mont(const FieldConsts& fc, u64* dst, const u64* src, size_t n, u64 r2, void* stream);
int launch_from_mont(const FieldConsts& fc, u64* dst, const u64* src, size_t n, void* stream);
int launch_mulnorm(const FieldConsts& fc, u64* dst, const u64* a, const u64* b, size_t n, void* stream);
// dst[ld_dst * c + r] = src[ld_src * r + c]; dst == src with rows == cols and equal leading dimensions
// is the in-place square form
int launch_transpose(u64* dst, const u64* src, u64 rows, u64 cols, u64 ld_dst, u64 ld_src, void* stream);
// Kinnaes sum: fills prm.partial[0 .. 2 * blocks) (device memory) with one Montgomery-form fraction per CTA
int launch_kinnaes(const KinnaesParams& prm, unsigned blocks, void* stream);
int microbench(int kind, int iters, double* gops, double* ms);

}  // namespace be
}  // namespace xntt
