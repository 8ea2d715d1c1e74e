# SPDX-License-Identifier: Apache-2.0
"""One transform sharded over the GPUs of a node: the six-step column/row split with a single
all-to-all (the reference has no distributed code - its only analogue is the
global transposition of GenericSVELayer, include/sventt/layer/sve/generic.hpp:112-161).

Layout, m = n0 * n1 viewed as an n0 x n1 row-major matrix, G ranks:
  time domain      rank r holds the column block  A[:, r*n1/G : (r+1)*n1/G]   as [n0][n1/G]
  frequency domain rank r holds rows  [r*n0/G, (r+1)*n0/G)  of the transformed matrix as [n0/G][n1],
                   i.e. the r-th contiguous 1/G of the reference's bit-reversed output.
forward  = column passes + twiddle (local) -> all-to-all of (n0/G x n1/G) tiles -> row passes (local)
inverse  = row passes -> all-to-all -> column passes; back in the column-block layout.

Exchange, fastest first:
  "peer"      the pass next to the exchange stores its output directly into every rank's buffer over
              NVLink (torch symmetric memory provides the peer mappings and the stream-ordered
              barrier): compute and all-to-all are ONE kernel, no collective is issued
              (xntt_shard_forward_cols_peer / xntt_shard_inverse_rows_peer);
  "pipelined" NCCL all_to_all_single per column chunk, overlapped with the next chunk's compute;
  "simple"    whole-block all-to-all + explicit tile interleave (any modulus).

Pipelining: the column block is processed in K column chunks; chunk c's all-to-all (async, NCCL's
own stream) runs while chunk c+1 computes, and the row half reads the received tiles in place
(xntt_shard_*_tiled), so there is no separate pack/unpack pass over memory.

torch.distributed is plumbing here (NCCL all_to_all_single over NVLink on the box, gloo in the CPU
tests); the compute is libxntt's xntt_shard_* entry points.
"""
import contextlib

import torch
import torch.distributed as dist


def _stream_scope(stream):
    """The exchange (symmetric-memory barrier, NCCL all-to-all, tensor copies) runs on torch's CURRENT stream, the
    kernels on the raw handle the caller passes: make the two the same stream.  0 / None = torch's current stream;
    any other handle becomes the current stream for the duration of the call.  Returns (context, raw handle)."""
    if not torch.cuda.is_available():
        return contextlib.nullcontext(), 0
    cur = torch.cuda.current_stream()
    if not stream or stream == cur.cuda_stream:
        return contextlib.nullcontext(), cur.cuda_stream
    return torch.cuda.stream(torch.cuda.ExternalStream(stream)), stream


class ShardedNTT:
    def __init__(self, library, log2_m, world, rank, device=-1, splits=None, group=None, inverse_factor=None,
                 chunks=None, modulus=None, generator=None, mode=None):
        self.world, self.rank, self.group = world, rank, group
        kw = {}
        if modulus is not None:
            kw.update(modulus=modulus, generator=generator)
        self.plan = library.plan(log2_m, splits=splits, shard_count=world, shard_rank=rank, device=device,
                                 inverse_factor=inverse_factor, **kw)
        self.splits = self.plan.splits
        self.n0 = 1 << self.splits[0]
        self.n1 = (1 << log2_m) // self.n0
        self.local_words = (1 << log2_m) // world
        self.tiled = True  # the address-mapped kernels exist for every field flavour
        self.chunks = self._pick_chunks(chunks) if self.tiled else 1
        self.extra_launches_per_roundtrip = 0
        self._tmp = None
        self.mode = mode or ("peer" if self.tiled else "simple")
        if not self.tiled:
            self.mode = "simple"
        self._peer = None  # (buffers, handles, parity)
        if self.mode == "peer":
            try:
                self._setup_peer(device)
            except Exception as exc:  # no P2P / symmetric memory here (e.g. gloo on CPU): NCCL path
                self._peer_error = repr(exc)
                self.mode = "pipelined"

    def _setup_peer(self, device):
        import torch.distributed._symmetric_memory as symm_mem
        if not torch.cuda.is_available() or dist.get_backend(self.group) != "nccl" or self.world > 8:
            raise RuntimeError("peer mode needs CUDA + NCCL and at most 8 ranks")
        dev = torch.device("cuda", torch.cuda.current_device() if device is None or device < 0 else device)
        group = self.group if self.group is not None else dist.group.WORLD
        bufs, hdls = [], []
        for _ in range(2):  # two alternating exchange buffers: one barrier per transform suffices
            t = symm_mem.empty(self.local_words, dtype=torch.int64, device=dev)
            h = symm_mem.rendezvous(t, group)
            bufs.append(t)
            hdls.append(h)
        self._peer = {"bufs": bufs, "hdls": hdls, "parity": 0,
                      "ptrs": [[int(p) for p in h.buffer_ptrs] for h in hdls]}

    def _next_peer(self):
        i = self._peer["parity"]
        self._peer["parity"] ^= 1
        return self._peer["bufs"][i], self._peer["hdls"][i], self._peer["ptrs"][i]

    def _pick_chunks(self, want):
        """Largest power of two <= want (default 4) that the tile shapes allow."""
        k = want or 4
        # w = n1 / (G*K) must hold a whole column tile and, for 3-pass plans, whole inner runs of pass 1
        tile_w = 1 << min(5, max(0, 13 - self.splits[0]))
        inner1 = 1 << sum(self.splits[2:]) if len(self.splits) > 2 else 1
        while k > 1 and (self.n1 // (self.world * k) < max(tile_w, inner1) or self.n1 % (self.world * k)):
            k //= 2
        return max(k, 1)

    def _scratch(self, like):
        if self._tmp is None or self._tmp[0].device != like.device or self._tmp[0].numel() != like.numel():
            self._tmp = (torch.empty_like(like), torch.empty_like(like))
        return self._tmp

    # ---- pipelined path --------------------------------------------------------------------------
    def forward(self, dst, src, stream=0):
        """src: this rank's column block [n0][n1/G]; dst: this rank's row block [n0/G][n1]."""
        scope, stream = _stream_scope(stream)
        with scope:
            return self._forward(dst, src, stream)

    def _forward(self, dst, src, stream):
        if self.mode == "simple":
            return self._forward_simple(dst, src, stream)
        if self.mode == "peer":
            buf, hdl, ptrs = self._next_peer()
            self.plan.shard_forward_cols_peer(ptrs, src.data_ptr(), stream)
            hdl.barrier(channel=0)  # every rank's tiles have landed in everybody's buffer
            self.plan.shard_forward_rows_tiled(dst.data_ptr(), buf.data_ptr(), 1, stream)
            return
        K = self.chunks
        send, recv = self._scratch(src)
        sv, rv = send.view(K, -1), recv.view(K, -1)
        works = []
        for c in range(K):
            self.plan.shard_forward_cols_chunk(send.data_ptr(), src.data_ptr(), c, K, stream)
            works.append(dist.all_to_all_single(rv[c], sv[c], group=self.group, async_op=True))
        for w in works:
            w.wait()
        self.plan.shard_forward_rows_tiled(dst.data_ptr(), recv.data_ptr(), K, stream)

    def inverse(self, dst, src, stream=0):
        """src: row block [n0/G][n1] (bit-reversed order); dst: column block [n0][n1/G]."""
        scope, stream = _stream_scope(stream)
        with scope:
            return self._inverse(dst, src, stream)

    def _inverse(self, dst, src, stream):
        if self.mode == "simple":
            return self._inverse_simple(dst, src, stream)
        if self.mode == "peer":
            buf, hdl, ptrs = self._next_peer()
            work = self._scratch(src)[0]
            self.plan.shard_inverse_rows_peer(ptrs, src.data_ptr(), work.data_ptr(), stream)
            hdl.barrier(channel=0)
            self.plan.shard_inverse_cols_chunk(dst.data_ptr(), buf.data_ptr(), 0, 1, stream)
            return
        K = self.chunks
        send, recv = self._scratch(src)
        # the row half leaves its result tiled in `send`; `recv` doubles as its scratch
        self.plan.shard_inverse_rows_tiled(send.data_ptr(), src.data_ptr(), recv.data_ptr(), K, stream)
        sv, rv = send.view(K, -1), recv.view(K, -1)
        works = [dist.all_to_all_single(rv[c], sv[c], group=self.group, async_op=True) for c in range(K)]
        for c in range(K):
            works[c].wait()
            self.plan.shard_inverse_cols_chunk(dst.data_ptr(), recv.data_ptr(), c, K, stream)

    # ---- reference path: whole-block exchange + explicit tile interleave (any modulus) -------------
    def _forward_simple(self, dst, src, stream=0):
        G, n0, n1 = self.world, self.n0, self.n1
        send, recv = self._scratch(src)
        self.plan.shard_forward_cols(send.data_ptr(), src.data_ptr(), stream)
        dist.all_to_all_single(recv, send, group=self.group)
        dst.view(n0 // G, G, n1 // G).copy_(recv.view(G, n0 // G, n1 // G).permute(1, 0, 2))
        self.plan.shard_forward_rows(dst.data_ptr(), dst.data_ptr(), stream)

    def _inverse_simple(self, dst, src, stream=0):
        G, n0, n1 = self.world, self.n0, self.n1
        send, recv = self._scratch(src)
        self.plan.shard_inverse_rows(recv.data_ptr(), src.data_ptr(), stream)
        send.view(G, n0 // G, n1 // G).copy_(recv.view(n0 // G, G, n1 // G).permute(1, 0, 2))
        dist.all_to_all_single(recv, send, group=self.group)
        self.plan.shard_inverse_cols(dst.data_ptr(), recv.data_ptr(), stream)

    # ---- host buffers: what a caller without device-resident data runs (bench.py's e2e figure) --------------
    def forward_host(self, dst_host, src_host, stream=0):
        """src_host: this rank's column block in (pinned) host memory; dst_host: its slice of the spectrum.  H2D copy,
        sharded forward, D2H copy, synchronise."""
        dev_src, dev_dst = self._host_staging(src_host)
        scope, stream = _stream_scope(stream)
        with scope:
            dev_src.copy_(src_host, non_blocking=True)
            self._forward(dev_dst, dev_src, stream)
            dst_host.copy_(dev_dst, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def inverse_host(self, dst_host, src_host, stream=0):
        dev_src, dev_dst = self._host_staging(src_host)
        scope, stream = _stream_scope(stream)
        with scope:
            dev_src.copy_(src_host, non_blocking=True)
            self._inverse(dev_dst, dev_src, stream)
            dst_host.copy_(dev_dst, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def _host_staging(self, like):
        if getattr(self, "_hstage", None) is None:
            dev = torch.device("cuda", torch.cuda.current_device())
            self._hstage = (torch.empty(self.local_words, dtype=like.dtype, device=dev),
                            torch.empty(self.local_words, dtype=like.dtype, device=dev))
        return self._hstage

    def close(self):
        self.plan.close()
