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

torch.distributed is plumbing here (NCCL all_to_all_single over NVLink on the box, gloo in the CPU
tests); the compute is libxntt's xntt_shard_* entry points.
"""
import torch
import torch.distributed as dist


class ShardedNTT:
    def __init__(self, library, log2_m, world, rank, device=-1, splits=None, group=None, inverse_factor=None):
        self.world, self.rank, self.group = world, rank, group
        self.plan = library.plan(log2_m, splits=splits, shard_count=world, shard_rank=rank, device=device,
                                 inverse_factor=inverse_factor)
        self.splits = self.plan.splits
        self.n0 = 1 << self.splits[0]
        self.n1 = (1 << log2_m) // self.n0
        self.local_words = (1 << log2_m) // world
        self.extra_launches_per_roundtrip = 0  # the tile (un)packing copies are torch's kernels, not ours
        self._tmp = None

    def _scratch(self, like):
        if self._tmp is None or self._tmp[0].device != like.device:
            self._tmp = (torch.empty_like(like), torch.empty_like(like))
        return self._tmp

    def forward(self, dst, src, stream=0):
        """src: this rank's column block [n0][n1/G]; dst: this rank's row block [n0/G][n1]."""
        G, n0, n1 = self.world, self.n0, self.n1
        send, recv = self._scratch(src)
        self.plan.shard_forward_cols(send.data_ptr(), src.data_ptr(), stream)
        # chunk s of `send` = rows [s*n0/G, (s+1)*n0/G) of my columns -> rank s
        dist.all_to_all_single(recv, send, group=self.group)
        # recv[s] = my rows x rank s's columns: interleave the G column blocks into whole rows
        dst.view(n0 // G, G, n1 // G).copy_(recv.view(G, n0 // G, n1 // G).permute(1, 0, 2))
        self.plan.shard_forward_rows(dst.data_ptr(), dst.data_ptr(), stream)

    def inverse(self, dst, src, stream=0):
        """src: row block [n0/G][n1] (bit-reversed order); dst: column block [n0][n1/G]."""
        G, n0, n1 = self.world, self.n0, self.n1
        send, recv = self._scratch(src)
        self.plan.shard_inverse_rows(recv.data_ptr(), src.data_ptr(), stream)
        send.view(G, n0 // G, n1 // G).copy_(recv.view(n0 // G, G, n1 // G).permute(1, 0, 2))
        dist.all_to_all_single(recv, send, group=self.group)
        # recv[s] = rows of rank s x my columns, already in [n0][n1/G] order
        self.plan.shard_inverse_cols(dst.data_ptr(), recv.data_ptr(), stream)

    def close(self):
        self.plan.close()
