# SPDX-License-Identifier: Apache-2.0
"""torchrun probe: time the phases of the sharded forward separately (column chunks, all-to-all, row half)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sve-ntt_b200"))
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
import dist_ntt  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = pkg.load()
st = torch.cuda.current_stream().cuda_stream
dev = torch.device("cuda", local)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 28
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sh = dist_ntt.ShardedNTT(lib, L, world, rank, device=local, chunks=K)
K = sh.chunks
n = (1 << L) // world
src = torch.randint(0, 2**62, (n,), dtype=torch.int64, device=dev)
dst = torch.empty_like(src)
send, recv = sh._scratch(src)
sv, rv = send.view(K, -1), recv.view(K, -1)


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def cols():
    for c in range(K):
        sh.plan.shard_forward_cols_chunk(send.data_ptr(), src.data_ptr(), c, K, st)


def a2a_sync():
    for c in range(K):
        dist.all_to_all_single(rv[c], sv[c])


def a2a_whole():
    dist.all_to_all_single(recv, send)


def rows():
    sh.plan.shard_forward_rows_tiled(dst.data_ptr(), recv.data_ptr(), K, st)


def full():
    sh.forward(dst, src, st)


def full_noverlap():
    cols()
    a2a_sync()
    rows()


res = {"L": L, "world": world, "K": K, "cols_ms": timed(cols), "a2a_chunks_ms": timed(a2a_sync), "a2a_whole_ms": timed(a2a_whole),
       "rows_ms": timed(rows), "full_ms": timed(full), "full_noverlap_ms": timed(full_noverlap)}
if rank == 0:
    gb = 8 * n * (world - 1) / world / 1e9
    res["a2a_GBps_per_gpu"] = gb / (res["a2a_whole_ms"] * 1e-3)
    print(res, flush=True)
dist.destroy_process_group()
