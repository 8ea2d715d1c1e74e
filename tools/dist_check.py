# SPDX-License-Identifier: Apache-2.0
"""torchrun script: sharded transform over N GPUs == single-GPU transform (memcmp), plus timings.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py [log2_m ...]"""
import json
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
sizes = [int(a) for a in sys.argv[1:]] or [20, 24, 26]
out = []
for L in sizes:
    m = 1 << L
    sh = dist_ntt.ShardedNTT(lib, L, world, rank, device=local, mode=os.environ.get("XNTT_DIST_MODE"))
    if rank == 0:
        print("mode", sh.mode, getattr(sh, "_peer_error", ""), flush=True)
    n0, n1 = sh.n0, sh.n1
    # every rank builds the same full input, keeps its column block
    gen = torch.Generator(device=dev)
    gen.manual_seed(1000 + L)
    check = True  # every size: a 2^30 input plus its single-GPU transform is 16 GiB of the 180 GB per GPU
    if check:
        full = torch.randint(0, 2**62, (m,), dtype=torch.int64, device=dev, generator=gen)
        src = full.view(n0, n1)[:, rank * n1 // world:(rank + 1) * n1 // world].contiguous().view(-1)
    else:
        src = torch.randint(0, 2**62, (m // world,), dtype=torch.int64, device=dev, generator=gen)
    dst = torch.empty_like(src)
    sh.forward(dst, src, st)
    ok_f = ok_i = None
    if check:
        ref_plan = lib.plan(L, device=local)
        want = torch.empty_like(full)
        ref_plan.forward(want.data_ptr(), full.data_ptr(), st)
        ok_f = bool(torch.equal(dst, want[rank * m // world:(rank + 1) * m // world]))
        ref_plan.close()
        del want, full
    back = torch.empty_like(src)
    sh.inverse(back, dst, st)
    ok_i = bool(torch.equal(back, src))
    # timing
    for _ in range(3):
        sh.forward(dst, src, st)
        sh.inverse(back, dst, st)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        sh.forward(dst, src, st)
    e1.record()
    torch.cuda.synchronize()
    tf = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    e0.record()
    for _ in range(reps):
        sh.inverse(back, dst, st)
    e1.record()
    torch.cuda.synchronize()
    ti = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(tf, op=dist.ReduceOp.MAX)
    dist.all_reduce(ti, op=dist.ReduceOp.MAX)
    flags = torch.tensor([int(bool(ok_f) or ok_f is None), int(ok_i)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        rec = {"log2_m": L, "world": world, "mode": sh.mode, "splits": sh.splits, "forward_equal_single_gpu": bool(flags[0].item()) if check else None,
               "roundtrip": bool(flags[1].item()), "fwd_ms": float(tf.item()), "inv_ms": float(ti.item()),
               "fwd_gelem_s": m / float(tf.item()) / 1e6, "inv_gelem_s": m / float(ti.item()) / 1e6}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    sh.close()
    del src, dst, back
    torch.cuda.empty_cache()
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"dist_check_{world}.json"), "w") as fh:
        json.dump(out, fh, indent=1)
dist.destroy_process_group()
