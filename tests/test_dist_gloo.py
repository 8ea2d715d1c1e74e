# SPDX-License-Identifier: Apache-2.0
"""The N > 1 path on CPU: world_size 2 and 4 over gloo, one process per rank, compute by the host
emulator of the kernel templates.  The gathered result must equal the oracle's transform."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P0, G0, SEED = 0xFFFFFC6E80000001, 3, 0x9E3779B97F4A7C15


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, log2_m, splits, modulus, emu_path, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "sve-ntt_b200"))
    import __graft_entry__ as ge
    import oracle_lib
    pkg = ge.load_package()
    import dist_ntt
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        lib = pkg.Library(emu_path)
        orc = oracle_lib.Oracle()
        m = 1 << log2_m
        N, g = (P0, G0) if modulus is None else modulus
        a = orc.fill_xorshift(m, SEED, N)
        sh = dist_ntt.ShardedNTT(lib, log2_m, world, rank, splits=splits, modulus=N, generator=g)
        n0, n1 = sh.n0, sh.n1
        block = np.ascontiguousarray(a.reshape(n0, n1)[:, rank * n1 // world:(rank + 1) * n1 // world])
        src = torch.from_numpy(block.view(np.int64).reshape(-1).copy())
        dst = torch.empty_like(src)
        sh.forward(dst, src)
        want = orc.ntt_forward(a, N, g)[rank * m // world:(rank + 1) * m // world]
        ok_f = bool(np.array_equal(dst.numpy().view(np.uint64), want))
        back = torch.empty_like(src)
        sh.inverse(back, dst)
        ok_i = bool(torch.equal(back, src))
        q.put((rank, ok_f, ok_i))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,log2_m,splits,modulus", [
    (2, 14, [7, 7], None), (4, 16, [6, 5, 5], None), (2, 15, None, None), (4, 18, [7, 11], None),
    (2, 14, [7, 7], (0x3A00000000000001, 3)),  # runtime-modulus address-mapped kernels
    (4, 16, [6, 5, 5], (0xFFFFFFFF00000001, 7)),
])
def test_sharded_transform_over_gloo(emu, world, log2_m, splits, modulus):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, log2_m, splits, modulus, emu.path, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert res == [(r, True, True) for r in range(world)]
