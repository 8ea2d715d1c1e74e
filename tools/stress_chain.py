# SPDX-License-Identifier: Apache-2.0
"""Chains of dependent transforms, in place on one buffer, enqueued back to back (every pass a programmatic dependent of
the kernel before it) against the same chain with a device synchronisation after every call: the two must agree word for
word, and a forward/inverse chain must return its input.  Looks for ordering bugs between dependent launches that a
single transform cannot show.  python tools/stress_chain.py [--reps R]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=200)
    a = ap.parse_args()
    pkg = ge.load_package()
    lib = pkg.load()
    st = torch.cuda.current_stream().cuda_stream
    bad = 0
    for L, batch, tiles in [(8, 1, None), (10, 1, None), (12, 1, None), (13, 1, None), (13, 16, None), (14, 1, None), (16, 1, None),
                            (17, 1, None), (17, 8, None), (19, 1, None), (19, 1, "wide"), (20, 1, None), (21, 1, None), (22, 1, None),
                            (24, 1, None), (26, 1, None)]:
        m = 1 << L
        reps = a.reps if L <= 22 else max(8, a.reps // 10)
        plan = lib.plan(L, batch=batch, tiles=tiles)
        gen = torch.Generator(device="cuda")
        gen.manual_seed(L)
        x0 = torch.randint(0, 2**62, (m * batch,), dtype=torch.int64, device="cuda", generator=gen)
        # 1. forward chain, free-running
        x = x0.clone()
        for _ in range(reps):
            plan.forward(x.data_ptr(), x.data_ptr(), st)
        torch.cuda.synchronize()
        # 2. the same chain, synchronised after every call
        y = x0.clone()
        for _ in range(reps):
            plan.forward(y.data_ptr(), y.data_ptr(), st)
            torch.cuda.synchronize()
        same = bool(torch.equal(x, y))
        # 3. forward / inverse ping-pong, free-running, must return the input
        z = x0.clone()
        for _ in range(reps):
            plan.forward(z.data_ptr(), z.data_ptr(), st)
            plan.inverse(z.data_ptr(), z.data_ptr(), st)
        torch.cuda.synchronize()
        back = bool(torch.equal(z, x0))
        # 4. out-of-place ping-pong between two buffers (the next call overwrites what the previous one read)
        p, q = x0.clone(), torch.empty_like(x0)
        for _ in range(reps):
            plan.forward(q.data_ptr(), p.data_ptr(), st)
            plan.inverse(p.data_ptr(), q.data_ptr(), st)
        torch.cuda.synchronize()
        back2 = bool(torch.equal(p, x0))
        rec = {"log2_m": L, "batch": batch, "tiles": tiles or "default", "splits": plan.splits, "tile_log2": plan.tile_log2, "reps": reps,
               "chain_equals_synchronised": same, "pingpong_in_place": back, "pingpong_two_buffers": back2}
        print(json.dumps(rec), flush=True)
        bad += (not same) + (not back) + (not back2)
        plan.close()
    print(json.dumps({"failures": bad}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
