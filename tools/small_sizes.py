# SPDX-License-Identifier: Apache-2.0
"""Latency of ONE transform of 2^10 .. 2^22 residues on one GPU: whole tiles against narrow tiles
(XNTT_TILES_WIDE / the planner's default).  CUDA events around back-to-back forward calls on one stream (every call
depends on the one before it: dst of one is src of the next), so the figure is the time one transform occupies the
GPU.  python tools/small_sizes.py [--out file.json]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def time_plan(plan, m, batch, reps):
    a = torch.randint(0, 2**62, (m * batch,), dtype=torch.int64, device="cuda")
    b = torch.empty_like(a)
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for name, fn in (("forward", plan.forward), ("inverse", plan.inverse)):
        for _ in range(5):
            fn(b.data_ptr(), a.data_ptr(), st)
            fn(a.data_ptr(), b.data_ptr(), st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(3):
            e0.record()
            for _ in range(reps // 2):
                fn(b.data_ptr(), a.data_ptr(), st)
                fn(a.data_ptr(), b.data_ptr(), st)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / (reps // 2 * 2))
        out[name + "_us"] = round(best, 2)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--reps", type=int, default=200)
    a = ap.parse_args()
    pkg = ge.load_package()
    lib = pkg.load()
    recs = []
    for L, batch in [(10, 1), (12, 1), (13, 1), (14, 1), (15, 1), (16, 1), (17, 1), (18, 1), (19, 1), (20, 1), (21, 1), (22, 1),
                     (15, 8), (17, 8), (12, 256)]:
        m = 1 << L
        rec = {"log2_m": L, "batch": batch}
        for tiles in ("wide", None):
            plan = lib.plan(L, batch=batch, tiles=tiles)
            key = "wide" if tiles else "default"
            rec[key] = {"splits": plan.splits, "tile_log2": plan.tile_log2, **time_plan(plan, m, batch, a.reps)}
            plan.close()
        rec["forward_speedup"] = round(rec["wide"]["forward_us"] / rec["default"]["forward_us"], 2)
        rec["default_forward_gelem_s"] = round(m * batch / rec["default"]["forward_us"] / 1e3, 2)
        recs.append(rec)
        print(json.dumps(rec), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(recs, f, indent=1)


if __name__ == "__main__":
    main()
