# SPDX-License-Identifier: Apache-2.0
"""Bandwidth of the stand-alone transposition (xntt_transpose; the reference's tests/bench-transpose.cpp sweeps the
same shapes): out of place and in-place square, bytes moved = 16 per element, against the measured HBM copy rate."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
lib = pkg.load()
st = torch.cuda.current_stream().cuda_stream
peak = 6559.7
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[0]


out = []
for rows, cols in [(8192, 8192), (4096, 16384), (16384, 16384), (2048, 8192)]:
    src = torch.arange(rows * cols, dtype=torch.int64, device="cuda").view(rows, cols)
    dst = torch.empty((cols, rows), dtype=torch.int64, device="cuda")
    ms = timeit(lambda: lib.transpose(dst.data_ptr(), src.data_ptr(), rows, cols, rows, cols, st))
    assert torch.equal(dst, src.t())
    out.append({"rows": rows, "cols": cols, "ms": ms, "GBps": 16.0 * rows * cols / ms / 1e6,
                "frac_of_copy_peak": 16.0 * rows * cols / ms / 1e6 / peak})
    del src, dst
for dim in (8192, 16384):
    a = torch.arange(dim * dim, dtype=torch.int64, device="cuda").view(dim, dim)
    want = a.t().contiguous()
    lib.transpose(a.data_ptr(), a.data_ptr(), dim, dim, dim, dim, st)
    assert torch.equal(a, want)
    ms = timeit(lambda: lib.transpose(a.data_ptr(), a.data_ptr(), dim, dim, dim, dim, st))
    out.append({"inplace": dim, "ms": ms, "GBps": 16.0 * dim * dim / ms / 1e6,
                "frac_of_copy_peak": 16.0 * dim * dim / ms / 1e6 / peak})
    del a, want
for row in out:
    print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "transpose_bw.json"), "w"), indent=1)
