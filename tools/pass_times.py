# SPDX-License-Identifier: Apache-2.0
"""Development probe: per-pass device times of a plan (xntt_run_pass), HBM-streaming and L2-resident sizes."""
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


def timeit(fn, reps=40, warm=5):
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
    return ts[0] * 1e3


cases = [(24, [11, 13], 1), (24, [12, 12], 1), (20, [7, 13], 1), (20, [7, 13], 16), (20, [8, 12], 16), (13, None, 128), (13, None, 2048),
         (12, None, 256), (12, None, 4096), (11, None, 8192), (22, [11, 11], 4), (22, [9, 13], 4)]
cases += [tuple(json.loads(a)) for a in sys.argv[1:]]
out = []
for L, splits, batch in cases:
    for compact in (False, True):
        plan = lib.plan(L, splits=splits, batch=batch, compact_tables=compact)
        n = (1 << L) * batch
        src = torch.randint(0, 2**62, (n,), dtype=torch.int64, device="cuda")
        dst = torch.empty_like(src)
        row = {"L": L, "splits": plan.splits, "batch": batch, "compact": compact, "MB": n * 8 / 1e6}
        for i, ln in enumerate(plan.splits):
            for inv in (0, 1):
                t = timeit(lambda: plan.run_pass(i, inv, dst.data_ptr(), src.data_ptr(), st))
                row[f"p{i}_{'inv' if inv else 'fwd'}_us"] = round(t, 1)
                # ns per element-level (per butterfly level and element)
                row[f"p{i}_{'inv' if inv else 'fwd'}_ps_per_elem_level"] = round(t * 1e6 / n / ln, 3)
        print(json.dumps(row), flush=True)
        out.append(row)
        plan.close()
        if len(plan.splits) == 1:
            break
        del src, dst
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "pass_times.json"), "w"), indent=1)
