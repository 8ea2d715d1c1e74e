# SPDX-License-Identifier: Apache-2.0
"""Latency of xntt_forward_host / xntt_inverse_host on page-locked host buffers for small transforms (host clock around
the blocking calls).  python tools/host_small.py"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import oracle_lib  # noqa: E402

pkg = ge.load_package()
lib = pkg.Library(sys.argv[1]) if len(sys.argv) > 1 else pkg.load()
orc = oracle_lib.Oracle()
P0, G0 = pkg.P0, pkg.G0
torch.cuda.init()
for L, batch in [(10, 1), (13, 1), (15, 1), (17, 1), (18, 1), (19, 1), (20, 1), (21, 1), (12, 64)]:
    m = (1 << L) * batch
    plan = lib.plan(L, batch=batch)
    a = torch.from_numpy(orc.fill_xorshift(m, 3 + L, P0).view(np.int64)).pin_memory()
    b = torch.empty(m, dtype=torch.int64).pin_memory()
    c = torch.empty(m, dtype=torch.int64).pin_memory()
    plan.forward_host(b.data_ptr(), a.data_ptr())
    ok = True
    if L <= 17:
        want = np.concatenate([orc.ntt_forward(a.numpy().view(np.uint64)[i << L:(i + 1) << L].copy(), P0, G0) for i in range(batch)])
        ok = bool(np.array_equal(b.numpy().view(np.uint64), want))
    plan.inverse_host(c.data_ptr(), b.data_ptr())
    ok = ok and bool(torch.equal(a, c))
    reps = 200
    res = {}
    for name, fn in (("forward", lambda: plan.forward_host(b.data_ptr(), a.data_ptr())),
                     ("inverse", lambda: plan.inverse_host(c.data_ptr(), b.data_ptr()))):
        for _ in range(10):
            fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        res[name + "_us"] = round((time.perf_counter() - t0) / reps * 1e6, 1)
    print(json.dumps({"log2_m": L, "batch": batch, "bytes": m * 8, "ok": ok, **res}), flush=True)
    plan.close()
