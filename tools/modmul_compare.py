# SPDX-License-Identifier: Apache-2.0
"""Development probe: Montgomery (PAdic64) vs Shoup (FixedPoint64) kernels at the reference's 62-bit test prime, next to
the production prime's baked-in Montgomery kernels - forward / inverse device times of the same plans."""
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


def timeit(fn, reps=30, warm=5):
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


out = []
for L, batch in [(24, 1), (20, 64), (13, 2048), (26, 1)]:
    row = {"log2_m": L, "batch": batch}
    for name, kw in [("production_montgomery_static", {}),
                     ("p62_montgomery_runtime", {"modulus": 0x3A00000000000001, "generator": 3}),
                     ("p62_shoup_runtime", {"modulus": 0x3A00000000000001, "generator": 3, "fixed_point": True}),
                     ("p50_shoup_runtime", {"modulus": 0x0003F00000000001, "generator": 11, "fixed_point": True}),
                     ("goldilocks_montgomery_static", {"modulus": 0xFFFFFFFF00000001, "generator": 7}),
                     ("p64_other_montgomery_runtime", {"modulus": 0xFFFFFFFF70000001, "generator": 3})]:
        plan = lib.plan(L, batch=batch, **kw)
        n = batch << L
        src = torch.randint(0, 2**49, (n,), dtype=torch.int64, device="cuda")
        dst = torch.empty_like(src)
        f = timeit(lambda: plan.forward(dst.data_ptr(), src.data_ptr(), st))
        i = timeit(lambda: plan.inverse(dst.data_ptr(), src.data_ptr(), st))
        plan.inverse(dst.data_ptr(), src.data_ptr(), st)
        row[name] = {"modmul": plan.modmul, "forward_us": round(f, 1), "inverse_us": round(i, 1),
                     "forward_gelem_s": round(n / f / 1e3, 1)}
        plan.close()
    print(json.dumps(row), flush=True)
    out.append(row)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "modmul_compare.json"), "w"), indent=1)
