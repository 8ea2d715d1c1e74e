# SPDX-License-Identifier: Apache-2.0
"""Development probe run on the GPU box: micro-benchmarks, parity sweep against the oracle and
first timings.  Writes gpurun_out/gpu_check.json.  Not part of the product or the test-suite."""
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
lib = pkg.load()
orc = oracle_lib.Oracle()
P0, G0 = pkg.P0, pkg.G0
out = {"version": lib.version(), "gpu": torch.cuda.get_device_name(0)}
torch.cuda.set_device(0)
st = torch.cuda.current_stream().cuda_stream


def dev(a):
    return torch.from_numpy(a.view(np.int64)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint64)


# ---- micro-benchmarks
mb = {}
for kind, name in [(0, "imad"), (1, "imad_wide"), (5, "imad_hi"), (2, "lop3"), (4, "imad+lop3"), (3, "butterfly")]:
    g, ms = lib.microbench(kind, 2000)
    mb[name] = {"gops": g, "ms": ms}
    print("microbench", name, f"{g:.1f} Gop/s", f"{ms:.3f} ms", flush=True)
out["microbench"] = mb

# ---- parity sweep
par = []
ok_all = True
cases = [(L, None, 1) for L in range(1, 21)] + [(17, [8, 9], 1), (13, [9, 4], 1), (15, [9, 6], 2), (12, None, 5),
                                              (18, [6, 6, 6], 1), (10, None, 37), (3, None, 1000), (22, None, 1)]
for L, splits, batch in cases:
    m = 1 << L
    a = orc.fill_xorshift(m * batch, 0x9E3779B97F4A7C15 + L, P0)
    plan = lib.plan(L, splits=splits, batch=batch)
    src = dev(a)
    dst = torch.empty_like(src)
    plan.forward(dst.data_ptr(), src.data_ptr(), st)
    got = host(dst)
    ok = True
    for b in range(min(batch, 3)):
        want = orc.ntt_forward(a[b * m:(b + 1) * m].copy(), P0, G0)
        if not np.array_equal(got[b * m:(b + 1) * m], want):
            ok = False
            bad = np.flatnonzero(got[b * m:(b + 1) * m] != want)
            print(f"  FWD MISMATCH L={L} splits={plan.splits} batch={b}: {bad.size} words, first {bad[:4]}")
            break
    back = torch.empty_like(src)
    plan.inverse(back.data_ptr(), dst.data_ptr(), st)
    rt = bool(np.array_equal(host(back), a))
    canon = bool((got < np.uint64(P0)).all())
    par.append({"log2_m": L, "splits": plan.splits, "batch": batch, "forward": ok, "roundtrip": rt, "canonical": canon})
    ok_all &= ok and rt and canon
    print(f"parity L={L} splits={plan.splits} batch={batch}: fwd={ok} roundtrip={rt} canonical={canon}", flush=True)
    plan.close()
out["parity"] = par
out["parity_all"] = ok_all

# ---- timings
def timeit(fn, reps=20, warm=5):
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
    return ts[0], ts[len(ts) // 2]


tim = []
for L, splits, batch in [(24, None, 1), (24, [12, 12], 1), (24, [10, 14 - 1], 1) if False else (24, [11, 13], 1),
                         (20, None, 256), (20, [10, 10], 256), (22, None, 1), (26, None, 1), (28, None, 1), (12, None, 4096),
                         (24, [8, 8, 8], 1), (24, [10, 12 + 2 - 2 + 2], 1) if False else (24, [12, 12], 4)]:
    m = 1 << L
    try:
        plan = lib.plan(L, splits=splits, batch=batch)
    except pkg.XnttError as e:
        print("plan failed", L, splits, e)
        continue
    src = torch.randint(0, 2**62, (m * batch,), dtype=torch.int64, device="cuda")
    dst = torch.empty_like(src)
    fb, fm = timeit(lambda: plan.forward(dst.data_ptr(), src.data_ptr(), st))
    ib, im = timeit(lambda: plan.inverse(dst.data_ptr(), src.data_ptr(), st))
    ge_f = m * batch / (fb * 1e-3) / 1e9
    ge_i = m * batch / (ib * 1e-3) / 1e9
    tim.append({"log2_m": L, "splits": plan.splits, "batch": batch, "fwd_ms_best": fb, "fwd_ms_med": fm,
                "inv_ms_best": ib, "inv_ms_med": im, "fwd_gelem_s": ge_f, "inv_gelem_s": ge_i})
    print(f"time L={L} splits={plan.splits} batch={batch}: fwd {fb:.3f} ms ({ge_f:.1f} Gelem/s) inv {ib:.3f} ms ({ge_i:.1f} Gelem/s)",
          flush=True)
    plan.close()
    del src, dst
out["timings"] = tim
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
    json.dump(out, f, indent=1)
print("PARITY_ALL", ok_all)
sys.exit(0 if ok_all else 1)
