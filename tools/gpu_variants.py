# SPDX-License-Identifier: Apache-2.0
"""Development probe: time tuning variants of libxntt (sve-ntt_b200/lib_*/libxntt.so) side by side."""
import glob
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import oracle_lib  # noqa: E402

pkg = ge.load_package()
orc = oracle_lib.Oracle()
P0, G0 = pkg.P0, pkg.G0
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
    return ts[0], ts[len(ts) // 2]


libs = sorted(glob.glob(os.path.join(ROOT, "sve-ntt_b200", "lib*", "libxntt.so")))
cases = [(24, None, 1), (13, None, 2048), (26, None, 1)] + \
        [tuple(json.loads(a)) for a in sys.argv[1:]]
res = {}
gen = torch.Generator(device="cuda")
gen.manual_seed(7)
probe_in = torch.randint(0, 2**62, (1 << 24,), dtype=torch.int64, device="cuda", generator=gen)
probe_ref = {}
for path in libs:
    name = os.path.basename(os.path.dirname(path))
    lib = pkg.Library(path)
    # parity guard
    a = orc.fill_xorshift(1 << 16, 5, P0)
    plan = lib.plan(16)
    d = torch.from_numpy(a.view(np.int64)).cuda()
    o = torch.empty_like(d)
    plan.forward(o.data_ptr(), d.data_ptr(), st)
    ok = bool(np.array_equal(o.cpu().numpy().view(np.uint64), orc.ntt_forward(a, P0, G0)))
    plan.close()
    # the 2^24 plan (2^11 columns x 2^13 rows) of every variant against the product library, all words, both directions
    plan = lib.plan(24)
    f_out, i_out = torch.empty_like(probe_in), torch.empty_like(probe_in)
    plan.forward(f_out.data_ptr(), probe_in.data_ptr(), st)
    plan.inverse(i_out.data_ptr(), f_out.data_ptr(), st)
    torch.cuda.synchronize()
    plan.close()
    if name == "lib":
        probe_ref["f"] = f_out.clone()
    same = bool(torch.equal(f_out, probe_ref["f"])) if "f" in probe_ref else None
    res[name] = {"parity": ok, "forward_2p24_equals_product": same, "roundtrip_2p24": bool(torch.equal(i_out, probe_in))}
    print(name, res[name], flush=True)
    for L, splits, batch in cases:
        m = 1 << L
        try:
            plan = lib.plan(L, splits=splits, batch=batch)
        except pkg.XnttError as e:
            print(name, L, splits, "plan failed", e)
            continue
        src = torch.randint(0, 2**62, (m * batch,), dtype=torch.int64, device="cuda")
        dst = torch.empty_like(src)
        fb, _ = timeit(lambda: plan.forward(dst.data_ptr(), src.data_ptr(), st))
        ib, _ = timeit(lambda: plan.inverse(dst.data_ptr(), src.data_ptr(), st))
        key = f"2^{L}x{batch}{plan.splits}"
        per = []
        for i in range(len(plan.splits)):
            tf, _ = timeit(lambda: plan.run_pass(i, 0, dst.data_ptr(), src.data_ptr(), st), reps=15)
            ti, _ = timeit(lambda: plan.run_pass(i, 1, dst.data_ptr(), src.data_ptr(), st), reps=15)
            per.append((round(tf * 1e3, 1), round(ti * 1e3, 1)))
        print(f"{name:10s} {key:26s} per-pass (fwd, inv) us: {per}", flush=True)
        res[name][key] = [round(fb * 1e3, 1), round(ib * 1e3, 1), round(m * batch / fb / 1e6, 1)]
        print(f"{name:10s} parity={ok} {key:26s} fwd {fb*1e3:9.1f} us inv {ib*1e3:9.1f} us  {m*batch/fb/1e6:6.1f} Gelem/s fwd",
              flush=True)
        plan.close()
        del src, dst
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "variants.json"), "w"), indent=1)
