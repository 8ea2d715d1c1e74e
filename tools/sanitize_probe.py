# SPDX-License-Identifier: Apache-2.0
"""Small transforms of every kernel shape, for compute-sanitizer (memcheck / racecheck) runs:
   compute-sanitizer --tool racecheck python tools/sanitize_probe.py"""
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
lib = pkg.load()
orc = oracle_lib.Oracle()
P0, G0 = pkg.P0, pkg.G0
st = torch.cuda.current_stream().cuda_stream
cases = [(13, None, 1, False), (12, None, 2, False), (11, None, 3, False), (10, None, 9, False), (9, None, 17, False),
         (8, None, 33, False), (6, None, 70, False), (17, [8, 9], 1, False), (17, [8, 9], 1, True), (14, [11, 3], 1, False),
         (15, [12, 3], 1, True), (16, [10, 6], 1, False), (18, [6, 6, 6], 1, False), (18, [6, 6, 6], 1, True),
         (16, [7, 9], 2, False), (16, [9, 7], 1, False), (24, None, 1, False)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
bad = 0
# small plans take narrow tiles by default: every case runs on both tile shapes
cases = [c + (t,) for c in cases for t in ((None, "wide") if c[0] <= 20 else (None,))]
for L, splits, batch, compact, tiles in cases:
    m = 1 << L
    a = orc.fill_xorshift(m * batch, 99 + L, P0)
    plan = lib.plan(L, splits=splits, batch=batch, compact_tables=compact, tiles=tiles)
    d = torch.from_numpy(a.view(np.int64)).cuda()
    o = torch.empty_like(d)
    plan.forward(o.data_ptr(), d.data_ptr(), st)
    b = torch.empty_like(d)
    plan.inverse(b.data_ptr(), o.data_ptr(), st)
    torch.cuda.synchronize()
    ok = bool(np.array_equal(b.cpu().numpy().view(np.uint64), a))
    if L <= 18:
        ok = ok and bool(np.array_equal(o.cpu().numpy().view(np.uint64)[:m], orc.ntt_forward(a[:m].copy(), P0, G0)))
    # fused point-wise product path
    plan.forward_multiply(o.data_ptr(), d.data_ptr(), d.data_ptr(), st)
    torch.cuda.synchronize()
    print(L, plan.splits, plan.tile_log2, batch, "compact" if compact else "matrix", "ok" if ok else "MISMATCH", flush=True)
    bad += 0 if ok else 1
    plan.close()
# runtime-modulus kernels (Montgomery and Shoup) and the multi-GPU plan with every rank on this device (peer stores,
# address-mapped passes)
for N, g, fixed in [(0xFFFFFFFF00000001, 7, False), (0x3A00000000000001, 3, True)]:
    for L, splits in [(13, None), (16, [5, 5, 6])]:
        a = orc.fill_xorshift(1 << L, 5 + L, N)
        plan = lib.plan(L, modulus=N, generator=g, splits=splits, fixed_point=fixed)
        d = torch.from_numpy(a.view(np.int64)).cuda()
        o = torch.empty_like(d)
        plan.forward(o.data_ptr(), d.data_ptr(), st)
        torch.cuda.synchronize()
        ok = bool(np.array_equal(o.cpu().numpy().view(np.uint64), orc.ntt_forward(a, N, g)))
        print(hex(N), L, "shoup" if plan.modmul else "montgomery", "ok" if ok else "MISMATCH", flush=True)
        bad += 0 if ok else 1
        plan.close()
for L, G, splits, N, g in [(16, 4, None, P0, G0), (15, 8, [3, 12], P0, G0), (18, 2, [6, 6, 6], P0, G0),
                           (16, 4, None, 0x3A00000000000001, 3)]:
    a = orc.fill_xorshift(1 << L, 77 + L, N)
    mg = lib.mgpu(L, [0] * G, splits=splits, modulus=N, generator=g)
    out, back = np.empty_like(a), np.empty_like(a)
    mg.forward_host(out.ctypes.data, a.ctypes.data)
    mg.inverse_host(back.ctypes.data, out.ctypes.data)
    ok = bool(np.array_equal(out, orc.ntt_forward(a, N, g))) and bool(np.array_equal(back, a))
    print("mgpu", L, G, splits, hex(N), "ok" if ok else "MISMATCH", flush=True)
    bad += 0 if ok else 1
    mg.close()
# in-place and out-of-place transposition
q = torch.arange(200 * 200, dtype=torch.int64, device="cuda").view(200, 200).clone()
want = q.t().contiguous()
lib.transpose(q.data_ptr(), q.data_ptr(), 200, 200, 200, 200, st)
torch.cuda.synchronize()
bad += 0 if torch.equal(q, want) else 1
# stand-alone kernels
t = torch.empty((200, 130), dtype=torch.int64, device="cuda")
s = torch.arange(130 * 200, dtype=torch.int64, device="cuda").view(130, 200)
lib.transpose(t.data_ptr(), s.data_ptr(), 130, 200, 130, 200, st)
torch.cuda.synchronize()
bad += 0 if torch.equal(t, s.t()) else 1
print("kinnaes", hex(lib.kinnaes_sum(0xFFFFFFFFFECA467F, 5, 100, 495017, 0, 600)))
print("FAILED" if bad else "ALL OK")
sys.exit(1 if bad else 0)
