# SPDX-License-Identifier: Apache-2.0
"""The largest transform the C ABI accepts (log2_m = 31, 16 GiB per buffer) on one B200: directly evaluated output words
of a sparse input, and the round trip.  One-off check (not part of the -m gpu suite: 32+ GiB of buffers)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
lib = pkg.load()
P0, G0 = pkg.P0, pkg.G0
L = int(sys.argv[1]) if len(sys.argv) > 1 else 31
m = 1 << L
st = torch.cuda.current_stream().cuda_stream
plan = lib.plan(L)
rng = np.random.default_rng(L)
pos = np.unique(rng.integers(0, m, 64, dtype=np.int64))
val = rng.integers(1, P0, pos.size, dtype=np.uint64)
sparse = torch.zeros(m, dtype=torch.int64, device="cuda")
sparse[torch.from_numpy(pos).cuda()] = torch.from_numpy(val.view(np.int64)).cuda()
out = torch.empty_like(sparse)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
plan.forward(out.data_ptr(), sparse.data_ptr(), st)
torch.cuda.synchronize()
e0.record()
plan.forward(out.data_ptr(), sparse.data_ptr(), st)
e1.record()
torch.cuda.synchronize()
fwd_ms = e0.elapsed_time(e1)
omega = pow(G0, (P0 - 1) >> L, P0)
idx = np.unique(np.concatenate([rng.integers(0, m, 1000, dtype=np.int64), [0, 1, m // 2, m - 1, m - 2]]))
got = out[torch.from_numpy(idx).cuda()].cpu().numpy().view(np.uint64)
wrong = 0
for i, g in zip(idx, got):
    k = int(f"{int(i):0{L}b}"[::-1], 2)
    want = 0
    for p_, v_ in zip(pos, val):
        want = (want + int(v_) * pow(omega, (k * int(p_)) % m, P0)) % P0
    wrong += int(g) != want
e0.record()
plan.inverse(out.data_ptr(), out.data_ptr(), st)
e1.record()
torch.cuda.synchronize()
inv_ms = e0.elapsed_time(e1)
rec = {"log2_m": L, "splits": plan.splits, "twiddle_forms_fwd": plan.twiddle_forms(False), "twiddle_forms_inv": plan.twiddle_forms(True),
       "words_checked": int(idx.size), "words_wrong": wrong, "roundtrip_equal": bool(torch.equal(out, sparse)),
       "forward_ms": fwd_ms, "inverse_ms": inv_ms, "forward_gelem_s": m / fwd_ms / 1e6}
print(json.dumps(rec))
