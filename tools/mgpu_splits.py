# SPDX-License-Identifier: Apache-2.0
"""One process, G GPUs (xntt_mgpu_*): time ONE sharded transform for several decompositions and check that they all
produce the same words.  python tools/mgpu_splits.py --devices 0,1,2,3,4,5,6,7 --log2m 30 --splits 9,9,12 10,10,10
With --devices 0,0 every rank shares one GPU (dry run of the script on a single-GPU box)."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def fnv(t):
    """order-dependent 64-bit digest of an int64 tensor (on its device)"""
    idx = torch.arange(t.numel(), dtype=torch.int64, device=t.device)
    return int(((t ^ (idx * -7046029254386353131)) * 1099511628211).sum().item()) & 0xFFFFFFFFFFFFFFFF


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", default="0,1")
    ap.add_argument("--log2m", type=int, default=30)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--splits", nargs="*", default=["default"])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    pkg = ge.load_package()
    lib = pkg.load()
    devs = [int(x) for x in a.devices.split(",")]
    G, L = len(devs), a.log2m
    m = 1 << L
    recs, digests0 = [], None
    for spec in a.splits:
        splits = None if spec == "default" else [int(x) for x in spec.split(",")]
        rec = {"log2_m": L, "devices": devs, "splits_asked": spec}
        try:
            mg = lib.mgpu(L, devs, splits=splits)
        except Exception as e:  # unsupported decomposition
            rec["error"] = str(e)[:200]
            recs.append(rec)
            print(json.dumps(rec), flush=True)
            continue
        n0, n1 = mg.n0, mg.n1
        rec["n0_log2"] = n0.bit_length() - 1
        # rank r's column block of ONE global input x[i] = hash(i) (i = k * n1 + column): whatever the decomposition, the
        # spectrum slices must come out the same
        src, dst, back = [], [], []
        w = n1 // G
        for r, d in enumerate(devs):
            dv = f"cuda:{d}"
            with torch.cuda.device(d):
                idx = (torch.arange(n0, dtype=torch.int64, device=dv)[:, None] * n1 + (r * w)
                       + torch.arange(w, dtype=torch.int64, device=dv)[None, :]).reshape(-1)
                src.append(((idx * -7046029254386353131 + 0x1234567) ^ (idx >> 7)) & ((1 << 62) - 1))
                del idx
                dst.append(torch.empty(m // G, dtype=torch.int64, device=dv))
                back.append(torch.empty(m // G, dtype=torch.int64, device=dv))
        sp = lambda ts: [t.data_ptr() for t in ts]  # noqa: E731
        for d in set(devs):
            torch.cuda.synchronize(d)
        mg.forward(sp(dst), sp(src))   # column blocks -> spectrum slices
        mg.inverse(sp(back), sp(dst))  # and back
        mg.synchronize()
        rec["roundtrip_ok"] = all(bool(torch.equal(back[r], src[r])) for r in range(G))
        rec["digest_spectrum"] = [fnv(dst[r]) for r in range(G)]
        if digests0 is None:
            digests0 = rec["digest_spectrum"]
        rec["spectrum_equals_first_run"] = rec["digest_spectrum"] == digests0
        for _ in range(2):
            mg.forward(sp(dst), sp(src))
            mg.inverse(sp(back), sp(dst))
        mg.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            mg.forward(sp(dst), sp(src))
        mg.synchronize()
        t1 = time.perf_counter()
        for _ in range(a.reps):
            mg.inverse(sp(back), sp(dst))
        mg.synchronize()
        t2 = time.perf_counter()
        rec["forward_ms"] = (t1 - t0) * 1e3 / a.reps
        rec["inverse_ms"] = (t2 - t1) * 1e3 / a.reps
        rec["roundtrip_gelem_s"] = 2 * m / ((t2 - t0) / a.reps) / 1e9
        rec["timing"] = "host clock around xntt_mgpu_synchronize, %d back-to-back transforms per direction" % a.reps
        mg.close()
        del src, dst, back
        for d in set(devs):
            with torch.cuda.device(d):
                torch.cuda.empty_cache()
        recs.append(rec)
        print(json.dumps(rec), flush=True)
    summary = {"all_spectra_equal": all(r.get("spectrum_equals_first_run", True) for r in recs),
               "all_roundtrips_ok": all(r.get("roundtrip_ok", True) for r in recs)}
    print(json.dumps(summary), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump({"runs": recs, "summary": summary}, f, indent=1)


if __name__ == "__main__":
    main()
