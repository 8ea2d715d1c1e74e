# SPDX-License-Identifier: Apache-2.0
"""bench.py - 64-bit NTT throughput (Gelem/s) on B200, BASELINE.json's metric:
"64-bit NTT Gelem/s at 2^24 (1 GPU) and 2^30 (1/2/4/8 GPU); % of roofline".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

A "step" is one out-of-place compute_forward followed by one compute_inverse (a full round trip, the pair the
reference's tests/bench-ntt.cpp:47-58 times as "Forward, ..." / "Inverse, ...") over synthetic residues already
resident in HBM.  Workloads (default: ntt24 on one GPU, dist30 under torchrun with N > 1):

  ntt24     single blocked six-step transform, n = 2^24, p = 0xfffffc6e80000001, g = 3 (BASELINE.json configs[1]).
            The line also carries the 1-GPU figures of the other configs (dist30: the plain three-pass 2^30 plan;
            batch20: 256 x 2^20) as extra keys, and forward / inverse separately (fwd_inv).
  distNN    ONE 2^NN transform sharded over the N ranks (configs[3], strong scaling): six-step column/row split,
            the single exchange fused into the producing pass as peer stores over NVLink (NCCL-pipelined as the
            recorded fallback).  Before timing, every rank's slice of the sharded forward is compared with the
            single-GPU plan's transform of the same input (all words) and with directly evaluated DFT sums.
  batch20   256 x 2^20 batched transforms, the batch sharded across ranks (configs[2]), no collective.

value = elements transformed per second (2 transforms x n x batch per step) over all ranks, from CUDA events on the
launching stream around exactly K steps, max over ranks.  e2e = the same metric through the host-buffer API (pinned
host memory, both PCIe copies inside the timed region).  roofline = the binding roof of the dominant kernel: the
integer multiplier for the transforms (11 32-bit multiply instructions per modular product at the measured IMAD rate,
SURVEY.md section 8d), NVLink for the pass fused with the exchange at N > 1; roofline_hbm keeps the HBM fraction.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "sve-ntt_b200"))

P0, G0 = 0xFFFFFC6E80000001, 3
SEED = 0x9E3779B97F4A7C15
METRIC = "64-bit NTT throughput (forward+inverse round trip)"
NVLINK_PEAK_GBS, NVLINK_NOMINAL_GBS = 770.0, 900.0  # B200_PROFILING.md: measured peer copy / nominal, per direction
# ncu --set full capture of the default command, per launch (profiles/, archival: not measured in this run)
TRAFFIC_PROFILE = "profiles/r4_ncu_pass_kernels.json"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the dist30 / batch20 1-GPU extras of the ntt24 line")
    ap.add_argument("--mode", default=None, help="exchange of the sharded workload: peer (default) | pipelined | simple")
    return ap.parse_args()


def default_workload(world):
    return "ntt24" if world == 1 else "dist30"


def workload_name(w, world):
    if w == "ntt24":
        return "blocked six-step forward+inverse NTT n=2^24 uint64, p=0xfffffc6e80000001 g=3, one transform per GPU"
    if w == "batch20":
        return "batched 256x NTT n=2^20 forward+inverse, batch sharded across GPUs"
    if w.startswith("dist"):
        if world == 1:
            return f"forward+inverse NTT n=2^{int(w[4:])} on one GPU (three-pass plan)"
        return (f"distributed six-step forward+inverse NTT n=2^{int(w[4:])}: ONE transform sharded over {world} GPUs, "
                "one all-to-all per transform")
    return w


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def archived_traffic(workload, splits, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture
    of this command (archival: says which file); None when no capture covers the configuration."""
    path = os.path.join(ROOT, TRAFFIC_PROFILE)
    if workload != "ntt24" or not os.path.exists(path):
        return None, None
    with open(path) as fh:
        prof = json.load(fh)
    if list(prof.get("splits", [])) != list(splits):
        return None, None
    for k in prof.get("kernels", []):
        if k.get("name") == kernel:
            return float(k["dram_bytes"]), f"{TRAFFIC_PROFILE} (ncu --set full, archival)"
    return None, None


# ------------------------------------------------------------------------------------------------
# clocks: NVML sampled in a thread during the timed region
def warm_up(torch, step, min_steps, agree_max=None, min_seconds=0.4, max_steps=4000):
    """At least `min_steps` untimed steps AND about `min_seconds` of GPU work: a GPU that sat idle while the host
    checked results needs more than three short steps to be back at its clocks (observed on the 2-GPU sharded
    workload: 3 warm-up steps of 31 ms left the first timed steps 12 % slow).  Every rank runs the same number of steps
    (the sharded step synchronises the ranks): the count follows from the slowest rank's time for the first
    `min_steps`, agreed through `agree_max` (an all-reduce MAX)."""
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(min_steps):
        step(i)
    torch.cuda.synchronize()
    elapsed = time.perf_counter() - t0
    if agree_max is not None:
        elapsed = agree_max(elapsed)
    extra = 0
    if elapsed < min_seconds:
        extra = min(max_steps, int((min_seconds - elapsed) / max(elapsed / min_steps, 1e-6)) + 1)
    for i in range(extra):
        step(min_steps + i)
    torch.cuda.synchronize()
    return min_steps + extra


class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread:
            self._stop.set()
            self._thread.join()
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU baselines.  The only place besides the reference arm that executes oracle/.
def cpu_baselines(workload, budget_s=25.0):
    """B1: NTTReference (tests/ntt-reference.hpp, the real prime, serial).  B2: the reference's portable scalar kernel
    IterativeNTT<RadixEightScalarLayer<PAdic64Scalar>...> (62-bit test prime - its arithmetic is wrong above 2^62), one
    thread.  B3: B2 under OpenMP over a batch on all host cores.  Bounded samples; returns (headline, all)."""
    import oracle_lib
    cores = os.cpu_count() or 1
    res = []
    orc = oracle_lib.Oracle()
    if oracle_lib.have_reference():
        impl, kind = oracle_lib.Reference(), "reference"
    else:
        impl, kind = orc, "port"
    n1 = 1 << 22
    a = orc.fill_xorshift(n1, SEED, P0)
    t0 = time.perf_counter()
    f = impl.ntt_forward(a, P0, G0)
    t1 = time.perf_counter()
    back = impl.ntt_inverse(f, P0, G0)
    t2 = time.perf_counter()
    assert np.array_equal(back, a)
    res.append({"name": "B1", "value": 2.0 * n1 / (t2 - t0) / 1e9, "unit": "Gelem/s", "cores": 1, "kind": kind,
                "forward_gelem_s": n1 / (t1 - t0) / 1e9, "inverse_gelem_s": n1 / (t2 - t1) / 1e9,
                "sample": "one forward+inverse of NTTReference (tests/ntt-reference.hpp:43-83) at n=2^22, "
                          "p=0xfffffc6e80000001, 1 thread (the class is serial)"})
    if oracle_lib.have_reference_scalar():
        sc = oracle_lib.ReferenceScalar()
        L = 24
        a = orc.fill_xorshift(1 << L, SEED, sc.N)
        t0 = time.perf_counter()
        f = sc.run(L, False, a)
        t1 = time.perf_counter()
        back = sc.run(L, True, f)
        t2 = time.perf_counter()
        assert np.array_equal(back % np.uint64(sc.N), a)
        res.append({"name": "B2", "value": 2.0 * (1 << L) / (t2 - t0) / 1e9, "unit": "Gelem/s", "cores": 1,
                    "kind": "reference", "forward_gelem_s": (1 << L) / (t1 - t0) / 1e9,
                    "inverse_gelem_s": (1 << L) / (t2 - t1) / 1e9,
                    "sample": "one forward+inverse of the reference's scalar IterativeNTT<RadixEightScalarLayer<PAdic64Scalar> x8> "
                              "(layer/scalar/radix-eight.hpp) at n=2^24, 62-bit test prime 0x3a00000000000001 (the scalar "
                              "path is incorrect for p >= 2^62), 1 thread (the kernel is serial)"})
        # B3: the batched config on all cores, a bounded slice of the 256 transforms
        batch = max(cores, 8)
        a = orc.fill_xorshift(batch << 20, SEED + 1, sc.N)
        out = np.empty_like(a)
        sc.run(20, False, a[:1 << 20].copy(), 1, 1)
        t0 = time.perf_counter()
        sc.run(20, False, a, batch, cores, out=out)
        sc.run(20, True, out, batch, cores, out=out)
        dt = time.perf_counter() - t0
        assert np.array_equal(out % np.uint64(sc.N), a)
        res.append({"name": "B3", "value": 2.0 * (batch << 20) / dt / 1e9, "unit": "Gelem/s", "cores": cores,
                    "kind": "reference",
                    "sample": f"forward+inverse of {batch} x 2^20 transforms (a slice of the 256 x 2^20 batch), the scalar "
                              f"kernel of B2 under `omp parallel for` over the batch on {cores} host threads, 62-bit prime"})
    want = "B3" if workload == "batch20" else "B2"
    head = next((r for r in res if r["name"] == want), None) or max(res, key=lambda r: r["value"])
    head = {k: head[k] for k in ("value", "unit", "cores", "kind", "sample")}
    head["host_cores"] = cores
    return head, res


# ------------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path, on the box's host cores, rank 0 only.  Fastest correct
    reference code first: its scalar IterativeNTT kernel compiled from /root/reference (B2; B3 = OpenMP over the batch
    for the batched workload), else NTTReference, else the oracle port.  Every step is a bounded sample of the
    workload: one 2^24 round trip (the 2^30 transform of distNN would take ~10 minutes per step)."""
    if rank != 0:
        return
    import oracle_lib
    w = args.workload or default_workload(args.gpus)
    orc = oracle_lib.Oracle()
    cores = os.cpu_count() or 1
    env_log2 = os.environ.get("XNTT_BENCH_REF_LOG2")  # tests shrink the sample
    if oracle_lib.have_reference_scalar() and not (env_log2 and int(env_log2) not in oracle_lib.ReferenceScalar.SIZES):
        sc = oracle_lib.ReferenceScalar()
        N = sc.N
        if w == "batch20":
            log2_n, batch, threads = 20, max(cores, 8), cores
        else:
            log2_n, batch, threads = 24, 1, 1
        if env_log2:
            log2_n = int(env_log2)
        a = orc.fill_xorshift(batch << log2_n, SEED, N)
        mid, out = np.empty_like(a), np.empty_like(a)

        def one():
            sc.run(log2_n, False, a, batch, threads, out=mid)
            sc.run(log2_n, True, mid, batch, threads, out=out)

        kind = "reference"
        what = (f"the reference's scalar IterativeNTT<RadixEightScalarLayer<PAdic64Scalar>...> (include/sventt/layer/scalar/"
                f"radix-eight.hpp, compiled from the reference: oracle/_ref/libnttref_scalar.so), {batch} x n=2^{log2_n}, "
                f"modulus 0x3a00000000000001 (62-bit test prime: the scalar path is incorrect for p >= 2^62; the SVE path "
                f"cannot be built on x86), {threads} thread(s)" + (" via omp parallel for over the batch" if threads > 1 else
                                                                  " (the kernel is serial)"))
        check = lambda: np.array_equal(out % np.uint64(N), a)  # noqa: E731
    else:
        impl, kind = (oracle_lib.Reference(), "reference") if oracle_lib.have_reference() else (orc, "port")
        log2_n, batch, threads = int(env_log2 or 22), 1, 1
        a = orc.fill_xorshift(1 << log2_n, SEED, P0)
        box = {}

        def one():
            box["out"] = impl.ntt_inverse(impl.ntt_forward(a, P0, G0), P0, G0)

        what = f"NTTReference (tests/ntt-reference.hpp) n=2^{log2_n}, p=0xfffffc6e80000001, 1 thread (the class is serial)"
        check = lambda: np.array_equal(box["out"], a)  # noqa: E731
    for _ in range(args.warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t0
    one()
    assert check(), "reference arm: inverse(forward(x)) != x"
    value = 2.0 * (batch << log2_n) * args.steps / dt / 1e9
    same = (w == "ntt24" and log2_n == 24) or (w == "batch20" and log2_n == 20)
    sample = f"forward+inverse per step of {what}; host has {cores} cores"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gelem/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak" if w == "ntt24" else "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(w, args.gpus), "sample": sample, "log2_n": log2_n, "batch": batch,
                   "same_size_as_workload": same},
        "cpu_baseline": {"value": value, "unit": "Gelem/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Gelem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
def integer_roofline(lib, modmuls, seconds):
    """north_star's integer roof: 11 32-bit multiply-class instructions per modular product (mul.lo64 = 1 wide + 2
    narrow, two mul.hi64 = 4 wide each) at the peak IMAD rate; also the same count weighted with the measured
    IMAD.WIDE rate (a 32x32->64 product issues at well under half the IMAD rate on sm_100a)."""
    imad, _ = lib.microbench(0, 1000)
    wide, _ = lib.microbench(1, 1000)
    hi, _ = lib.microbench(5, 1000)
    bf, _ = lib.microbench(3, 1000)
    t_narrow = modmuls * 11.0 / (imad * 1e9)
    t_weighted = modmuls * (2.0 / (imad * 1e9) + 9.0 / (wide * 1e9))
    return {"achieved": modmuls * 11.0 / seconds / 1e12, "peak": imad / 1e3, "unit": "Tinstr/s",
            "frac": t_narrow / seconds, "frac_wide_weighted": t_weighted / seconds,
            "imad_gops": imad, "imad_wide_gops": wide, "imad_hi_gops": hi, "butterfly_gops": bf,
            "modmuls": modmuls,
            "definition": "achieved = 11 multiply instructions x (n/2) log2 n modular products / measured time; peak = "
                          "IMAD rate measured on this GPU by a register-resident loop (xntt_microbench, SASS of the loops: "
                          "profiles/r2_microbench_sass.txt); frac_wide_weighted charges the 9 32x32->64 products at the "
                          "measured IMAD.WIDE rate instead"}


def time_events(torch, stream, fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(reps):
        fn(i)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def make_sets(torch, dev, words, rank, max_sets=8):
    """Ring of (src, mid, out) buffer sets larger than L2 so that every step starts on data that is not cached."""
    set_bytes = 3 * 8 * words
    nsets = max(2, min(max_sets, int(1.5 * 2**30 // set_bytes))) if set_bytes < 2**30 else 1
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    sets = []
    for _ in range(nsets):
        src = torch.randint(0, 2**62, (words,), dtype=torch.int64, device=dev, generator=gen)
        sets.append((src, torch.empty_like(src), torch.empty_like(src)))
    return sets, nsets, set_bytes


def one_gpu_roundtrip(torch, lib, stream, dev, log2_m, batch, reps, rank=0):
    """forward / inverse / round-trip Gelem/s of a plain single-GPU plan (extras of the ntt24 line, and the
    same-size single-GPU figure next to a sharded run)."""
    plan = lib.plan(log2_m, batch=batch, device=dev.index)
    words = batch << log2_m
    sets, nsets, _ = make_sets(torch, dev, words, rank, max_sets=3)
    st = stream.cuda_stream
    src, mid, out = sets[0]
    plan.forward(mid.data_ptr(), src.data_ptr(), st)
    plan.inverse(out.data_ptr(), mid.data_ptr(), st)
    torch.cuda.synchronize()
    assert torch.equal(src, out), "inverse(forward(x)) != x"
    f = lambda i: plan.forward(sets[i % nsets][1].data_ptr(), sets[i % nsets][0].data_ptr(), st)  # noqa: E731
    g = lambda i: plan.inverse(sets[i % nsets][2].data_ptr(), sets[i % nsets][1].data_ptr(), st)  # noqa: E731
    for i in range(2):
        f(i), g(i)
    ms_f = time_events(torch, stream, f, reps)
    ms_i = time_events(torch, stream, g, reps)
    ms_rt = time_events(torch, stream, lambda i: (f(i), g(i)), reps)
    res = {"log2_m": log2_m, "batch": batch, "splits": plan.splits, "forward_ms": ms_f, "inverse_ms": ms_i,
           "roundtrip_ms": ms_rt, "forward_gelem_s": words / ms_f / 1e6, "inverse_gelem_s": words / ms_i / 1e6,
           "value": 2.0 * words / ms_rt / 1e6, "unit": "Gelem/s", "reps": reps}
    plan.close()
    del sets
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------
def verify_sharded(torch, lib, a2a, stream, dev, log2_m, rank, world):
    """Parity of what is about to be timed (outside the timed region): (1) every word of this rank's slice of the
    sharded forward against the single-GPU plan's transform of the same input; (2) directly evaluated output words
    (a 64-term DFT sum each, plain Python integers - independent of every kernel and table of this repo) on a sparse
    input; (3) the round trip.  Returns the report and the single-GPU timing taken on the way."""
    st = stream.cuda_stream
    m, n0, n1 = 1 << log2_m, a2a.n0, a2a.n1
    local = m // world
    rep = {}
    gen = torch.Generator(device=dev)
    gen.manual_seed(4242 + log2_m)  # same full input on every rank
    full = torch.randint(0, 2**62, (m,), dtype=torch.int64, device=dev, generator=gen)
    src = full.view(n0, n1)[:, rank * n1 // world:(rank + 1) * n1 // world].contiguous().view(-1)
    mid = torch.empty_like(src)
    a2a.forward(mid, src, st)
    ref = lib.plan(log2_m, device=dev.index)
    want = torch.empty_like(full)
    ref.forward(want.data_ptr(), full.data_ptr(), st)
    torch.cuda.synchronize()
    rep["forward_equal_single_gpu"] = bool(torch.equal(mid, want[rank * local:(rank + 1) * local]))
    rep["single_gpu_splits"] = ref.splits
    back = torch.empty_like(src)
    a2a.inverse(back, mid, st)
    torch.cuda.synchronize()
    rep["roundtrip_equal"] = bool(torch.equal(back, src))
    # the same-size single-GPU figure, measured on this GPU in this run
    f = lambda i: ref.forward(want.data_ptr(), full.data_ptr(), st)  # noqa: E731
    g = lambda i: ref.inverse(want.data_ptr(), want.data_ptr(), st)  # noqa: E731
    f(0), g(0)
    ms_f = time_events(torch, stream, f, 3)
    ms_i = time_events(torch, stream, g, 3)
    one = {"forward_ms": ms_f, "inverse_ms": ms_i, "value": 2.0 * m / (ms_f + ms_i) / 1e6, "unit": "Gelem/s",
           "forward_gelem_s": m / ms_f / 1e6, "inverse_gelem_s": m / ms_i / 1e6, "splits": ref.splits,
           "note": "plain single-GPU plan of the same 2^%d transform, timed on this GPU in this run (3 reps)" % log2_m}
    ref.close()
    del want, full, back
    # (2) sparse input, direct evaluation
    rng = np.random.default_rng(99)
    pos = np.unique(rng.integers(0, m, 64, dtype=np.int64))
    val = rng.integers(1, 2**62, pos.size, dtype=np.int64)
    r_, c_ = pos // n1, pos % n1
    blk = n1 // world
    mine = (c_ // blk) == rank
    sparse = torch.zeros(local, dtype=torch.int64, device=dev)
    if mine.any():
        idx = r_[mine] * blk + (c_[mine] - rank * blk)
        sparse[torch.from_numpy(idx).to(dev)] = torch.from_numpy(val[mine]).to(dev)
    a2a.forward(mid, sparse, st)
    torch.cuda.synchronize()
    omega = pow(G0, (P0 - 1) >> log2_m, P0)
    probe = np.unique(np.concatenate([rng.integers(0, local, 12, dtype=np.int64), [0, local - 1]]))
    got = mid[torch.from_numpy(probe).to(dev)].cpu().numpy().view(np.uint64)
    bad = 0
    for i, gword in zip(probe, got):
        k = int(f"{int(i) + rank * local:0{log2_m}b}"[::-1], 2)  # output index holds frequency bitrev(index)
        acc = 0
        for p_, v_ in zip(pos, val):
            acc = (acc + int(v_) * pow(omega, (k * int(p_)) % m, P0)) % P0
        bad += int(gword) != acc
    rep["direct_dft_words_checked"] = int(probe.size)
    rep["direct_dft_words_wrong"] = int(bad)
    del sparse, mid, src
    torch.cuda.empty_cache()
    ok = rep["forward_equal_single_gpu"] and rep["roundtrip_equal"] and bad == 0
    return ok, rep, one


def bench_sharded(args, lib, torch, dist, rank, world, local_rank, log2_m):
    import dist_ntt  # sve-ntt_b200/dist_ntt.py
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream
    dev = torch.device("cuda", local_rank)
    a2a = dist_ntt.ShardedNTT(lib, log2_m, world, rank, local_rank, mode=args.mode)
    plan = a2a.plan
    m = 1 << log2_m
    local = m // world

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    ok, parity, one_gpu = verify_sharded(torch, lib, a2a, stream, dev, log2_m, rank, world)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) != 1:
        raise SystemExit(f"rank {rank}: sharded transform failed its parity check: {parity}")

    sets, nsets, set_bytes = make_sets(torch, dev, local, rank, max_sets=2)

    def step(i):
        src, mid, out = sets[i % nsets]
        a2a.forward(mid, src, st)
        a2a.inverse(out, mid, st)

    def agree_max(v):
        tv = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        return float(tv.item())

    warm_n = warm_up(torch, step, max(args.warmup, 3), agree_max)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = 2.0 * m * args.steps / (ms_max * 1e-3) / 1e9
    assert torch.equal(sets[(args.steps - 1) % nsets][0], sets[(args.steps - 1) % nsets][2]), "round trip broke under load"

    # forward and inverse separately (max over ranks)
    src, mid, out = sets[0]
    reps = max(3, min(args.steps, 10))
    fi = {}
    for name, fn in (("forward", lambda i: a2a.forward(mid, src, st)), ("inverse", lambda i: a2a.inverse(out, mid, st))):
        barrier()
        ms = torch.tensor([time_events(torch, stream, fn, reps)], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        fi[name + "_ms"] = float(ms.item())
        fi[name + "_gelem_s"] = m / float(ms.item()) / 1e6

    # the pass fused with the exchange, alone: NVLink roofline (peer mode only - otherwise NCCL moves the data)
    roofline = None
    link_bytes = 8.0 * local * (world - 1) / world
    if a2a.mode == "peer":
        phases = {}
        for name in ("forward_cols_peer", "inverse_rows_peer"):
            tot = 0.0
            for r in range(reps + 1):
                buf, hdl, ptrs = a2a._next_peer()
                barrier()
                k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                k0.record(stream)
                if name == "forward_cols_peer":
                    plan.shard_forward_cols_peer(ptrs, src.data_ptr(), st)
                else:
                    plan.shard_inverse_rows_peer(ptrs, mid.data_ptr(), a2a._scratch(mid)[0].data_ptr(), st)
                k1.record(stream)
                torch.cuda.synchronize()
                if r:
                    tot += k0.elapsed_time(k1)
            ms = torch.tensor([tot / reps], dtype=torch.float64, device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            phases[name + "_ms"] = float(ms.item())
        dom = "forward_cols_peer"
        ach = link_bytes / (phases[dom + "_ms"] * 1e-3) / 1e9
        roofline = {"bound": "nvlink", "kernel": f"pass_kernel<2^{plan.splits[0]} columns, MAP, peer stores> ({dom})",
                    "achieved": ach, "peak": NVLINK_PEAK_GBS, "unit": "GB/s", "frac": ach / NVLINK_PEAK_GBS,
                    "frac_of_nominal": ach / NVLINK_NOMINAL_GBS, "traffic": None,
                    "peak_source": "B200_PROFILING.md: measured peer copy 770 GB/s per direction per GPU (900 nominal)",
                    "link_bytes_per_gpu_per_transform": link_bytes, "phases": phases,
                    "note": "bytes one GPU stores into its peers per transform, 8 (n/G) (G-1)/G, divided by the time of the "
                            "column pass that performs those stores (its arithmetic included); inverse_rows_peer additionally "
                            "runs the inner passes of a three-pass plan before its storing pass"}
    roofline_int = None
    if rank == 0:
        try:
            roofline_int = integer_roofline(lib, 2.0 * (m // 2) * log2_m / world, ms_max * 1e-3 / args.steps)
            roofline_int["bound"] = "imad"
            roofline_int["scope"] = "whole round trip per GPU (exchange included)"
        except Exception as exc:  # pragma: no cover
            roofline_int = {"error": str(exc)}

    # end to end: host buffers in, host buffers out (each rank its block / slice)
    e2e = None
    if not args.no_e2e:
        h_src = torch.empty(local, dtype=torch.int64).pin_memory()
        h_mid = torch.empty(local, dtype=torch.int64).pin_memory()
        h_out = torch.empty(local, dtype=torch.int64).pin_memory()
        h_src.copy_(sets[0][0].cpu())
        esteps = 3
        a2a.forward_host(h_mid, h_src, st)
        a2a.inverse_host(h_out, h_mid, st)
        assert torch.equal(h_out, h_src)
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            a2a.forward_host(h_mid, h_src, st)
            a2a.inverse_host(h_out, h_mid, st)
        torch.cuda.synchronize()
        td = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(td, op=dist.ReduceOp.MAX)
        # the copies alone, same pattern: what the host side (PCIe, host DRAM) allows
        d_a, d_b = a2a._host_staging(h_src)
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            for _ in range(2):
                d_a.copy_(h_src, non_blocking=True)
                h_mid.copy_(d_b, non_blocking=True)
                torch.cuda.synchronize()
        tc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        nbytes = 8 * local
        e2e = {"value": 2.0 * m * esteps / float(td.item()) / 1e9, "unit": "Gelem/s",
               "h2d_bytes_per_step": 2 * nbytes * world, "d2h_bytes_per_step": 2 * nbytes * world, "steps": esteps,
               "api": "ShardedNTT.forward_host + inverse_host: pinned host block -> H2D -> sharded transform -> D2H, per rank",
               "copies_only_value": 2.0 * m * esteps / float(tc.item()) / 1e9,
               "copies_only_gbs_per_gpu_per_direction": 2.0 * nbytes * esteps / float(tc.item()) / 1e9,
               "note": "copies_only_* = the same H2D + D2H traffic without any transform: the ceiling the host side "
                       "(PCIe, host memory) sets for this figure"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Gelem/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(f"dist{log2_m}", world), "warmup_steps_run": warm_n, "log2_m": log2_m, "splits": plan.splits,
                       "modulus": "0xfffffc6e80000001", "exchange": a2a.mode,
                       "exchange_fallback_reason": getattr(a2a, "_peer_error", None),
                       "l2": f"per-rank working set {set_bytes * nsets >> 20} MiB, far larger than L2",
                       "parallelism": f"column/row sharded over {world} GPUs, 1 all-to-all per transform "
                                      f"({'fused into the producing pass as peer stores' if a2a.mode == 'peer' else 'NCCL'})"},
            "clocks": clocks, "e2e": e2e,
            "gpu_launches": (2 * plan.launches + a2a.extra_launches_per_roundtrip) * args.steps,
            "roofline": roofline, "roofline_int": roofline_int, "fwd_inv": fi, "parity": parity,
            "one_gpu_same_workload": one_gpu,
            "speedup_vs_one_gpu_same_workload": value / one_gpu["value"],
        }
        emit(line)


# ------------------------------------------------------------------------------------------------
class _StdoutGuard:
    """Everything but the result goes to stderr: libraries (NCCL's version banner, for one) write to file descriptor 1
    behind Python's back, and the contract is ONE JSON line on stdout.  emit() writes to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.real, (text + "\n").encode())


_OUT = None


def emit(line):
    text = json.dumps(line)
    if _OUT is not None:
        _OUT.emit(text)
    else:
        print(text, flush=True)


def main():
    global _OUT
    args = parse_args()
    _OUT = _StdoutGuard()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return 0

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge

    pkg = ge.load_package()
    lib = pkg.load()  # raises if libxntt.so is missing - no fallback
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream
    dev = torch.device("cuda", local_rank)

    w = args.workload or default_workload(world)
    if w.startswith("dist") and world > 1:
        bench_sharded(args, lib, torch, dist, rank, world, local_rank, int(w[4:]))
        dist.destroy_process_group()
        return 0

    # ---- workloads without a data-path collective -----------------------------------------------------------
    scaling = "weak"
    if w == "ntt24":
        log2_m, batch = 24, 1
    elif w == "batch20":
        log2_m, total_batch = 20, 256
        assert total_batch % world == 0
        batch = total_batch // world
        scaling = "strong"
    elif w.startswith("dist"):
        log2_m, batch = int(w[4:]), 1
    else:
        raise SystemExit(f"unknown workload {w}")
    plan = lib.plan(log2_m, batch=batch, device=local_rank)
    m = 1 << log2_m
    local_words = m * batch
    sets, nsets, set_bytes = make_sets(torch, dev, local_words, rank)

    def fwd(i):
        plan.forward(sets[i % nsets][1].data_ptr(), sets[i % nsets][0].data_ptr(), st)

    def inv(i):
        plan.inverse(sets[i % nsets][2].data_ptr(), sets[i % nsets][1].data_ptr(), st)

    def step(i):
        fwd(i)
        inv(i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness of what is being timed (cheap, outside the timed region) -------------------
    step(0)
    torch.cuda.synchronize()
    assert torch.equal(sets[0][0], sets[0][2]), "inverse(forward(x)) != x"

    # ---- timed region --------------------------------------------------------------------------
    def agree_max(v):
        if world <= 1:
            return v
        tv = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        return float(tv.item())

    warm_n = warm_up(torch, step, max(args.warmup, 3), agree_max)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = 2.0 * local_words * world * args.steps / (ms_max * 1e-3) / 1e9

    # ---- forward and inverse separately (tests/bench-ntt.cpp:47-58 reports them as two benchmarks) ------------
    reps = max(10, min(args.steps, 50))
    for i in range(nsets):
        fwd(i)
    ms_f = time_events(torch, stream, fwd, reps)
    ms_i = time_events(torch, stream, inv, reps)
    fwd_inv = {"forward_ms": ms_f, "inverse_ms": ms_i, "forward_gelem_s": local_words / ms_f / 1e6,
               "inverse_gelem_s": local_words / ms_i / 1e6, "reps": reps, "per": "GPU"}

    # ---- per-kernel times (live, CUDA events, same buffers) -> roofline of the dominant kernel ----
    per_kernel = []
    hbm_peak, peak_src = measured_peaks()
    for inverse in (False, True):
        order = range(plan.launches) if not inverse else reversed(range(plan.launches))
        for p_i in order:
            run = lambda r, p_i=p_i, inverse=inverse: plan.run_pass(  # noqa: E731
                p_i, inverse, sets[r % nsets][1].data_ptr(), sets[r % nsets][0].data_ptr(), st)
            for r in range(3):
                run(r)
            us = time_events(torch, stream, run, reps) * 1e3
            per_kernel.append({"kernel": f"{'inv' if inverse else 'fwd'}_pass{p_i}_2^{plan.splits[p_i]}", "us": us,
                               "levels": plan.splits[p_i], "alg_gbs": 16.0 * local_words / (us * 1e-6) / 1e9})
    dom = max(per_kernel, key=lambda k: k["us"])
    traffic, traffic_src = archived_traffic(w, plan.splits, dom["kernel"])
    roofline_hbm = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["alg_gbs"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": dom["alg_gbs"] / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peak_src, "alg_bytes_per_launch": 16 * local_words,
                    "note": "a pass reads and writes every residue once: 16 B/element per launch.  traffic also holds the "
                            "16 B/element twiddle matrix the row pass streams where the plan stores one (DESIGN.md section "
                            "2): spare HBM bandwidth traded for one modular product per residue, not a re-read of data"}
    roofline = None
    if rank == 0:
        try:
            # the dominant kernel: its butterfly levels x (n/2) modular products
            roofline = integer_roofline(lib, batch * (m // 2) * dom["levels"], dom["us"] * 1e-6)
            roofline = dict({"bound": "imad", "kernel": dom["kernel"]}, **roofline)
            roofline["traffic"], roofline["traffic_source"] = traffic, traffic_src
            step_int = integer_roofline(lib, 2.0 * batch * (m // 2) * log2_m, ms_max * 1e-3 / args.steps)
            roofline["whole_step"] = {k: step_int[k] for k in ("achieved", "frac", "frac_wide_weighted")}
        except Exception as exc:  # pragma: no cover
            roofline = {"bound": "imad", "error": str(exc)}

    # ---- end to end through the host-buffer C ABI --------------------------------------------------
    e2e = None
    if not args.no_e2e:
        nbytes = 8 * local_words
        h_src = torch.empty(local_words, dtype=torch.int64).pin_memory()
        h_mid = torch.empty(local_words, dtype=torch.int64).pin_memory()
        h_out = torch.empty(local_words, dtype=torch.int64).pin_memory()
        h_src.copy_(sets[0][0].cpu())
        esteps = max(3, min(args.steps, 20))
        for _ in range(2):
            plan.forward_host(h_mid.data_ptr(), h_src.data_ptr())
            plan.inverse_host(h_out.data_ptr(), h_mid.data_ptr())
        assert torch.equal(h_out, h_src)
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            plan.forward_host(h_mid.data_ptr(), h_src.data_ptr())
            plan.inverse_host(h_out.data_ptr(), h_mid.data_ptr())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        td = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        e2e = {"value": 2.0 * local_words * world * esteps / float(td.item()) / 1e9, "unit": "Gelem/s",
               "h2d_bytes_per_step": 2 * nbytes * world, "d2h_bytes_per_step": 2 * nbytes * world, "steps": esteps,
               "api": "xntt_forward_host + xntt_inverse_host on pinned host buffers"}

    # ---- the other BASELINE configs on this one GPU (extras of the default line) -----------------------------
    extras = {}
    if rank == 0 and world == 1 and w == "ntt24" and not args.no_extras:
        del sets
        torch.cuda.empty_cache()
        try:
            extras["batch20"] = one_gpu_roundtrip(torch, lib, stream, dev, 20, 256, 10)
            extras["dist30"] = one_gpu_roundtrip(torch, lib, stream, dev, 30, 1, 3)
        except Exception as exc:  # pragma: no cover
            extras["error"] = repr(exc)

    # ---- CPU baselines on the box's host cores (rank 0, N = 1, bounded samples) ----------------------
    cpu, cpu_all = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, cpu_all = cpu_baselines(w)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Gelem/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(w, world), "warmup_steps_run": warm_n, "log2_m": log2_m, "batch_per_gpu": batch, "splits": plan.splits,
                       "modulus": "0xfffffc6e80000001",
                       "l2": f"ring of {nsets} buffer sets ({nsets * set_bytes >> 20} MiB) larger than L2, each step touches "
                             "the next set", "parallelism": f"{world}x independent"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": 2 * plan.launches * args.steps,
            "roofline": roofline, "roofline_hbm": roofline_hbm, "per_kernel": per_kernel, "fwd_inv": fwd_inv,
            "cpu_baseline": cpu, "cpu_baselines": cpu_all,
        }
        line.update(extras)
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
