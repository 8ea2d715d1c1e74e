# SPDX-License-Identifier: Apache-2.0
"""bench.py - 64-bit NTT throughput (Gelem/s) on B200, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

A "step" is one out-of-place compute_forward followed by one compute_inverse (a full round trip,
the pair the reference's bench-ntt.cpp times as "Forward, ..." / "Inverse, ...") over one batch of
synthetic residues already resident in HBM.  Workloads:

  ntt24   (default)  single blocked six-step transform, n = 2^24, p = 0xfffffc6e80000001, g = 3
                     (BASELINE.json configs[1]).  With --gpus N > 1 every rank runs its own
                     transform on its own data (independent units, no data-path collective):
                     weak scaling.
  batch20            256 x 2^20 batched transforms, the batch sharded across ranks (configs[2]):
                     strong scaling, no collective.
  dist30 / distNN    one 2^NN transform sharded over the ranks with a single NCCL all-to-all
                     (configs[3]); on one GPU the plain three-pass plan.

value = elements transformed per second (2 transforms x n x batch per step) summed over ranks,
from CUDA events on the launching stream around exactly K steps, max over ranks.  e2e = the same
metric through the host-buffer entry points of the C ABI (xntt_forward_host + xntt_inverse_host on
pinned host memory: both PCIe copies inside the timed region).
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ntt24")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, splits, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed
    ncu --set full capture of this very command (profiles/r1_v6_ncu_pass_kernels.json); None when the
    capture does not cover the configuration."""
    path = os.path.join(ROOT, "profiles", "r1_v6_ncu_pass_kernels.json")
    if workload != "ntt24" or list(splits) != [11, 13] or not os.path.exists(path):
        return None
    order = ["fwd_pass0_2^11", "fwd_pass1_2^13", "inv_pass1_2^13", "inv_pass0_2^11"]  # launch order in the capture
    if kernel not in order:
        return None
    with open(path) as fh:
        k = json.load(fh)["kernels"][order.index(kernel)]
    return (float(k["dram__bytes_read.sum [Mbyte]"]) + float(k["dram__bytes_write.sum [Mbyte]"])) * 1e6


# ------------------------------------------------------------------------------------------------
# clocks: NVML sampled in a thread during the timed region
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
def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path: NTTReference (tests/ntt-reference.hpp)
    compiled into oracle/_ref; the oracle port if that library was never built.  Serial code: one
    host core is all it can use."""
    if rank != 0:
        return
    import oracle_lib
    if oracle_lib.have_reference():
        impl, kind = oracle_lib.Reference(), "reference"
    else:
        impl, kind = oracle_lib.Oracle(), "port"
    orc = oracle_lib.Oracle()
    total = args.steps + args.warmup
    # one step of NTTReference at 2^24 costs ~14 s: shrink the sample so the run ends in minutes
    log2_n = 24 if total <= 8 else (22 if total <= 40 else 20)
    log2_n = int(os.environ.get("XNTT_BENCH_REF_LOG2", log2_n))  # tests shrink the sample
    n = 1 << log2_n
    # the class is serial; what shards over N GPUs on our side (one independent transform per GPU) runs here as N
    # independent transforms on N host threads (ctypes releases the GIL) - all the threads this workload can use
    from concurrent.futures import ThreadPoolExecutor
    lanes = max(1, min(args.gpus, os.cpu_count() or 1))
    inputs = [orc.fill_xorshift(n, SEED + i, P0) for i in range(lanes)]

    def one(a):
        return impl.ntt_inverse(impl.ntt_forward(a, P0, G0), P0, G0)

    with ThreadPoolExecutor(max_workers=lanes) as pool:
        for _ in range(args.warmup):
            list(pool.map(one, inputs))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            list(pool.map(one, inputs))
        dt = time.perf_counter() - t0
    value = 2.0 * n * lanes * args.steps / dt / 1e9
    sample = (f"forward+inverse NTTReference n=2^{log2_n}, p=0xfffffc6e80000001, {lanes} independent transform(s) on "
              f"{lanes} host thread(s) (the class itself is serial)")
    line = {
        "impl": "reference", "metric": "64-bit NTT throughput (forward+inverse round trip)", "value": value,
        "unit": "Gelem/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Gelem/s", "cores": lanes, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Gelem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(args):
    w = args.workload
    if w == "ntt24":
        return "blocked six-step forward+inverse NTT n=2^24 uint64, p=0xfffffc6e80000001 g=3, one transform per GPU"
    if w == "batch20":
        return "batched 256x NTT n=2^20 forward+inverse, batch sharded across GPUs"
    if w.startswith("dist"):
        return f"distributed six-step forward+inverse NTT n=2^{int(w[4:])}, one all-to-all per transform"
    return w


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
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

    # ---- workload ----------------------------------------------------------------------------
    w = args.workload
    scaling = "weak"
    a2a = None
    if w == "ntt24":
        log2_m, batch, plan = 24, 1, None
        plan = lib.plan(log2_m, device=local_rank)
    elif w == "batch20":
        log2_m, total_batch = 20, 256
        assert total_batch % world == 0
        batch = total_batch // world
        plan = lib.plan(log2_m, batch=batch, device=local_rank)
        scaling = "strong"
    elif w.startswith("dist"):
        log2_m, batch = int(w[4:]), 1
        scaling = "strong"
        if world == 1:
            plan = lib.plan(log2_m, device=local_rank)
        else:
            import dist_ntt  # sve-ntt_b200/dist_ntt.py
            a2a = dist_ntt.ShardedNTT(lib, log2_m, world, rank, local_rank)
            plan = a2a.plan
    else:
        raise SystemExit(f"unknown workload {w}")
    m = 1 << log2_m
    local_words = (m * batch) if a2a is None else (m // world)
    units_per_step_local = 2 * local_words  # elements transformed: forward + inverse

    # ring of buffer sets so that every step starts on data that is not in L2
    set_bytes = 3 * 8 * local_words
    nsets = max(2, min(8, int(1.5 * 2**30 // set_bytes))) if set_bytes < 2**30 else 1
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    sets = []
    for _ in range(nsets):
        src = torch.randint(0, 2**62, (local_words,), dtype=torch.int64, device=dev, generator=gen)
        sets.append((src, torch.empty_like(src), torch.empty_like(src)))

    def step(i):
        src, mid, out = sets[i % nsets]
        if a2a is None:
            plan.forward(mid.data_ptr(), src.data_ptr(), st)
            plan.inverse(out.data_ptr(), mid.data_ptr(), st)
        else:
            a2a.forward(mid, src, st)
            a2a.inverse(out, mid, st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness of what is being timed (cheap, outside the timed region) -------------------
    step(0)
    torch.cuda.synchronize()
    assert torch.equal(sets[0][0], sets[0][2]), "inverse(forward(x)) != x"

    # ---- timed region --------------------------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        step(i)
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
    total_units = units_per_step_local * world * args.steps
    value = total_units / (ms_max * 1e-3) / 1e9
    launches_per_step = 2 * plan.launches + (0 if a2a is None else a2a.extra_launches_per_roundtrip)

    # ---- per-kernel times (live, CUDA events, same buffers) -> roofline of the dominant kernel ----
    roofline, per_kernel = None, []
    hbm_peak, peak_src = measured_peaks()
    if a2a is None:
        src, mid, out = sets[0]
        reps = max(10, min(args.steps, 50))
        for inverse in (False, True):
            order = range(plan.launches) if not inverse else reversed(range(plan.launches))
            for p_i in order:
                for _ in range(3):
                    plan.run_pass(p_i, inverse, mid.data_ptr(), src.data_ptr(), st)
                torch.cuda.synchronize()
                k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                k0.record(stream)
                for r in range(reps):
                    s_, m_, _ = sets[r % nsets]
                    plan.run_pass(p_i, inverse, m_.data_ptr(), s_.data_ptr(), st)
                k1.record(stream)
                torch.cuda.synchronize()
                us = k0.elapsed_time(k1) / reps * 1e3
                per_kernel.append({"kernel": f"{'inv' if inverse else 'fwd'}_pass{p_i}_2^{plan.splits[p_i]}",
                                   "us": us, "alg_gbs": 16.0 * local_words / (us * 1e-6) / 1e9})
        dom = max(per_kernel, key=lambda k: k["us"])
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["alg_gbs"], "peak": hbm_peak,
                    "unit": "GB/s", "frac": dom["alg_gbs"] / hbm_peak, "traffic": ncu_traffic(w, plan.splits, dom["kernel"]),
                    "peak_source": peak_src,
                    "alg_bytes_per_launch": 16 * local_words,
                    "note": "kernel reads and writes every residue once: 16 B/element per launch; the kernels are "
                            "bound by the IMAD pipe, see roofline_int.  traffic (ncu dram bytes of this launch) also "
                            "holds the 16 B/element twiddle matrix the row pass / inverse column pass streams where the "
                            "plan stores one (DESIGN.md section 2): spare HBM bandwidth traded for one modular product "
                            "per residue, not a re-read of data"}

    # ---- integer roofline: measured IMAD / IMAD.WIDE rates ------------------------------------------
    roofline_int = None
    if rank == 0:
        try:
            imad, _ = lib.microbench(0, 1000)
            wide, _ = lib.microbench(1, 1000)
            bf, _ = lib.microbench(3, 1000)
            # canonical count: (n/2) log2 n modmuls, 11 32-bit multiply-class instructions
            # each (mul.lo64 = 1 wide + 2 narrow, two mul.hi64 = 4 wide each)
            modmuls = 2 * batch * (m // 2) * log2_m * (1 if a2a is None else 1.0 / world)
            t_int = modmuls * (2.0 / (imad * 1e9) + 9.0 / (wide * 1e9))
            roofline_int = {"bound": "imad", "imad_gops": imad, "imad_wide_gops": wide, "butterfly_gops": bf,
                            "modmuls_per_step": modmuls, "t_int_ms": t_int * 1e3,
                            "frac": t_int / (ms_max * 1e-3 / args.steps),
                            "note": "time of the canonical 11 multiply instructions per modmul at the measured "
                                    "IMAD / IMAD.WIDE issue rates, divided by the measured step time"}
        except Exception as exc:  # pragma: no cover
            roofline_int = {"error": str(exc)}

    # ---- end to end through the host-buffer C ABI --------------------------------------------------
    e2e = None
    if not args.no_e2e and a2a is None:
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
        e2e = {"value": units_per_step_local * world * esteps / float(td.item()) / 1e9, "unit": "Gelem/s",
               "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": 2 * nbytes, "steps": esteps,
               "api": "xntt_forward_host + xntt_inverse_host on pinned host buffers"}

    # ---- CPU baseline on the box's host cores (rank 0, N = 1, bounded sample) ----------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle_lib
        kind = "reference" if oracle_lib.have_reference() else "port"
        impl = oracle_lib.Reference() if kind == "reference" else oracle_lib.Oracle()
        n_cpu = 1 << 22
        a = oracle_lib.Oracle().fill_xorshift(n_cpu, SEED, P0)
        t0 = time.perf_counter()
        back = impl.ntt_inverse(impl.ntt_forward(a, P0, G0), P0, G0)
        dt = time.perf_counter() - t0
        assert np.array_equal(back, a)
        cpu = {"value": 2.0 * n_cpu / dt / 1e9, "unit": "Gelem/s", "cores": 1, "kind": kind,
               "sample": "one forward+inverse of NTTReference (tests/ntt-reference.hpp) at n=2^22, real prime, "
                         "1 thread (the class is serial); host has %d cores" % (os.cpu_count() or 0)}

    if rank == 0:
        line = {
            "metric": "64-bit NTT throughput (forward+inverse round trip)", "value": value, "unit": "Gelem/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(args), "log2_m": log2_m, "batch_per_gpu": batch, "splits": plan.splits,
                       "modulus": "0xfffffc6e80000001", "l2": f"ring of {nsets} buffer sets ({nsets * set_bytes >> 20} MiB) "
                       "larger than L2, each step touches the next set", "parallelism": f"{world}x independent"
                       if a2a is None else f"column/row sharded over {world} GPUs, 1 all-to-all per transform"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "roofline_int": roofline_int, "per_kernel": per_kernel, "cpu_baseline": cpu,
            "fwd_inv": None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
