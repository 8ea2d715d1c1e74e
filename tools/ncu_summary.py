# SPDX-License-Identifier: Apache-2.0
"""Reduce an ncu --set full report to the handful of per-launch metrics DESIGN.md and bench.py quote.
   usage: python tools/ncu_summary.py REPORT.ncu-rep OUT.json "note" [name0,name1,... [split0,split1,...]]
Runs `ncu -i REP --page raw --csv` here (no GPU needed).  The optional kernel names (in launch order, the ones bench.py's
per_kernel block uses, e.g. fwd_pass0_2^11) and the plan's splits let bench.py attach `traffic` to its roofline."""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
]
STALLS = ["math_pipe_throttle", "wait", "not_selected", "dispatch_stall", "long_scoreboard", "barrier", "short_scoreboard",
          "mio_throttle", "lg_throttle", "no_instruction"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    names = sys.argv[4].split(",") if len(sys.argv) > 4 else []
    splits = [int(v) for v in sys.argv[5].split(",")] if len(sys.argv) > 5 else []
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    kernels = []
    for r in data:
        k = {}
        for name in KEEP:
            if name in ix:
                u = units[ix[name]]
                k[name + (f" [{u}]" if u else "")] = r[ix[name]]
        for s in STALLS:
            name = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if name in ix:
                k[f"stall_{s}_per_issue"] = r[ix[name]]
        if len(kernels) < len(names):
            k["name"] = names[len(kernels)]
        rd, wr = "dram__bytes_read.sum", "dram__bytes_write.sum"
        if rd in ix and wr in ix:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            k["dram_bytes"] = float(r[ix[rd]]) * scale[units[ix[rd]]] + float(r[ix[wr]]) * scale[units[ix[wr]]]
        kernels.append(k)
    json.dump({"source": rep, "note": note, "splits": splits, "kernels": kernels}, open(out, "w"), indent=1)
    for k in kernels:
        print(k["Kernel Name"][:110], k.get("gpu__time_duration.sum [us]", k.get("gpu__time_duration.sum [ns]")))


main()
