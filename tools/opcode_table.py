# SPDX-License-Identifier: Apache-2.0
"""Per-opcode executed-instruction table of one kernel from an ncu report (source page, needs -lineinfo /
--import-source): warp-level executed instructions summed per SASS opcode, normalised per butterfly.

    python tools/opcode_table.py REPORT.ncu-rep KERNEL_ID BUTTERFLIES [--json out.json]

IMAD.WIDE, IMAD.HI and IMAD (and IMAD.MOV / IMAD.X / IMAD.SHL / IMAD.IADD, which ptxas emits for moves and adds)
all issue on the fma pipe; the 'fma_pipe_non_product' line says how many of those issues are not products."""
import collections
import csv
import io
import json
import subprocess
import sys


def main():
    rep, kid, bfly = sys.argv[1], sys.argv[2], float(sys.argv[3])
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"],
                         capture_output=True, text=True, check=True).stdout
    lines = out.splitlines()
    name = lines[0]
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]
    i_src, i_ex = hdr.index("Source"), hdr.index("Instructions Executed")
    i_samples = hdr.index("# Samples")
    per, samples = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= i_ex:
            continue
        toks = r[i_src].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") else toks[0]
        # keep the modifiers that change the pipe / cost: IMAD.WIDE(.U32)(.X), IMAD.HI, IMAD.MOV, IMAD.X, IMAD.SHL, IMAD.IADD
        parts = op.split(".")
        key = parts[0]
        if key == "IMAD":
            for m in ("WIDE", "HI", "MOV", "SHL", "IADD", "X"):
                if m in parts[1:]:
                    key = "IMAD." + m
                    break
        elif key in ("LDG", "STG", "LDS", "STS", "BAR"):
            key = ".".join(p for p in parts if p in (key, "128", "64", "SYNC"))
        per[key] += int(r[i_ex])
        samples[key] += int(r[i_samples])
    total = sum(per.values())
    fma = {k: v for k, v in per.items() if k.startswith("IMAD")}
    products = sum(v for k, v in fma.items() if k in ("IMAD.WIDE", "IMAD.HI", "IMAD"))
    table = {"kernel": name.strip('",'), "butterflies": bfly, "warp_instructions": total,
             "per_butterfly_thread": {k: round(v * 32 / bfly, 3) for k, v in per.most_common()},
             "total_per_butterfly": round(total * 32 / bfly, 2),
             "fma_pipe_issues_per_butterfly": round(sum(fma.values()) * 32 / bfly, 2),
             "fma_pipe_non_product_per_butterfly": round((sum(fma.values()) - products) * 32 / bfly, 2),
             "stall_samples_share": {k: round(v / max(1, sum(samples.values())), 4) for k, v in samples.most_common(12)}}
    txt = json.dumps(table, indent=1)
    if "--json" in sys.argv:
        open(sys.argv[sys.argv.index("--json") + 1], "w").write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()
