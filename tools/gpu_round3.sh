#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
python - <<'PY'
import sys, os, json, torch
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge
pkg = ge.load_package(); lib = pkg.load(); st = torch.cuda.current_stream().cuda_stream
res = []
for rows, cols in [(8192, 8192), (4096, 16384), (16384, 16384), (2048, 8192)]:
    src = torch.randint(0, 2**62, (rows, cols), dtype=torch.int64, device="cuda"); dst = torch.empty((cols, rows), dtype=torch.int64, device="cuda")
    for _ in range(3): lib.transpose(dst.data_ptr(), src.data_ptr(), rows, cols, rows, cols, st)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): lib.transpose(dst.data_ptr(), src.data_ptr(), rows, cols, rows, cols, st)
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)/10
    gbs = 16.0*rows*cols/ms/1e6
    res.append({"rows": rows, "cols": cols, "ms": ms, "GBps": gbs, "frac_of_6559.7": gbs/6559.7})
    print(res[-1])
    del src, dst
sq = torch.randint(0, 2**62, (8192, 8192), dtype=torch.int64, device="cuda")
for _ in range(3): lib.transpose(sq.data_ptr(), sq.data_ptr(), 8192, 8192, 8192, 8192, st)
torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): lib.transpose(sq.data_ptr(), sq.data_ptr(), 8192, 8192, 8192, 8192, st)
e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)/10
res.append({"inplace": 8192, "ms": ms, "GBps": 16.0*8192*8192/ms/1e6}); print(res[-1])
json.dump(res, open("gpurun_out/transpose_bw.json","w"), indent=1)
PY
python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline_int']['frac'])"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pass_kernel -s 12 -c 4 -o gpurun_out/prof_pass \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
