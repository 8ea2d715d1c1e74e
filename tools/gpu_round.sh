#!/bin/bash
# Round-end measurement run (one B200): GPU tests, bench (full), launch list and one ncu --set full capture of the
# four pass kernels of the bench command, transposition bandwidth, per-pass times, field flavours.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline_int']['frac'], d['cpu_baseline'])"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
python bench.py --workload batch20 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_batch20.json 2> gpurun_out/bench_batch20.err; echo "bench batch20 rc=$?"
python tools/pass_times.py > gpurun_out/pass_times.log 2>&1; echo "pass_times rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pass_kernel -s 12 -c 4 -f -o gpurun_out/prof_pass \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
