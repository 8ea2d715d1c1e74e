#!/bin/bash
# Round-2 measurement run on one B200: GPU tests, butterfly lab, bench (driver's flags), reference arm,
# launch list and one ncu --set full capture of the pass kernels of the bench command.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw --format=csv > gpurun_out/smi.txt
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 120 tools/lab/_build/bfly_lab 500 > gpurun_out/bfly_lab.jsonl 2>&1; echo "lab rc=$?"; cat gpurun_out/bfly_lab.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'] and d['e2e']['value'])
print('fwd_inv', d['fwd_inv'])
print('roofline', {k: d['roofline'].get(k) for k in ('kernel', 'frac', 'frac_wide_weighted', 'imad_gops', 'imad_wide_gops')})
print('batch20', d.get('batch20')); print('dist30', d.get('dist30')); print('cpu', d.get('cpu_baselines'))
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
cat gpurun_out/bench_ref.json | cut -c1-400
S="--steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
python bench.py $S > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py $S > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
python bench.py $S > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pass_kernel -s 16 -c 4 -f -o gpurun_out/prof_pass_r2 \
    python bench.py $S > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
