#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log
python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','clocks')}); print(d['per_kernel'])
PY
python - <<'PY'
# runtime-modulus kernels vs baked-in modulus: same shape, Goldilocks / 62-bit prime
import sys, os, torch
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge
pkg = ge.load_package(); lib = pkg.load(); st = torch.cuda.current_stream().cuda_stream
def t(plan, n):
    src = torch.randint(0, 2**59, (n,), dtype=torch.int64, device="cuda"); dst = torch.empty_like(src)
    for _ in range(5): plan.forward(dst.data_ptr(), src.data_ptr(), st)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): plan.forward(dst.data_ptr(), src.data_ptr(), st)
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/20*1e3
for name,N,g in [("p0",pkg.P0,3),("goldilocks",0xFFFFFFFF00000001,7),("62bit",0x3A00000000000001,3),("64bit-random",0xA3B25F400C7A8001,5)]:
    L = 24
    if (N-1) % (1<<L): L = 15
    plan = lib.plan(L, modulus=N, generator=g, batch=(1 if L==24 else 512))
    n = (1<<L)*(1 if L==24 else 512)
    us = t(plan, n); print(f"field {name:14s} 2^{L}: fwd {us:8.1f} us  {n/us/1e3:6.1f} Gelem/s")
PY
