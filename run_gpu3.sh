#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'us/iter',round(d['us_per_admm_iter_in_kernel'],3),'solves/s',round(d['value']),'e2e',round(d['e2e']['value']), d['config']['launch'], 'reg' if d['w_in_registers'] else 'smem', d['phase_cycles_per_iter'])"; }
for args in "" "--w-residency 1" "--grid 138" "--grid 96 --w-residency 1" "--backoff 100" "--backoff 300" "--grid 60 --w-residency 1 --backoff 100" "--grid 60 --w-residency 1"; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $args 2>/dev/null | show "[$args]"
done
