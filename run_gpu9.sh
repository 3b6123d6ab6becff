#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'solves/s',round(d['value']),'ms/step',round(d['ms_per_step'],2),'iters',round(d['iters_per_solve'],1),d['iters_max'],'sweeps',d['sweeps'],'TF/s',round(d['roofline']['achieved'],1),'e2e',round(d['e2e']['value']),'solved',d['all_solved'])"; }
timeout 300 python bench.py --workload mpc_batched --steps 3 --warmup 1 --no-cpu-baseline 2>gpurun_out/e1 | show tc4096_rnalo
