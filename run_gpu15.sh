#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batched.py -m gpu -q > gpurun_out/pytest_batched.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_batched.log
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'solves/s',round(d['value']),'ms/step',round(d['ms_per_step'],2),'iters',round(d['iters_per_solve'],1),d['iters_max'],'sweeps',d['sweeps'],'TF/s',round(d['roofline']['achieved'],1),'e2e',round(d['e2e']['value']),'solved',d['all_solved'])"; }
for a in "--batch 4096" "--batch 4096 --batch-engine 2" "--batch 4096 --batch-engine 3" "--batch 16384" "--batch 65536"; do
timeout 300 python bench.py --workload mpc_batched --steps 3 --warmup 1 --no-cpu-baseline --no-extras $a 2>gpurun_out/e1 | show "[$a]"
done
