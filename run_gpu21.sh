#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batched.py -m gpu -q 2>&1 | tail -4
timeout 900 python bench_sweep.py --sizes "" --batches 256,4096,16384 --no-cpu --dtypes f64 2>gpurun_out/sweep3.err | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l)
    print('%s B %6d ms %8.2f solves/s %9.0f iters %.1f/%d sweeps %d TF/s %6.1f solved %s'%(r['dtype'],r['B'],r['ms'],r['solves_per_s'],r['iters_mean'],r['iters_max'],r['sweeps'],r['alg_tflops'],r['all_solved']))
"
tail -3 gpurun_out/sweep3.err
