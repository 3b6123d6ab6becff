#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_single.py -m gpu -q -x 2>&1 | tail -5
timeout 900 python bench_sweep.py --sizes 1600,2000,3200,4000 --batches "" --no-cpu --dtypes f64,f32 2>gpurun_out/sweep2.err | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l)
    print('%s nx %4d D %4d iters %4d us/iter %7.2f HBMeq %7.0f GB/s grid %3d rpc %2d smem %2d'%(r['dtype'],r['nx'],r['D'],r['iters'],r['us_per_iter_kernel'],r['hbm_equiv_gbs'],r['grid'],r['rows_per_cta'],r['rows_in_smem']))
"
tail -3 gpurun_out/sweep2.err
