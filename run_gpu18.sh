#!/bin/bash
cd /root/repo
timeout 900 python bench_sweep.py --sizes 2000,4000 --batches "" --no-cpu --dtypes f64,f32 2>gpurun_out/sweep2.err | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l)
    print('%s nx %4d D %4d iters %4d us/iter %7.2f HBMeq %7.0f GB/s  phases[wait_v,gemv,barrier,finalize,checks,failed_polls,slab] %s'%(r['dtype'],r['nx'],r['D'],r['iters'],r['us_per_iter_kernel'],r['hbm_equiv_gbs'],r['phase_cycles_per_iter']))
"
