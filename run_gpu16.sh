#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python bench_sweep.py --out gpurun_out/sweep_r01.json > gpurun_out/sweep.log 2> gpurun_out/sweep.err; echo "sweep rc=$?"
tail -5 gpurun_out/sweep.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/sweep_r01.json'))
for r in d['single']:
    print('%s nx %4d D %4d iters %4d %-6s us/iter %7.2f (cpu %8.1f x%5.1f) HBMeq %7.0f GB/s  grid %3d rpc %2d smem %2d reg %d setup %.2fs'%(r['dtype'],r['nx'],r['D'],r['iters'],r['status'][:6],r['us_per_iter_kernel'],r.get('cpu_us_per_iter',0),r.get('speedup_per_iter',0),r['hbm_equiv_gbs'],r['grid'],r['rows_per_cta'],r['rows_in_smem'],r['w_in_registers'],r['setup_s']))
print('cpu mpc solves/s', d.get('cpu_mpc_solves_per_s'))
for r in d['batched']:
    print('%s B %6d ms %8.2f solves/s %9.0f iters %.1f/%d sweeps %d TF/s %6.1f solved %s x%.0f'%(r['dtype'],r['B'],r['ms'],r['solves_per_s'],r['iters_mean'],r['iters_max'],r['sweeps'],r['alg_tflops'],r['all_solved'],r['speedup_vs_cpu'] or 0))
PY
