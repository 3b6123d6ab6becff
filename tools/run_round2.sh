#!/bin/bash
# round-2 evidence run, part A: the default bench line (mpc_batched through the sharded path + full-step extras), the
# reference arm, then the ncu launch list of the same command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
r = d['roofline']
print('value %.0f e2e %.0f (%.2f) iters %.1f launches %s | window %.3f ms achieved %.1f TF peak %.0f frac %.3f executed_frac %.3f | clocks %s' % (
    d['value'], d['e2e']['value'], d['e2e']['frac_of_device_timed'], d['iters_per_solve'], d['gpu_launches'], r['launch_ms'], r['achieved'], r['peak'], r['frac'], r['executed_frac'], d['clocks']))
print('cpu_baseline', {k: d['cpu_baseline'][k] for k in ('value', 'cores', 'processes', 'threads_per_process')})
for k, v in d.get('other_workloads', {}).items():
    print(' ', k, 'value', round(v.get('value', 0), 1), v.get('unit'), 'e2e', v.get('e2e', {}).get('value'), 'roofline', v.get('roofline', {}).get('frac'), 'iters', v.get('iters_per_solve'), v.get('error'))
ref = json.loads(open('gpurun_out/r02_bench_reference.json').read().strip().splitlines()[-1])
print('reference', round(ref['value'], 1), ref['cpu_baseline']['cores'], 'same config:', ref['config'] == d['config'])
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 $CMD > gpurun_out/plain_l.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:rqp|batch_|bgemm" --csv --log-file gpurun_out/r02_batched_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1; echo "launches rc=$?"
