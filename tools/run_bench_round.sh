#!/bin/bash
# round-end style run: default bench (contract line), reference arm, then ncu launch list + one full capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'us/iter', d.get('us_per_admm_iter_in_kernel'), 'frac', d['roofline']['frac'], 'clocks', d['clocks'])
for k, v in d.get('other_workloads', {}).items():
    print(' ', k, round(v['value']), v.get('unit'), 'e2e', v.get('e2e', {}).get('value'), 'roofline', v.get('roofline', {}).get('frac'))
r = json.loads(open('gpurun_out/bench_reference.json').read().strip().splitlines()[-1])
print('reference', round(r['value']), r['cpu_baseline']['cores'])
PY
