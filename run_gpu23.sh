#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'us/iter',round(d['us_per_admm_iter_in_kernel'],3),'frac',round(d['roofline']['frac'],3),'cpu',round(d['cpu_baseline']['value']),d['cpu_baseline']['cores'],'clocks',d['clocks'])
for k,v in d.get('other_workloads',{}).items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in ('roofline','e2e','note')}, 'roof', v.get('roofline',{}).get('achieved'), 'e2e', v.get('e2e',{}).get('value'))
PY
tail -3 gpurun_out/bench_default.err
