#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'us/iter',round(d['us_per_admm_iter_in_kernel'],3),'frac',round(d['roofline']['frac'],3),'cpu',d.get('cpu_baseline',{}).get('value'))
for k,v in d.get('other_workloads',{}).items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in ('roofline','e2e')}, 'roof', v.get('roofline',{}).get('achieved'), 'e2e', v.get('e2e',{}).get('value'))
PY
tail -3 gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_reference.json
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rqp_|FillFunctor<unsigned char>" -c 200 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
