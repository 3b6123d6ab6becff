#!/bin/bash
# GPU tests, then the single-QP bench line with its extras condensed to a few lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload mpc_single --no-cpu-baseline > gpurun_out/bench_single_r02.json 2> gpurun_out/bench_single_r02.err
python - <<'EOF'
import json
d=json.loads(open("gpurun_out/bench_single_r02.json").read().strip().splitlines()[-1])
print("mpc_single value %.0f e2e %.0f (%.2f of device) resolve %.0f us/iter %.3f launches %s" % (d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["e2e"]["resolve"]["value"], d["us_per_admm_iter_in_kernel"], d["gpu_launches"]))
for k,v in d["other_workloads"].items():
    print(" ", k, {kk: (round(vv,1) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("value","ms_per_step","iters_per_solve","error")}, "e2e", v.get("e2e",{}).get("value"))
EOF
tail -3 gpurun_out/bench_single_r02.err
