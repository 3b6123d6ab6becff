#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python tools/bench_sweep.py --out gpurun_out/sweep_r01b.json > gpurun_out/sweep.log 2> gpurun_out/sweep.err; echo "sweep rc=$?"
tail -5 gpurun_out/sweep.log | cut -c1-200
CMD1="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rqp_ --csv --log-file gpurun_out/launches_r01b.csv $CMD1 > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
timeout 200 python bench.py --workload mpc_batched --batch-dtype f64 --steps 2 --warmup 1 --no-extras > gpurun_out/bench_batched_f64.json 2>gpurun_out/bb64.err; echo "f64 batched rc=$?"
cut -c1-400 gpurun_out/bench_batched_f64.json
