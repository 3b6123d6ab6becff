#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD2="python bench.py --workload mpc_batched --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
timeout 200 $CMD2 > gpurun_out/plain2.log 2>&1; echo "plain2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rqp_batched_tc_kernel -s 4 -c 1 -o gpurun_out/prof_batched_win_r01d $CMD2 > gpurun_out/ncu_win.log 2>&1; echo "full rc=$?"
tail -3 gpurun_out/ncu_win.log
ls -la gpurun_out/prof_batched_win_r01d.ncu-rep
