#!/bin/bash
# ncu evidence for profiles/: launch list of the default bench command and one --set full capture of the batched
# tcgen05 kernel (each only after the same command exited 0 without ncu)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CMD1="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
CMD2="python bench.py --workload mpc_batched --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
timeout 200 $CMD1 > gpurun_out/plain1.log 2>&1; echo "plain1 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv $CMD1 > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
timeout 200 $CMD2 > gpurun_out/plain2.log 2>&1; echo "plain2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rqp_batched_tc_kernel -s 40 -c 2 -o gpurun_out/prof_batched_tc_r01b $CMD2 > gpurun_out/ncu_tc.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/*.ncu-rep
