#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
CMD="python bench.py --workload mpc_batched --batch 16384 --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_tc.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:rqp_batched_tc2 -s 10 -c 2 -o gpurun_out/prof_batched_tc2_r01 $CMD > gpurun_out/ncu_tc2.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_tc2.log
