#!/bin/bash
# ncu evidence for profiles/ (each command is first run plain and must exit 0):
#   launch list of the default bench, --set full of the single-QP kernel, --set full of one window launch of the
#   tcgen05 batched kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r01d}
CMD1="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
CMD2="python bench.py --workload mpc_batched --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
timeout 200 $CMD1 > gpurun_out/plain1.log 2>&1; echo "plain1 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rqp_ --csv --log-file gpurun_out/${TAG}_single_mpc_launches.csv $CMD1 > gpurun_out/ncu_l1.log 2>&1; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rqp_single -s 4 -c 2 -o gpurun_out/prof_single_${TAG} $CMD1 > gpurun_out/ncu_s.log 2>&1; echo "single full rc=$?"
timeout 200 $CMD2 > gpurun_out/plain2.log 2>&1; echo "plain2 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:rqp|batch_|bgemm" --csv --log-file gpurun_out/${TAG}_batched_launches.csv $CMD2 > gpurun_out/ncu_l2.log 2>&1; echo "launches2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rqp_batched_tc_kernel -s 4 -c 1 -o gpurun_out/prof_batched_win_${TAG} $CMD2 > gpurun_out/ncu_win.log 2>&1; echo "win full rc=$?"
ls -la gpurun_out/*.ncu-rep
