#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
$CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; tail -3 gpurun_out/ncu_list.log
$CMD > gpurun_out/plain2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:rqp_single -s 4 -c 2 -o gpurun_out/prof_single_r01 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out/
