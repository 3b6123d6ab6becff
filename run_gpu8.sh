#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'solves/s',round(d['value']),'ms/step',round(d['ms_per_step'],2),'iters',round(d['iters_per_solve'],1),d['iters_max'],'sweeps',d['sweeps'],'TF/s',round(d['roofline']['achieved'],1),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']),'solved',d['all_solved'])"; }
timeout 300 python bench.py --workload mpc_batched --batch-dtype f32 --steps 3 --warmup 1 --no-cpu-baseline 2>gpurun_out/e1 | tee gpurun_out/bench_batched_tc.json | show tc4096
timeout 300 python bench.py --workload mpc_batched --batch-dtype f32 --batch-engine 1 --steps 3 --warmup 1 --no-cpu-baseline 2>gpurun_out/e2 | show simt4096
timeout 300 python bench.py --workload mpc_batched --batch-dtype f32 --batch 16384 --steps 2 --warmup 1 --no-cpu-baseline 2>gpurun_out/e3 | show tc16384
timeout 300 python bench.py --workload mpc_batched --batch-dtype f64 --steps 2 --warmup 1 --no-cpu-baseline 2>gpurun_out/e4 | show f64_4096
tail -2 gpurun_out/e1 gpurun_out/e3
CMD="python bench.py --workload mpc_batched --batch-dtype f32 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_tc.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:rqp_batched_tc -s 30 -c 2 -o gpurun_out/prof_batched_tc_r01 $CMD > gpurun_out/ncu_tc.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_tc.log
