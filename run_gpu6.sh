#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batched.py -m gpu -q -x > gpurun_out/pytest_batched.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_batched.log
timeout 600 python bench.py --workload mpc_batched --steps 3 --warmup 1 --no-cpu-baseline 2> gpurun_out/bb.err | tee gpurun_out/bench_batched_f64.json | cut -c1-1500; tail -3 gpurun_out/bb.err
timeout 600 python bench.py --workload mpc_batched --batch-dtype f32 --steps 3 --warmup 1 --no-cpu-baseline 2> gpurun_out/bb32.err | tee gpurun_out/bench_batched_f32simt.json | cut -c1-1500; tail -3 gpurun_out/bb32.err
