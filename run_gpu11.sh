#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 single rc=$?"
tail -1 gpurun_out/bench_n2.json | cut -c1-700; tail -3 gpurun_out/bench_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 1 --workload mpc_batched > gpurun_out/bench_n2_batched.json 2> gpurun_out/bench_n2_batched.err; echo "n2 batched rc=$?"
tail -1 gpurun_out/bench_n2_batched.json | cut -c1-900; tail -3 gpurun_out/bench_n2_batched.err
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err; echo "n2 ref rc=$?"
tail -1 gpurun_out/bench_n2_ref.json | cut -c1-300
