#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $NG --steps 100 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "n8 single rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1]); print('N8 value',round(d['value']),'e2e',round(d['e2e']['value']),'n_gpus',d['n_gpus'],'clocks',d['clocks'])"
timeout 600 $TR bench.py --gpus $NG --steps 5 --warmup 3 --workload mpc_batched > gpurun_out/bench_n8_batched.json 2> gpurun_out/bench_n8_batched.err; echo "n8 batched rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_n8_batched.json').read().strip().splitlines()[-1]); print('N8 batched value',round(d['value']),'e2e',round(d['e2e']['value']),'n_gpus',d['n_gpus'],'ms',d['ms_per_step'])"
timeout 300 $TR bench.py --impl reference --gpus $NG --steps 10 --warmup 3 2>/dev/null | cut -c1-200
timeout 300 python -m pytest tests/test_sharding.py -q 2>&1 | tail -2
