#!/bin/bash
# fp64 DMMA engine: RQP_DMMA_BIG = fewest active columns of a check window that get 128x128 tiles
# (below: 64x64 split-K tiles), swept per batch size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/dmma_ab.txt
run() { tag=$1; shift
  python bench.py --workload mpc_batched --steps 2 --warmup 1 --no-cpu-baseline --no-extras --batch-dtype f64 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$tag value %.0f ms %.3f iters %.1f' % (d['value'], d['ms_per_step'], d['iters_per_solve']))" >> gpurun_out/dmma_ab.txt; }
for B in 2048 4096 8192 16384; do
  for thr in 600 1217 2049 3000 4097 6000 9000; do
    RQP_DMMA_BIG=$thr run "B=$B thr=$thr" --batch $B
  done
done
cat gpurun_out/dmma_ab.txt
