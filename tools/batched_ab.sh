#!/bin/bash
# A/B of the batched engines: sparsity map (RQP_NO_KMASK) and window rotation (RQP_NO_ROTATE) on / off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/batched_ab.txt
run() { tag=$1; shift
  python bench.py --workload mpc_batched --steps 3 --warmup 2 --no-cpu-baseline --no-extras "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$tag %s value %.0f ms %.3f iters %.1f max %d solved %s' % (d['dtype'], d['value'], d['ms_per_step'], d['iters_per_solve'], d['iters_max'], d['all_solved']))" >> gpurun_out/batched_ab.txt; }
for B in ${BATCHES:-4096 16384}; do
  for dt in f32 f64; do
    run "B=$B base        " --batch $B --batch-dtype $dt
    RQP_NO_KMASK=1 run "B=$B nomask      " --batch $B --batch-dtype $dt
    RQP_NO_ROTATE=1 run "B=$B norot       " --batch $B --batch-dtype $dt
    RQP_NO_KMASK=1 RQP_NO_ROTATE=1 run "B=$B nomask norot" --batch $B --batch-dtype $dt
  done
done
cat gpurun_out/batched_ab.txt
