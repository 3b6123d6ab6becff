#!/bin/bash
# A/B of the batched fp32 engine's scheduling switches: sparsity map (RQP_NO_KMASK), ticket scheduling of the
# window kernel (RQP_NO_TICKET -> rotated static assignment; + RQP_NO_ROTATE -> fixed static assignment)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/batched_ab.txt
run() { tag=$1; shift
  python bench.py --workload mpc_batched --steps 3 --warmup 2 --no-cpu-baseline --no-extras "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$tag %s value %.0f ms %.3f iters %.1f max %d solved %s' % (d['dtype'], d['value'], d['ms_per_step'], d['iters_per_solve'], d['iters_max'], d['all_solved']))" >> gpurun_out/batched_ab.txt; }
for B in ${BATCHES:-4096 16384}; do
    run "B=$B ticket          " --batch $B
    RQP_NO_TICKET=1 run "B=$B rotate          " --batch $B
    RQP_NO_TICKET=1 RQP_NO_ROTATE=1 run "B=$B static          " --batch $B
    RQP_NO_KMASK=1 run "B=$B ticket, no mask " --batch $B
    RQP_NO_NARROW=1 run "B=$B ticket, no narrow" --batch $B
done
cat gpurun_out/batched_ab.txt
