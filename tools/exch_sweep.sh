#!/bin/bash
# sweep of the exchange tuning knobs on the default workload (C2): replicas x pre-poll spin
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/exch_sweep.txt
for rep in 1 2 4 8; do
  for pp in 600 400 200 -1; do
    python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras --exch-flags $((rep*256)) --prepoll $pp 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('rep $rep prepoll $pp value %.0f us/iter %.3f phases %s' % (d['value'], d['us_per_admm_iter_in_kernel'], d['phase_cycles_per_iter']))" >> gpurun_out/exch_sweep.txt
  done
done
cat gpurun_out/exch_sweep.txt
