#!/bin/bash
# run the default bench once per variant library built by tools/variants.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp reluqp-py_b200/lib/librqp.so /tmp/librqp_keep.so
: > gpurun_out/variants.txt
for so in reluqp-py_b200/build/variants/librqp_*.so; do
  tag=$(basename $so .so)
  cp $so reluqp-py_b200/lib/librqp.so
  for nosm in "" 1; do
    RQP_NO_CHECK_SMEM=$nosm
    if [ -n "$nosm" ]; then export RQP_NO_CHECK_SMEM; else unset RQP_NO_CHECK_SMEM; fi
    if [ -n "$PROBE" ]; then echo "$tag: $(python $PROBE 2>&1 | tr "\n" " ")" >> gpurun_out/variants.txt; continue; fi
    python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extras $EXTRA 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$tag nosmem=$nosm value %.0f e2e %.0f us/iter %.3f phases %s' % (d['value'], d['e2e']['value'], d['us_per_admm_iter_in_kernel'], d['phase_cycles_per_iter']))" >> gpurun_out/variants.txt
  done
done
cp /tmp/librqp_keep.so reluqp-py_b200/lib/librqp.so
cat gpurun_out/variants.txt
