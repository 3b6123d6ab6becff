#!/bin/bash
# A/B: bench line summary for a library variant
run() { python bench.py --no-extras --no-cpu-baseline --steps 5 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'win ms %.3f' % d['roofline']['launch_ms'], 'iters %.1f' % d['iters_per_solve'])"; }
cp reluqp-py_b200/lib/librqp.so /tmp/keep.so
for v in "$@"; do
  if [ "$v" != base ]; then cp reluqp-py_b200/lib/librqp_$v.so reluqp-py_b200/lib/librqp.so; else cp /tmp/keep.so reluqp-py_b200/lib/librqp.so; fi
  echo "$v B4096: $(run)"; echo "$v B16384: $(run --batch 16384)"; echo "$v B1024: $(run --batch 1024)"
done
cp /tmp/keep.so reluqp-py_b200/lib/librqp.so
