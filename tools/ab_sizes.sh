#!/bin/bash
# A/B of library variants on whole solves at several batch sizes: bash tools/ab_sizes.sh "<sizes>" base pre ...
cd "$(dirname "$0")/.."
SIZES=$1; shift
run() { python bench.py --no-extras --no-cpu-baseline --steps 4 --warmup 2 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'ms %.3f' % d['ms_per_step'], 'win %.3f' % d['roofline']['launch_ms'])"; }
cp reluqp-py_b200/lib/librqp.so /tmp/keep.so
for rep in 1 2; do for v in "$@"; do
  if [ "$v" != base ]; then cp reluqp-py_b200/lib/librqp_$v.so reluqp-py_b200/lib/librqp.so; else cp /tmp/keep.so reluqp-py_b200/lib/librqp.so; fi
  for B in $SIZES; do echo "$v B=$B: $(run --batch $B)"; done
done; done
cp /tmp/keep.so reluqp-py_b200/lib/librqp.so
