#!/bin/bash
# A/B of library variants (reluqp-py_b200/lib/librqp_<tag>.so) on ONE full check window: tools/window_counters.py
cd "$(dirname "$0")/.."
cp reluqp-py_b200/lib/librqp.so /tmp/keep.so
for v in "$@"; do
  if [ "$v" != base ]; then cp reluqp-py_b200/lib/librqp_$v.so reluqp-py_b200/lib/librqp.so; else cp /tmp/keep.so reluqp-py_b200/lib/librqp.so; fi
  for B in ${BS:-4096 16384}; do echo "$v: $(python tools/window_counters.py $B 2>&1 | tail -1)"; done
done
cp /tmp/keep.so reluqp-py_b200/lib/librqp.so
