for c in 2 1; do
  for all in "" 1; do
    if [ -n "$all" ]; then export RQP_TC_CHUNK_ALL=1; else unset RQP_TC_CHUNK_ALL; fi
    RQP_TC_CHUNK=$c python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunk $c all=$all: value %.0f e2e %.0f iters %.1f max %d window_ms %.3f frac %.3f' % (d['value'], d['e2e']['value'], d['iters_per_solve'], d['iters_max'], d['roofline']['launch_ms'], d['roofline']['frac']))"
  done
done
