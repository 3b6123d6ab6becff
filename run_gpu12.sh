#!/bin/bash
cd /root/repo
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'us/iter',round(d['us_per_admm_iter_in_kernel'],3),'solves/s',round(d['value']),'e2e',round(d['e2e']['value']), d['phase_cycles_per_iter'])"; }
for args in "" "--exch-flags 1" "--prepoll 300" "--prepoll 600" "--prepoll 900" "--exch-flags 1 --prepoll 500" "--exch-flags 1 --prepoll 800"; do
  timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-extras $args 2>/dev/null | show "[$args]"
done
timeout 300 python -m pytest tests/test_gpu_single.py -m gpu -q -x 2>&1 | tail -3
