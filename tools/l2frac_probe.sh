#!/bin/bash
# ring kernels with and without the L2 evict_last fraction hint (RQP_L2_FRAC=0 turns it off)
cd "$(dirname "$0")/.."
for nx in 2000 2500 3200 4000; do
  for f in 0 auto; do
    if [ $f = auto ]; then unset RQP_L2_FRAC; else export RQP_L2_FRAC=$f; fi
    echo "nx=$nx frac=$f: $(python tools/ring_probe.py --nx $nx --modes dense --reps 3 2>/dev/null | grep '^{' | python -c "
import sys,json
print('  '.join('%s %.2f us/iter %.0f GB/s' % (r['dtype'], r['us_per_iter'], r['GBs']) for r in map(json.loads, sys.stdin)))")"
  done
done
