#!/bin/bash
# same-box A/B of two builds of the library: tools/ab.sh old.so new.so  (alternates twice to expose drift)
set -u
for rep in 1 2; do
  for lib in "$@"; do
    echo "== $lib (rep $rep)"
    KB_LIB=$lib python bench.py --no-nn --no-cpu --steps 200 --warmup 10 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); c = d['config']
        print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'phases', c.get('phase_us'), 'iters', c.get('icp_iterations_mean'))
"
  done
done
