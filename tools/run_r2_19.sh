#!/bin/bash
# A/B: warm-up run of the solver's code (KB_ICP_WARM=1) vs none
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in 0 1 0 1; do
  KB_ICP_WARM=$v timeout 300 python bench.py --steps 20 --warmup 5 --repeats 9 --no-nn --no-cpu --no-extra --streams 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('warm $v value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'blocking', round(d['blocking_calls']['value_resident'],1))"
done
KB_ICP_WARM=1 timeout 200 python tools/icp_timeline.py 100 2 2>&1 | grep -E "^iters" | tail -2
KB_ICP_WARM=0 timeout 200 python tools/icp_timeline.py 100 2 2>&1 | grep -E "^iters" | tail -2
KB_ICP_WARM=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "kitti or queued or smoke or align" 2>&1 | tail -2
