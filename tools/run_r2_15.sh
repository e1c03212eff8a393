#!/bin/bash
# A/B: warps of the solver's scheduler own no source points (KB_ICP_SOLVER_SMSP=3) vs all warps (-1); parity tests with the default
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in -1 3 -1 3; do
  KB_ICP_SOLVER_SMSP=$v timeout 300 python bench.py --steps 20 --warmup 5 --repeats 9 --no-nn --no-cpu --no-extra --streams 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('smsp $v value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'blocking', round(d['blocking_calls']['value_resident'],1))"
done
KB_ICP_SOLVER_SMSP=3 timeout 200 python tools/icp_timeline.py 100 3 2>&1 | grep -E "iters|map update" | tail -4
KB_ICP_SOLVER_SMSP=-1 timeout 200 python tools/icp_timeline.py 100 3 2>&1 | grep -E "iters|map update" | tail -4
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
