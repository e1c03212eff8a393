#!/bin/bash
# A/B of the ICP team size: source points per team CTA (KB_ICP_TEAM_Q; default 64 -> 38..43 CTAs at KITTI shape)
cd "$GRAFT_REPO_ROOT"
for v in 64 48 80 96 64 48 80 96; do
  KB_ICP_TEAM_Q=$v timeout 200 python bench.py --steps 20 --warmup 5 --repeats 7 --no-nn --no-cpu --no-extra --streams 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('q $v value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'blocking', round(d['blocking_calls']['value_resident'],1))"
done
