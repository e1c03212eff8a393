#!/bin/bash
# repeat the GPU parity file (hunting the rare failure of the 300-scan table-rebuild test); full logs, stop at the first failure
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
export KB_SYNC_TIMEOUT_S=15
for i in $(seq 1 ${1:-4}); do
  timeout -k 5 200 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -o faulthandler_timeout=180 > gpurun_out/r2_repeat_$i.log 2>&1
  rc=$?
  echo "run $i rc=$rc"; tail -1 gpurun_out/r2_repeat_$i.log
  if [ $rc -ne 0 ]; then grep -n -B2 -A25 "^E  \|Error\|phase marks" gpurun_out/r2_repeat_$i.log | cut -c1-400 | head -80; break; fi
done
