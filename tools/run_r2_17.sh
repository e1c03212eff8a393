#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout -k 10 500 compute-sanitizer --tool initcheck --check-api-memory-access no --print-limit 60 python tools/initcheck_target.py 8 > gpurun_out/r2_sanitizer_initcheck.log 2>&1
echo "rc=$?"; grep -c "Uninitialized" gpurun_out/r2_sanitizer_initcheck.log; grep -A12 "Uninitialized" gpurun_out/r2_sanitizer_initcheck.log | head -120; tail -5 gpurun_out/r2_sanitizer_initcheck.log
