#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in 3 -1; do for i in 1 2 3; do
  KB_ICP_SOLVER_SMSP=$v KB_SYNC_TIMEOUT_S=15 timeout -k 5 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "long_stream or capacity_veto" > gpurun_out/r2_ls_${v}_$i.log 2>&1
  echo "smsp $v run $i rc=$?"; tail -1 gpurun_out/r2_ls_${v}_$i.log
done; done
grep -h "Error\|assert \|AssertionError\|phase marks\|watchdog" gpurun_out/r2_ls_*.log | cut -c1-400 | head -20
for k in async bulk regs; do KB_NN_KERNEL=$k timeout -k 10 200 python tools/nn_ab.py 2>&1 | tail -1; done
