#!/bin/bash
# verification: the full gpu suite, the parity file three more times (full logs), smoke, both bench arms
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
export KB_SYNC_TIMEOUT_S=20
timeout -k 5 400 python -X faulthandler -m pytest tests -m gpu -x -q -p no:cacheprovider -o faulthandler_timeout=380 > gpurun_out/r2_final_suite.log 2>&1
echo "suite rc=$?"; tail -2 gpurun_out/r2_final_suite.log
for i in 1 2 3; do
  timeout -k 5 300 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -o faulthandler_timeout=280 > gpurun_out/r2_parity_$i.log 2>&1
  echo "parity $i rc=$?"; tail -1 gpurun_out/r2_parity_$i.log
done
grep -h -B5 -A40 "^E  " gpurun_out/r2_final_suite.log gpurun_out/r2_parity_*.log | cut -c1-300 | head -80
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
echo "ref rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_reference.json').read().strip().splitlines()[-1]); print('reference arm', d['value'], d['cpu_baseline']['cores'])"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'blocking', d['blocking_calls'], 'ms', d.get('multi_stream',{}).get('scans_per_s'))
print('nn', {k: d['nn_kernel'][k] for k in ('kernel','ms','frac','other_variant')})
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
