#!/bin/bash
# final evidence of the round: gpu suite, smoke, NN sweep (config 5, up to 5 M stored points), both bench arms, ouster128 bench
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
export KB_SYNC_TIMEOUT_S=20
timeout -k 5 400 python -X faulthandler -m pytest tests -m gpu -x -q -p no:cacheprovider -o faulthandler_timeout=380 > gpurun_out/r2_final_suite.log 2>&1
echo "suite rc=$?"; tail -2 gpurun_out/r2_final_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
[ -n "$SKIP_SWEEP" ] || timeout 600 python tools/nn_bench.py 100000 200000 500000 1000000 2000000 5000000 10000000 > gpurun_out/r2_nn_sweep.jsonl 2> gpurun_out/r2_nn_sweep.err
echo "sweep rc=$?"; tail -1 gpurun_out/r2_nn_sweep.jsonl | cut -c1-300
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
echo "ref rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_default.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r2_bench_reference.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'blocking', {k: v for k, v in d['blocking_calls'].items() if k != 'note'})
print('multi', d.get('multi_stream',{}).get('scans_per_s'), 'nn', d['nn_kernel']['ms'], d['nn_kernel']['frac'], d['nn_kernel']['other_variant']['ms'])
print('reference arm', r['value'], r['cpu_baseline']['cores'], 'ratio', d['e2e']['value']/r['value'], 'cpu leg', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
[ -n "$SKIP_OUSTER" ] || timeout 600 python bench.py --workload ouster128 --steps 20 --warmup 5 --repeats 5 --no-nn > gpurun_out/r2_bench_ouster128.json 2> gpurun_out/r2_bench_ouster128.err
echo "ouster rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_ouster128.json').read().strip().splitlines()[-1]); print('ouster value', d['value'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])"
