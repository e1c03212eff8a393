set -u
mkdir -p gpurun_out
echo "== full suite"
timeout -k 10 400 python -X faulthandler -m pytest tests -m gpu -q -x -o faulthandler_timeout=200 > gpurun_out/r2_t12.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r2_t12.log
echo "== config 2 full length"
ORACLE_THREADS=32 timeout -k 10 600 python tools/config2_full.py 4541 > gpurun_out/r2_config2_full.json 2> gpurun_out/r2_config2_full.err; echo rc=$?
tail -2 gpurun_out/r2_config2_full.err; cat gpurun_out/r2_config2_full.json | cut -c1-900
echo "== ouster128 bench"
timeout -k 10 600 python bench.py --workload ouster128 --steps 20 --warmup 5 --repeats 5 --prime 40 --cpu-sample 20 --streams 0 --no-nn > gpurun_out/r2_bench_ouster128.json 2> gpurun_out/r2_bench_ouster128.err; echo rc=$?
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench_ouster128.json") if l.startswith("{")][-1]); c=d["details"]
    print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "blocking", {k: round(v,1) for k,v in d["blocking_calls"].items() if isinstance(v,float)})
    print("phases", {k: round(v,1) for k,v in c["phase_us"].items()}, "iters", c["icp_iterations_per_scan"], "src pts", c["icp_source_points"], "det", c["deterministic_replay"])
    print("cpu", {k: v for k, v in d["cpu_baseline"].items() if k not in ("sample",)})
    print("quality", c["trajectory_quality"])
except Exception as e:
    print("failed", e)
PY
tail -3 gpurun_out/r2_bench_ouster128.err
echo "== compute-sanitizer (smoke)"
for tool in memcheck racecheck synccheck; do
KB_WATCHDOG_SHIFT=31 KB_SYNC_TIMEOUT_S=0 timeout -k 10 420 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke:|Error|error" gpurun_out/r2_sanitizer_$tool.log | head -8
done
