set -u
mkdir -p gpurun_out
echo "== driver-style bench"
timeout -k 10 500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo rc=$?
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench_default.json") if l.startswith("{")][-1]); c=d["details"]
    print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "blocking", {k: round(v,1) for k,v in d["blocking_calls"].items() if isinstance(v,float)}, "f32", round(d["e2e_f32"]["value"],1))
    print("phases", {k: round(v,1) for k,v in c["phase_us"].items()}, "iters", c["icp_iterations_per_scan"], "det", c["deterministic_replay"])
    print({k: (round(v["scans_per_s"]), round(v["min_ms"],2), round(v["max_ms"],2)) for k,v in d["windows"].items()})
    print("cpu", {k: v for k, v in d["cpu_baseline"].items() if k not in ("sample",)}, "nn", d["nn_kernel"]["frac"] if d["nn_kernel"] else None, "ms", d["multi_stream"])
    print("quality", c["trajectory_quality"])
except Exception as e:
    print("failed", e)
PY
tail -3 gpurun_out/r2_bench_default.err
echo "== reference arm"
timeout -k 10 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-600
echo "== full suite x2 (hang hunt)"
for i in 1 2; do
timeout -k 10 240 python -X faulthandler -m pytest tests -m gpu -q -x -o faulthandler_timeout=120 > gpurun_out/r2_t9_$i.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2_t9_$i.log
done
