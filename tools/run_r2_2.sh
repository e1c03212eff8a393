set -u
mkdir -p gpurun_out
echo "== tests"
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/r2_t2.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r2_t2.log
for q in 128 64 160; do
echo "== timeline team q=$q"
KB_ICP_TEAM_Q=$q timeout -k 10 200 python tools/icp_timeline.py 100 4 2>&1 | tail -8
done
echo "== bench team"
timeout -k 10 400 python bench.py --no-nn --no-cpu --streams 0 > gpurun_out/r2_b2_team.json 2> gpurun_out/r2_b2_team.err; echo rc=$?
python - <<'PY'
import json
for tag in ("team",):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2_b2_{tag}.json") if l.startswith("{")][-1]); c=d["details"]
        print(tag, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "blocking", {k: round(v,1) for k,v in d["blocking_calls"].items() if isinstance(v,float)}, "phases", {k: round(v,1) for k,v in c["phase_us"].items()}, "iters", c["icp_iterations_per_scan"], "det", c["deterministic_replay"])
        print(json.dumps(d["windows"]))
    except Exception as e:
        print(tag, "failed", e)
PY
tail -5 gpurun_out/r2_b2_team.err
