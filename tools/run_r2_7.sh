set -u
mkdir -p gpurun_out
for q in 64 120; do
echo "== q=$q"
KB_ICP_TEAM_Q=$q timeout -k 10 200 python tools/icp_timeline.py 100 2 2>&1 | tail -8
KB_ICP_TEAM_Q=$q timeout -k 10 200 python tools/queue_timeline.py 100 40 2>&1 | tail -3
done
echo "== legacy queue timeline"
KB_ICP_TEAM_Q=0 timeout -k 10 200 python tools/queue_timeline.py 100 40 2>&1 | tail -3
echo "== bench team"
timeout -k 10 400 python bench.py --no-nn --no-cpu --streams 0 > gpurun_out/r2_b7_team.json 2> gpurun_out/r2_b7_team.err; echo rc=$?
python - <<'PY'
import json
for tag in ("team",):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2_b7_{tag}.json") if l.startswith("{")][-1]); c=d["details"]
        print(tag, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "blocking", {k: round(v,1) for k,v in d["blocking_calls"].items() if isinstance(v,float)}, "phases", {k: round(v,1) for k,v in c["phase_us"].items()}, "iters", c["icp_iterations_per_scan"], "det", c["deterministic_replay"])
        print({k: (round(v["scans_per_s"]), round(v["min_ms"],2), round(v["max_ms"],2)) for k,v in d["windows"].items()})
    except Exception as e:
        print(tag, "failed", e)
PY
tail -5 gpurun_out/r2_b7_team.err
echo "== full suite with faulthandler (hang hunt)"
timeout -k 10 200 python -X faulthandler -m pytest tests -m gpu -q -x -o faulthandler_timeout=90 > gpurun_out/r2_t7.log 2>&1; echo "tests rc=$?"
tail -30 gpurun_out/r2_t7.log
