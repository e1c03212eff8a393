set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt 2>&1
echo "== tests" 
timeout -k 10 900 python -m pytest tests -m gpu -q -x --timeout=300 > gpurun_out/r2_t1.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r2_t1.log
echo "== timeline team"
timeout -k 10 200 python tools/icp_timeline.py 100 6 > gpurun_out/r2_tl_team.log 2>&1; echo rc=$?; tail -7 gpurun_out/r2_tl_team.log
echo "== timeline legacy"
KB_ICP_TEAM_Q=0 timeout -k 10 200 python tools/icp_timeline.py 100 6 > gpurun_out/r2_tl_legacy.log 2>&1; echo rc=$?; tail -7 gpurun_out/r2_tl_legacy.log
echo "== nn A/B (checksums must agree)"
timeout -k 10 200 python tools/nn_ab.py 2>&1 | tail -1 | sed 's/^/bulk: /'
KB_NN_KERNEL=regs timeout -k 10 200 python tools/nn_ab.py 2>&1 | tail -1 | sed 's/^/regs: /'
echo "== bench team"
timeout -k 10 400 python bench.py --steps 200 --warmup 10 --no-nn --no-cpu > gpurun_out/r2_b_team.json 2> gpurun_out/r2_b_team.err; echo rc=$?
echo "== bench legacy"
KB_ICP_TEAM_Q=0 timeout -k 10 400 python bench.py --steps 200 --warmup 10 --no-nn --no-cpu > gpurun_out/r2_b_legacy.json 2> gpurun_out/r2_b_legacy.err; echo rc=$?
python - <<'PY'
import json
for tag in ("team","legacy"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2_b_{tag}.json") if l.startswith("{")][-1]); c=d["config"]
        print(tag, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "blocking", {k: round(v,1) for k,v in d["blocking_calls"].items() if isinstance(v,float)}, "phases", {k: round(v,1) for k,v in c["phase_us"].items()}, "iters", c["icp_iterations_per_scan"], "det", c["deterministic_replay"], "quality", c["trajectory_quality"]["gpu"] if c.get("trajectory_quality") else None)
    except Exception as e:
        print(tag, "failed", e)
PY
tail -3 gpurun_out/r2_b_team.err
