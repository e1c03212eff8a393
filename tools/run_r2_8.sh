set -u
for q in 64 40; do
echo "== q=$q"
KB_ICP_TEAM_Q=$q timeout -k 10 200 python tools/icp_timeline.py 100 2 2>&1 | tail -10
KB_ICP_TEAM_Q=$q timeout -k 10 200 python tools/queue_timeline.py 100 40 2>&1 | tail -3
done
