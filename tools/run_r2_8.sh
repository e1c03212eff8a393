set -u
for r in 0.2 0.1 0.05; do
echo "== R=$r"
KB_ICP_RADIUS=$r timeout -k 10 200 python tools/icp_timeline.py 100 2 2>&1 | grep -v "map update\|per member" | tail -8
KB_ICP_RADIUS=$r timeout -k 10 200 python tools/queue_timeline.py 100 40 2>&1 | tail -3
done
