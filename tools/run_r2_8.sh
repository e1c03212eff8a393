set -u
timeout -k 10 200 python tools/icp_timeline.py 100 2 2>&1 | grep "iters\|fill pass" | tail -4
timeout -k 10 200 python tools/queue_timeline.py 100 40 2>&1 | tail -3
