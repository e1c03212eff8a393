set -u
KB_ICP_TEAM_Q=64 timeout -k 10 200 python tools/icp_timeline.py 100 2 2>&1 | tail -12
