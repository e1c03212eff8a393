set -u
timeout -k 10 300 python tools/multi_stream.py 1 2 4 8 2>&1 | tail -4
KB_ICP_SMEM_KB=160 timeout -k 10 300 python tools/multi_stream.py 4 8 2>&1 | tail -2
KB_ICP_TEAM_Q=0 timeout -k 10 300 python tools/multi_stream.py 4 8 2>&1 | tail -2
