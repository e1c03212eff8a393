set -u
echo "== nn A/B (checksums must agree)"
timeout -k 10 200 python tools/nn_ab.py 2>&1 | tail -1 | sed 's/^/bulk2: /'
KB_NN_KERNEL=regs timeout -k 10 200 python tools/nn_ab.py 2>&1 | tail -1 | sed 's/^/regs:  /'
echo "== nn parity tests"
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -q -x -k "closest or compact" 2>&1 | tail -3
