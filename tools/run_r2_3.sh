set -u
mkdir -p gpurun_out
echo "== test 39 (team)"
timeout -k 10 150 python -m pytest tests/test_gpu_parity.py -q -x -k "long_stream" --timeout=120 2>&1 | tail -15
echo "== test 39 (legacy)"
KB_ICP_TEAM_Q=0 timeout -k 10 150 python -m pytest tests/test_gpu_parity.py -q -x -k "long_stream" --timeout=120 2>&1 | tail -5
echo "== remaining tests"
timeout -k 10 400 python -m pytest tests -m gpu -q --timeout=120 --deselect tests/test_gpu_parity.py::test_pipeline_long_stream_with_table_rebuilds 2>&1 | tail -15
