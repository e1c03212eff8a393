#!/bin/bash
# r2 evidence: launch list of a bench run, ncu --set full of two k_register_frame launches (queued stream) and of the
# NN kernel, then the un-profiled bench (both arms)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
K='regex:k_register_frame|k_map_|k_nn_|k_icp|k_downsample|k_preprocess'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_r2.csv \
  python bench.py --steps 20 --warmup 5 --repeats 3 --prime 30 --no-nn --no-cpu > gpurun_out/launches_r2_bench.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_register_frame -s 34 -c 2 -f -o gpurun_out/prof_r2_frame \
  python tools/profile_target.py 40 queued > gpurun_out/prof_r2_frame.log 2>&1
echo "frame rc=$?"; tail -2 gpurun_out/prof_r2_frame.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_nn_query -s 2 -c 1 -f -o gpurun_out/prof_r2_nn \
  python tools/profile_nn.py > gpurun_out/prof_r2_nn.log 2>&1
echo "nn rc=$?"; tail -2 gpurun_out/prof_r2_nn.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
echo "ref rc=$?"; tail -c 600 gpurun_out/r2_bench_reference.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r2_bench_default.json
ls -la gpurun_out/*.ncu-rep
