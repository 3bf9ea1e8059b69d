#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/prof_sweep.py 256 4 592 5 2 > gpurun_out/plain_left.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_left -s 2 -c 1 -o gpurun_out/prof_left_r256_v1 -f python tools/prof_sweep.py 256 4 592 5 2 > gpurun_out/ncu_left.log 2>&1
echo "ncu c128 rc=$?"; tail -3 gpurun_out/ncu_left.log
timeout 120 python tools/prof_sweep.py 256 4 888 5 2 f64 > gpurun_out/plain_left_f64.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_left -s 2 -c 1 -o gpurun_out/prof_left_r256_f64_v1 -f python tools/prof_sweep.py 256 4 888 5 2 f64 > gpurun_out/ncu_left_f64.log 2>&1
echo "ncu f64 rc=$?"; tail -3 gpurun_out/ncu_left_f64.log
