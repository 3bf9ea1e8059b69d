#!/bin/bash
mkdir -p gpurun_out
export MF_LEFT_LOOKAHEAD=1
timeout 600 python -m pytest tests/test_gpu_sweep.py -m gpu -q -x > gpurun_out/pytest_sweep3.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_sweep3.log
tail -15 gpurun_out/pytest_sweep3.log
(
timeout 120 python tools/prof_sweep.py 256 4 2960 5 5
env -u MF_LEFT_LOOKAHEAD timeout 120 python tools/prof_sweep.py 256 4 2960 5 5
timeout 120 python tools/prof_sweep.py 128 4 8880 5 5
timeout 120 python tools/prof_sweep.py 160 4 5920 5 5
timeout 120 python tools/prof_sweep.py 192 4 2960 5 5
timeout 120 python tools/prof_sweep.py 512 8 592 5 3
env -u MF_LEFT_LOOKAHEAD timeout 120 python tools/prof_sweep.py 512 8 592 5 3
timeout 120 python tools/prof_sweep.py 96 2 10000 5 5
timeout 120 python tools/prof_sweep.py 256 4 2960 5 5 f64
env -u MF_LEFT_LOOKAHEAD timeout 120 python tools/prof_sweep.py 256 4 2960 5 5 f64
timeout 120 python tools/prof_sweep.py 160 4 5920 5 5 f64
timeout 120 python tools/prof_sweep.py 128 4 8880 5 5 f64
timeout 120 python tools/prof_sweep.py 512 8 1184 5 3 f64
MF_LEFT_CFG=3 timeout 120 python tools/prof_sweep.py 512 8 1184 5 3 f64
MF_LEFT_CFG=3 timeout 120 python tools/prof_sweep.py 256 4 2960 5 5
) > gpurun_out/left3_timings.log 2>&1
cat gpurun_out/left3_timings.log
timeout 120 python tools/prof_sweep.py 256 4 592 5 2 > gpurun_out/plain_left3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_left -s 2 -c 1 -o gpurun_out/prof_left_r256_v3 -f python tools/prof_sweep.py 256 4 592 5 2 > gpurun_out/ncu_left3.log 2>&1
echo "ncu c128 rc=$?"; tail -2 gpurun_out/ncu_left3.log
