#!/bin/bash
mkdir -p gpurun_out
export MF_LEFT_VER=3
timeout 600 python -m pytest tests/test_gpu_sweep.py -m gpu -q -x > gpurun_out/pytest_sweep4.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_sweep4.log
(
timeout 120 python tools/prof_sweep.py 256 4 2960 5 5
timeout 120 python tools/prof_sweep.py 224 4 2960 5 5
timeout 120 python tools/prof_sweep.py 160 4 5920 5 5
timeout 120 python tools/prof_sweep.py 128 4 8880 5 5
timeout 120 python tools/prof_sweep.py 96 2 10000 5 5
timeout 120 python tools/prof_sweep.py 512 8 592 5 3
timeout 120 python tools/prof_sweep.py 256 4 2960 5 5 f64
timeout 120 python tools/prof_sweep.py 160 4 5920 5 5 f64
timeout 120 python tools/prof_sweep.py 512 8 1184 5 3 f64
) > gpurun_out/left4_timings.log 2>&1
cat gpurun_out/left4_timings.log
export MF_LEFT_TIMING=1
timeout 120 python tools/prof_sweep.py 256 4 592 5 1 2>&1 | grep TIMING
