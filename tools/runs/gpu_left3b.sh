#!/bin/bash
mkdir -p gpurun_out
export MF_LEFT_VER=3
timeout 600 python -m pytest tests/test_gpu_sweep.py -m gpu -q -x > gpurun_out/pytest_sweep3.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_sweep3.log
tail -8 gpurun_out/pytest_sweep3.log
(
timeout 120 python tools/prof_sweep.py 256 4 2960 5 5
timeout 120 python tools/prof_sweep.py 128 4 8880 5 5
timeout 120 python tools/prof_sweep.py 160 4 5920 5 5
timeout 120 python tools/prof_sweep.py 512 8 592 5 3
timeout 120 python tools/prof_sweep.py 96 2 10000 5 5
timeout 120 python tools/prof_sweep.py 64 2 20000 5 5
timeout 120 python tools/prof_sweep.py 256 4 2960 5 5 f64
timeout 120 python tools/prof_sweep.py 160 4 5920 5 5 f64
timeout 120 python tools/prof_sweep.py 128 4 8880 5 5 f64
timeout 120 python tools/prof_sweep.py 512 8 1184 5 3 f64
timeout 120 python tools/prof_sweep.py 64 2 20000 5 5 f64
) > gpurun_out/left3_timings.log 2>&1
cat gpurun_out/left3_timings.log
export MF_LEFT_TIMING=1
for a in "256 4 592 5 1" "256 4 888 5 1 f64"; do
  timeout 120 python tools/prof_sweep.py $a 2>&1 | grep -E "TIMING" | tail -1
done
unset MF_LEFT_TIMING
(
for v in 2 3; do
  MF_LEFT_VER=$v MF_LEFT_CFG=5 timeout 120 python tools/prof_sweep.py 256 4 2960 5 5 f64
  MF_LEFT_VER=$v MF_LEFT_CFG=5 timeout 120 python tools/prof_sweep.py 160 4 5920 5 5 f64
  MF_LEFT_VER=$v MF_LEFT_CFG=5 timeout 120 python tools/prof_sweep.py 128 4 8880 5 5 f64
  MF_LEFT_VER=$v MF_LEFT_CFG=1 timeout 120 python tools/prof_sweep.py 128 4 8880 5 5 f64
  MF_LEFT_VER=$v MF_LEFT_CFG=2 timeout 120 python tools/prof_sweep.py 128 4 8880 5 5
  MF_LEFT_VER=$v MF_LEFT_CFG=2 timeout 120 python tools/prof_sweep.py 96 2 10000 5 5
  MF_LEFT_VER=$v timeout 120 python tools/prof_sweep.py 224 4 2960 5 5
  MF_LEFT_VER=$v timeout 120 python tools/prof_sweep.py 384 4 1184 5 3
done
) > gpurun_out/left_geom.log 2>&1
cat gpurun_out/left_geom.log
