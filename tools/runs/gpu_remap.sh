#!/bin/bash
# warp-role remap of the look-ahead sweep body (panel warps on SM sub-partitions 0/1) + per-stage clocks of a pivot column step
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_gpu_sweep.py -x -q -k "left or variants" 2>&1 | tail -2
export MF_LEFT_VER=3
for rm in 0 1; do
  export MF_LEFT_REMAP=$rm
  echo "== remap $rm"
  for a in "256 4 2960 5 7" "192 4 4440 5 7" "128 4 8880 5 7" "512 8 592 5 5"; do
    timeout 120 python tools/prof_sweep.py $a 2>&1 | tail -1
  done
done
export MF_LEFT_REMAP=1
timeout 300 python -m pytest tests/test_gpu_sweep.py -x -q -k "left or variants" 2>&1 | tail -2
cp morfem_b200/csrc/build/libmorfem_b200_clk.so morfem_b200/libmorfem_b200.so
export MF_LEFT_TIMING=1
for rm in 0 1; do
  export MF_LEFT_REMAP=$rm
  echo "== clocks, remap $rm"
  for a in "256 4 296 5 1" "256 4 148 5 1" "256 4 444 5 1 f64"; do
    MF_LEFT_VER=3 timeout 120 python tools/prof_sweep.py $a 2>&1 | grep -E "TIMING|CLOCKS" | tail -3
  done
done
echo "== clocks, v2 body"
MF_LEFT_VER=2 timeout 120 python tools/prof_sweep.py 256 4 296 5 1 2>&1 | grep -E "CLOCKS" | tail -1
MF_LEFT_VER=2 timeout 120 python tools/prof_sweep.py 256 4 444 5 1 f64 2>&1 | grep -E "CLOCKS" | tail -1
} > gpurun_out/remap.log 2>&1
cat gpurun_out/remap.log
