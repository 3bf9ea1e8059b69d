#!/bin/bash
mkdir -p gpurun_out
{
MF_LEFT_CFG=8 timeout 300 python -m pytest tests/test_gpu_sweep.py -x -q -k "left or variants or real_sweep" 2>&1 | tail -2
MF_LEFT_CFG=8 MF_LEFT_TIMING=1 timeout 120 python tools/prof_sweep.py 256 4 5920 5 1 2>&1 | grep -E "TIMING|pts/s" | tail -2
for a in "256 4 29600 5 3" "192 4 44400 5 3" "128 4 88800 5 3"; do
  MF_LEFT_CFG=8 timeout 120 python tools/prof_sweep.py $a 2>&1 | tail -1
done
} > gpurun_out/v4t.log 2>&1
cat gpurun_out/v4t.log
