#!/bin/bash
mkdir -p gpurun_out
{
echo "== single launch sanity"
MF_LEFT_CFG=8 timeout 60 python tools/prof_sweep.py 64 2 100 5 1 2>&1 | tail -2
MF_LEFT_CFG=8 timeout 60 python tools/prof_sweep.py 256 4 600 5 1 2>&1 | tail -2
echo "== pytest with cfg 8"
MF_LEFT_CFG=8 timeout 600 python -m pytest tests/test_gpu_sweep.py -x -q 2>&1 | tail -5
echo "== timings cfg 8"
for a in "256 4 2960 5 5" "256 4 29600 5 3" "256 8 29600 5 3" "224 4 29600 5 3" "192 4 44400 5 3" "160 4 59200 5 3" "128 4 88800 5 3" "96 2 100000 5 3"; do
  MF_LEFT_CFG=8 timeout 120 python tools/prof_sweep.py $a 2>&1 | tail -1
done
} > gpurun_out/v4.log 2>&1
cat gpurun_out/v4.log
