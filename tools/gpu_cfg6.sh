#!/bin/bash
mkdir -p gpurun_out
{
MF_LEFT_CFG=6 timeout 600 python -m pytest tests/test_gpu_sweep.py -x -q 2>&1 | tail -3
export MF_LEFT_CFG=6
for ns in 4 2; do for rm in 1 0; do
  echo "== cfg6 nstage $ns remap $rm"
  for a in "256 4 2960 5 7" "192 4 4440 5 7"; do
    MF_LEFT_NSTAGE=$ns MF_LEFT_REMAP=$rm timeout 120 python tools/prof_sweep.py $a 2>&1 | tail -1
  done
done; done
MF_LEFT_TIMING=1 timeout 120 python tools/prof_sweep.py 256 4 148 5 1 2>&1 | grep TIMING
MF_LEFT_TIMING=1 MF_LEFT_REMAP=0 timeout 120 python tools/prof_sweep.py 256 4 148 5 1 2>&1 | grep TIMING
} > gpurun_out/cfg6.log 2>&1
cat gpurun_out/cfg6.log
