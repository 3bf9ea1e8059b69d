#!/bin/bash
mkdir -p gpurun_out
{
for cfg in 3 4; do for ver in 2 3; do for rm in 0 1; do
  if [ $ver = 2 ] && [ $rm = 1 ]; then continue; fi
  echo "== cfg $cfg ver $ver remap $rm"
  MF_LEFT_CFG=$cfg MF_LEFT_VER=$ver MF_LEFT_REMAP=$rm timeout 120 python tools/prof_sweep.py 256 4 2960 5 7 2>&1 | tail -1
done; done; done
} > gpurun_out/lat2.log 2>&1
cat gpurun_out/lat2.log
