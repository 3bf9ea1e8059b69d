#!/bin/bash
mkdir -p gpurun_out
{
timeout 120 tools/lat_bench
export MF_LEFT_VER=3
for rm in 0 1; do
  export MF_LEFT_REMAP=$rm
  echo "== remap $rm"
  for a in "144 4 7400 5 7" "160 4 5920 5 7" "128 4 8880 5 7" "128 2 8880 5 7"; do
    timeout 120 python tools/prof_sweep.py $a 2>&1 | tail -1
  done
done
} > gpurun_out/lat.log 2>&1
cat gpurun_out/lat.log
