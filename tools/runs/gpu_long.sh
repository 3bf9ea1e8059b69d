#!/bin/bash
mkdir -p gpurun_out
{
for ver in 3 2; do
  for F in 2960 29600; do
    echo "== MF_LEFT_VER=$ver F=$F with phase clocks"
    MF_LEFT_VER=$ver MF_LEFT_TIMING=1 timeout 120 python tools/prof_sweep.py 256 4 $F 5 2 2>&1 | grep -E "TIMING\] r=|pts/s" | tail -2
  done
done
echo "== clocks/power while v3 runs F=118400 (about 0.55 s per launch, 6 launches)"
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv,noheader -lms 100 > gpurun_out/smi_long.csv &
SMI=$!
MF_LEFT_VER=3 timeout 200 python tools/prof_sweep.py 256 4 118400 5 4 2>&1 | tail -1
MF_LEFT_VER=2 timeout 200 python tools/prof_sweep.py 256 4 118400 5 4 2>&1 | tail -1
kill $SMI
sort gpurun_out/smi_long.csv | uniq -c | sort -rn | head -12
} > gpurun_out/long.log 2>&1
cat gpurun_out/long.log
