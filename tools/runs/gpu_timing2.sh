#!/bin/bash
export MF_LEFT_VER=3 MF_LEFT_TIMING=1
for a in "256 4 148 5 1" "256 4 296 5 1" "256 4 148 5 1 f64" "256 4 444 5 1 f64"; do
  timeout 120 python tools/prof_sweep.py $a 2>&1 | grep -E "TIMING" | tail -2
done
