#!/bin/bash
export MF_LEFT_LOOKAHEAD=1 MF_LEFT_TIMING=1
for a in "256 4 592 5 1" "256 4 592 5 1 f64" "512 8 296 5 1" "128 4 1184 5 1"; do
  timeout 120 python tools/prof_sweep.py $a 2>&1 | grep -E "TIMING|variant" | tail -2
done
