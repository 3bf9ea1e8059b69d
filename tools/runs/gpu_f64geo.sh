#!/bin/bash
mkdir -p gpurun_out
{
for cfg in 5 2; do
  echo "== f64 MF_LEFT_CFG=$cfg"
  for a in "160 4 88800 0 3 f64" "176 4 66600 0 3 f64" "192 4 66600 0 3 f64" "208 4 44400 0 3 f64"; do
    MF_LEFT_CFG=$cfg timeout 120 python tools/prof_sweep.py $a 2>&1 | tail -1
  done
done
echo "== f64 cfg1 vs cfg5 at r=128, 144"
for cfg in 1 5; do for a in "128 4 133200 0 3 f64" "144 4 88800 0 3 f64"; do MF_LEFT_CFG=$cfg timeout 120 python tools/prof_sweep.py $a 2>&1 | tail -1; done; done
} > gpurun_out/f64geo.log 2>&1
cat gpurun_out/f64geo.log
