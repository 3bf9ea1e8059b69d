#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_sweep.py -x -q 2>&1 | tail -3
for a in "256 4 44400 0 3 f64" "192 4 66600 0 3 f64" "160 4 88800 0 3 f64" "128 4 133200 0 3 f64" "96 2 100000 0 3 f64" "512 8 2368 0 3 f64" "256 4 29600 0 3"; do
  timeout 120 python tools/prof_sweep.py $a 2>&1 | tail -1
done
} > gpurun_out/f64fifo.log 2>&1
cat gpurun_out/f64fifo.log
