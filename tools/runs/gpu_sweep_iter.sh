#!/bin/bash
# GPU iteration on the sweep kernels: parity tests, then timings of the variants.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q > gpurun_out/pytest_sweep.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_sweep.log
tail -15 gpurun_out/pytest_sweep.log
for cfg in "64 2 10000 3" "64 2 10000 2" "32 2 20000 3" "32 2 20000 2" "16 2 20000 3" "16 2 20000 2" "48 2 20000 3" "96 4 4000 3" "112 4 4000 3" "112 4 400 1" "24 2 20000 3" "8 2 20000 3"; do
  timeout 120 python tools/prof_sweep.py $cfg 2>&1 | tail -1
done | tee gpurun_out/sweep_timings.log
