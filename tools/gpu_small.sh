#!/bin/bash
# GPU check of the r x r kernels (fused Cholesky + inverse, Jacobi with the flag barrier): tests, then timings.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_path.py tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/pytest_small.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest_small.log)" | tee gpurun_out/small.log
for r in 64 128 256 512; do timeout 120 python tools/prof_small.py $r 2>&1 | tail -3; done | tee -a gpurun_out/small.log
