#!/bin/bash
# Quick GPU iteration on the streamed sweep kernel: parity of the r > 112 cases, then timings.
mkdir -p gpurun_out
K="streamed_kernel or 113 or 128 or 160 or 200 or 256 or 257 or 384 or 512"
timeout 400 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q -k "$K" > gpurun_out/pytest_stream.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest_stream.log)" | tee gpurun_out/stream_quick.log
for shape in "256 4 2072 3 5" "128 4 8880 3 5" "192 4 4144 3 3" "512 8 592 3 3"; do
  timeout 120 python tools/prof_sweep.py $shape 2>&1 | tail -1 | tee -a gpurun_out/stream_quick.log
done
