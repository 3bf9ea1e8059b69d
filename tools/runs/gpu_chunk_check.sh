#!/bin/bash
# after making short launches the default of the left-looking sweep: all GPU tests, then the driver's bench command
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_s3b.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_s3b.log
tail -4 gpurun_out/pytest_gpu_s3b.log
SECONDS=0
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_s3b.log 2> gpurun_out/bench_s3b.err; echo "bench rc=$? wall ${SECONDS}s"
tail -c 300 gpurun_out/bench_s3b.log; tail -2 gpurun_out/bench_s3b.err
