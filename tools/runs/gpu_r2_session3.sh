#!/bin/bash
# single-GPU check after the host-side full-order solver: GPU tests, config-1 driver example, the driver's own bench commands
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_s3.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_s3.log
tail -3 gpurun_out/pytest_gpu_s3.log
timeout 100 python examples/rom_sweep.py --points 100 > gpurun_out/examples_s3.log 2>&1; echo "example rc=$?"; cut -c1-600 gpurun_out/examples_s3.log | head -3
/usr/bin/time -f "bench wall %e s" timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_s3.log 2> gpurun_out/bench_s3.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_s3.log; tail -2 gpurun_out/bench_s3.err
