#!/bin/bash
# single-GPU check after the host-side full-order solver: config-1 driver example, the driver's own bench commands (both arms)
mkdir -p gpurun_out
timeout 100 python examples/rom_sweep.py --points 100 > gpurun_out/examples_s3.log 2>&1; echo "example rc=$?"; cut -c1-700 gpurun_out/examples_s3.log | head -3
SECONDS=0
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_s3.log 2> gpurun_out/bench_s3.err; echo "bench rc=$? wall ${SECONDS}s"
tail -c 400 gpurun_out/bench_s3.log; tail -2 gpurun_out/bench_s3.err
SECONDS=0
timeout 200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_s3.log 2>&1; echo "reference rc=$? wall ${SECONDS}s"
tail -c 300 gpurun_out/bench_ref_s3.log
