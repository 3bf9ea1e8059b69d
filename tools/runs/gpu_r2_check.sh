#!/bin/bash
# Round-2 GPU check: smoke, GPU parity tests, the default bench (north-star config, N=1) and a short reference arm.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3_n1.log 2> gpurun_out/bench_cfg3_n1.err; echo "bench cfg3 rc=$?"
tail -c 3000 gpurun_out/bench_cfg3_n1.log; tail -5 gpurun_out/bench_cfg3_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_cfg3.log 2>&1; echo "reference rc=$?"
tail -c 600 gpurun_out/bench_reference_cfg3.log
