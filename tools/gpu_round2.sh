#!/bin/bash
# GPU parity tests, sweep timings (r = 64 blocked, f64 twin, streamed sizes), default bench and cfg3 bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for shape in "64 2 10000 3" "32 2 20000 3" "96 4 4000 3" "256 4 2072 3 5" "128 4 8880 3 5"; do
  timeout 120 python tools/prof_sweep.py $shape 2>&1 | tail -1
done | tee gpurun_out/sweep_timings.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; echo "bench cfg2 rc=$?"; tail -c 300 gpurun_out/bench_cfg2.log
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 > gpurun_out/bench_cfg3_n1.log 2>&1; echo "bench cfg3 rc=$?"; tail -c 300 gpurun_out/bench_cfg3_n1.log
