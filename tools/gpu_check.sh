#!/bin/bash
# First-line GPU check: smoke, GPU parity tests, a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload small --steps 5 --warmup 3 > gpurun_out/bench_small.log 2>&1; echo "bench small rc=$?"
tail -5 gpurun_out/bench_small.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; echo "bench cfg2 rc=$?"
tail -5 gpurun_out/bench_cfg2.log
