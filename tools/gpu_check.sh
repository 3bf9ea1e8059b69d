#!/bin/bash
# GPU check: smoke, GPU parity tests, a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; echo "bench cfg2 rc=$?"
tail -3 gpurun_out/bench_cfg2.log
