#!/bin/bash
# GPU check: smoke, GPU parity tests, a short bench, then the ncu launch list of the same bench command.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; echo "bench cfg2 rc=$?"
tail -c 1500 gpurun_out/bench_cfg2.log
if [ "$1" == "ncu" ]; then
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu launch list rc=$?"; tail -2 gpurun_out/launches.csv
fi
