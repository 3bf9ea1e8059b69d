#!/bin/bash
# final single-GPU check of the round: smoke, all GPU tests, default bench (cfg3), reference arm, ncu launch list of the bench command
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg3_n1.log 2> gpurun_out/bench_cfg3_n1.err; echo "bench cfg3 rc=$?"
tail -c 1200 gpurun_out/bench_cfg3_n1.log; tail -3 gpurun_out/bench_cfg3_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_cfg3.log 2>&1; echo "reference rc=$?"
tail -c 600 gpurun_out/bench_reference_cfg3.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-parity --no-secondary --no-alt-dtype > gpurun_out/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_cfg3.csv \
    python bench.py --steps 2 --warmup 3 --no-parity --no-secondary --no-alt-dtype > gpurun_out/ncu_bench.log 2>&1
echo "ncu launch list rc=$?"; tail -2 gpurun_out/launches_cfg3.csv
