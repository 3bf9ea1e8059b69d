#!/bin/bash
# Final evidence of a round: smoke, all GPU tests, default bench (cfg2), cfg3 bench, reference arm, ncu launch lists of both
# bench commands and one ncu --set full capture of the streamed sweep kernel (r = 256).
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; echo "bench cfg2 rc=$?"
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 > gpurun_out/bench_cfg3_n1.log 2>&1; echo "bench cfg3 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "bench reference rc=$?"; tail -c 400 gpurun_out/bench_reference.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-alt-dtype > gpurun_out/ncu_bench.log 2>&1; echo "cfg2 launch list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-alt-dtype > gpurun_out/ncu_bench3.log 2>&1; echo "cfg3 launch list rc=$?"
timeout 300 python tools/prof_sweep.py 256 4 2072 3 3 > gpurun_out/plain_stream.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_stream -c 1 -s 2 -f -o gpurun_out/prof_stream_r256_v5 \
    python tools/prof_sweep.py 256 4 2072 3 1 > gpurun_out/ncu_stream.log 2>&1
echo "ncu stream rc=$?"; tail -1 gpurun_out/plain_stream.log
