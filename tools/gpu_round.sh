#!/bin/bash
# Full GPU check of the current tree: smoke, all GPU parity tests, geometry crossover of the streamed sweep, benches
# (cfg2 default, cfg3), launch list of the cfg3 bench and one ncu --set full capture of the streamed sweep kernel.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
: > gpurun_out/stream_xover.log
for r in 160 192 224; do for cfg in 1 2; do
  MF_STREAM_CFG=$cfg timeout 120 python tools/prof_sweep.py $r 4 4144 3 3 2>&1 | tail -1 | sed "s/^/cfg=$cfg /" | tee -a gpurun_out/stream_xover.log
done; done
for f in 100 148 296 592; do   # does the working set (slots x 1.08 MB) falling out of L2 matter?
  timeout 120 python tools/prof_sweep.py 256 4 $f 3 5 2>&1 | tail -1 | tee -a gpurun_out/stream_xover.log
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; echo "bench cfg2 rc=$?"; tail -c 600 gpurun_out/bench_cfg2.log
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 > gpurun_out/bench_cfg3_n1.log 2>&1; echo "bench cfg3 rc=$?"; tail -c 600 gpurun_out/bench_cfg3_n1.log
timeout 300 python tools/prof_sweep.py 256 4 1036 3 3 > gpurun_out/plain_stream.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_stream -c 1 -s 2 -f -o gpurun_out/prof_stream_r256_v4 \
    python tools/prof_sweep.py 256 4 1036 3 1 > gpurun_out/ncu_stream.log 2>&1
echo "ncu stream rc=$?"; cat gpurun_out/plain_stream.log | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-alt-dtype > gpurun_out/ncu_bench3.log 2>&1
echo "cfg3 launch list rc=$?"
