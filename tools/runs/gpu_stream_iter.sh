#!/bin/bash
# GPU iteration on the streamed sweep kernel: parity of every geometry (MF_STREAM_CFG) with the fused first step
# (MF_STREAM_FUSE), timings at r = 256 / 128 / 512, then the ncu launch list of one cfg3 bench step.
mkdir -p gpurun_out
: > gpurun_out/stream_iter.log
K="streamed_kernel or 113 or 128 or 160 or 200 or 256 or 257 or 384 or 512"
for cf in "0 1" "1 0" "1 1" "2 1" "3 1"; do
  set -- $cf
  MF_STREAM_CFG=$1 MF_STREAM_FUSE=$2 timeout 400 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q -k "$K" > gpurun_out/pytest_stream_$1_$2.log 2>&1
  echo "cfg=$1 fuse=$2 pytest rc=$? $(tail -1 gpurun_out/pytest_stream_$1_$2.log)" | tee -a gpurun_out/stream_iter.log
done
for cfg in 0 1 2 3; do for fuse in 0 1; do
  for shape in "256 4 2072 3 5" "128 4 8000 3 5"; do
    MF_STREAM_CFG=$cfg MF_STREAM_FUSE=$fuse timeout 120 python tools/prof_sweep.py $shape 2>&1 | tail -1 | sed "s/^/cfg=$cfg fuse=$fuse /" | tee -a gpurun_out/stream_iter.log
  done
done; done
for fuse in 0 1; do
  MF_STREAM_FUSE=$fuse timeout 120 python tools/prof_sweep.py 512 8 592 3 3 2>&1 | tail -1 | sed "s/^/fuse=$fuse /" | tee -a gpurun_out/stream_iter.log
done
timeout 300 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-alt-dtype > gpurun_out/plain_bench_cfg3.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-alt-dtype > gpurun_out/ncu_bench3.log 2>&1
echo "cfg3 launch list rc=$?"; tail -c 600 gpurun_out/plain_bench_cfg3.log
