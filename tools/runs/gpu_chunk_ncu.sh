#!/bin/bash
# evidence for the chunked look-ahead default: one ncu --set full capture of the kernel (r = 256, m = 4, 2 points per CTA = one launch of the
# chunked sequence), then the launch list of the default bench command
mkdir -p gpurun_out
timeout 40 python tools/prof_sweep.py 256 4 592 0 5 > gpurun_out/plain_left3_chunk.log 2>&1 && \
timeout 70 ncu --set full --clock-control none --import-source on -k regex:sweep_left -s 2 -c 1 -o gpurun_out/prof_left3_chunk_r256 -f python tools/prof_sweep.py 256 4 592 0 5 > gpurun_out/ncu_left3_chunk.log 2>&1
echo "ncu full rc=$?"; tail -1 gpurun_out/plain_left3_chunk.log
timeout 70 python bench.py --steps 2 --warmup 3 --no-parity --no-secondary --no-alt-dtype > gpurun_out/plain_bench_s3.log 2>&1 && \
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_cfg3_s3.csv \
    python bench.py --steps 2 --warmup 3 --no-parity --no-secondary --no-alt-dtype > gpurun_out/ncu_bench_s3.log 2>&1
echo "launch list rc=$?"; tail -1 gpurun_out/launches_cfg3_s3.csv | cut -c1-200
