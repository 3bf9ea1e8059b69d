#!/bin/bash
mkdir -p gpurun_out
export MF_LEFT_V2=1
timeout 900 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/pytest_sweep.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_sweep.log
tail -25 gpurun_out/pytest_sweep.log
(
timeout 120 python tools/prof_sweep.py 256 4 2960 5 5
env -u MF_LEFT_V2 timeout 120 python tools/prof_sweep.py 256 4 2960 5 5
timeout 120 python tools/prof_sweep.py 128 4 8880 5 5
timeout 120 python tools/prof_sweep.py 160 4 5920 5 5
timeout 120 python tools/prof_sweep.py 512 8 592 5 3
timeout 120 python tools/prof_sweep.py 96 2 10000 5 5
timeout 120 python tools/prof_sweep.py 256 4 2960 5 5 f64
timeout 120 python tools/prof_sweep.py 160 4 5920 5 5 f64
timeout 120 python tools/prof_sweep.py 128 4 8880 5 5 f64
timeout 120 python tools/prof_sweep.py 512 8 1184 5 3 f64
timeout 120 python tools/prof_sweep.py 64 2 20000 5 5 f64
timeout 120 python tools/prof_sweep.py 64 2 20000 3 5 f64
) > gpurun_out/left2_timings.log 2>&1
cat gpurun_out/left2_timings.log
(
timeout 200 python tools/prof_spmm.py 256 2000 5 c128 25 20
MF_SPMM_SPLIT=0 timeout 200 python tools/prof_spmm.py 256 2000 5 c128 25 20
timeout 200 python tools/prof_spmm.py 256 2000 5 f64 25 20
MF_SPMM_SPLIT=0 timeout 200 python tools/prof_spmm.py 256 2000 5 f64 25 20
timeout 200 python tools/prof_spmm.py 64 1000 5 c128
) > gpurun_out/spmm_timings.log 2>&1
cat gpurun_out/spmm_timings.log
timeout 120 python tools/prof_sweep.py 256 4 592 5 2 > gpurun_out/plain_left.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_left -s 2 -c 1 -o gpurun_out/prof_left_r256_v2 -f python tools/prof_sweep.py 256 4 592 5 2 > gpurun_out/ncu_left.log 2>&1
echo "ncu c128 rc=$?"; tail -2 gpurun_out/ncu_left.log
