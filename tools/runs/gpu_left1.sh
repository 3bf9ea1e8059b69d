#!/bin/bash
# first run of the left-looking sweep: parity tests, then timings against the right-looking kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sweep.py -m gpu -q -x > gpurun_out/pytest_sweep.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_sweep.log
tail -25 gpurun_out/pytest_sweep.log
(
for v in 4 5; do
  timeout 120 python tools/prof_sweep.py 256 4 2960 $v 5
  timeout 120 python tools/prof_sweep.py 128 4 8880 $v 5
  timeout 120 python tools/prof_sweep.py 160 4 5920 $v 5
  timeout 120 python tools/prof_sweep.py 512 8 592 $v 3
done
timeout 120 python tools/prof_sweep.py 64 2 20000 3 5
timeout 120 python tools/prof_sweep.py 64 2 20000 5 5
timeout 120 python tools/prof_sweep.py 96 2 10000 3 5
timeout 120 python tools/prof_sweep.py 96 2 10000 5 5
timeout 120 python tools/prof_sweep.py 256 4 2960 5 5 f64
timeout 120 python tools/prof_sweep.py 160 4 5920 5 5 f64
timeout 120 python tools/prof_sweep.py 128 4 8880 5 5 f64
timeout 120 python tools/prof_sweep.py 128 4 8880 3 5 f64
timeout 120 python tools/prof_sweep.py 512 8 1184 5 3 f64
MF_LEFT_CFG=3 timeout 120 python tools/prof_sweep.py 512 8 1184 5 3 f64
MF_LEFT_CFG=4 timeout 120 python tools/prof_sweep.py 512 8 592 5 3
) > gpurun_out/left_timings.log 2>&1
cat gpurun_out/left_timings.log
