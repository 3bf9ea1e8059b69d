#!/bin/bash
# one ncu --set full capture of a sweep kernel: bash tools/gpu_ncu_one.sh NAME "R M F VARIANT REPS [f64]" [ENV=VAL ...]
name=$1; args=$2; shift 2
mkdir -p gpurun_out
for kv in "$@"; do export "$kv"; done
timeout 120 python tools/prof_sweep.py $args > gpurun_out/plain_$name.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_left -s 2 -c 1 -o gpurun_out/prof_$name -f python tools/prof_sweep.py $args > gpurun_out/ncu_$name.log 2>&1
echo "ncu $name rc=$?"; tail -1 gpurun_out/plain_$name.log; tail -2 gpurun_out/ncu_$name.log
