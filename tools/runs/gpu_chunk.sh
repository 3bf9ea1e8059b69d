#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/chunk_sweep.py > gpurun_out/chunk_sweep.log 2>&1; echo "rc=$?"
cat gpurun_out/chunk_sweep.log | tail -70
