#!/bin/bash
# BASELINE configs[3], the error half: mean S-parameter error against the full-order sweep as a function of the basis size (thin geometry, host SuperLU yardstick)
mkdir -p gpurun_out
timeout 1200 python examples/basis_size_sweep.py --grid 12 6 400 --points 101 --first 3 --last 29 > gpurun_out/basis_size_sweep_n28800.jsonl 2> gpurun_out/basis_size_sweep.err
echo "rc=$?"; head -3 gpurun_out/basis_size_sweep_n28800.jsonl; tail -4 gpurun_out/basis_size_sweep_n28800.jsonl; tail -3 gpurun_out/basis_size_sweep.err
