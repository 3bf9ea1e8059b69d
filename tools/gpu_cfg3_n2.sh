#!/bin/bash
# strong-scaled default bench (cfg3) on 2 GPUs, final code
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 \
    > gpurun_out/bench_cfg3_n2.log 2> gpurun_out/bench_cfg3_n2.err
echo "bench N=2 rc=$?"; tail -c 1500 gpurun_out/bench_cfg3_n2.log; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/bench_cfg3_n2.err | tail -3
