#!/bin/bash
# 2-GPU check: multi-rank parity tests (NCCL), then the strong-scaled bench at N=2 (graph replay of the multi-rank step, then eager)
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x -s > gpurun_out/pytest_multirank.log 2>&1; echo "pytest multirank rc=$?" | tee -a gpurun_out/pytest_multirank.log
tail -25 gpurun_out/pytest_multirank.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_cfg3_n2.log 2> gpurun_out/bench_cfg3_n2.err; echo "bench N=2 rc=$?"
tail -c 2500 gpurun_out/bench_cfg3_n2.log; tail -5 gpurun_out/bench_cfg3_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-graph --no-alt-dtype > gpurun_out/bench_cfg3_n2_eager.log 2> gpurun_out/bench_cfg3_n2_eager.err; echo "bench N=2 eager rc=$?"
tail -c 800 gpurun_out/bench_cfg3_n2_eager.log; tail -3 gpurun_out/bench_cfg3_n2_eager.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --workload cfg2 --no-alt-dtype > gpurun_out/bench_cfg2_n2.log 2> gpurun_out/bench_cfg2_n2.err; echo "bench cfg2 N=2 rc=$?"
tail -c 800 gpurun_out/bench_cfg2_n2.log; tail -3 gpurun_out/bench_cfg2_n2.err
