#!/bin/bash
# 8-GPU run: 4-rank parity tests, strong-scaled cfg3 at 8 and 4 GPUs, cfg5 at 8 GPUs, cfg2 at 8 GPUs
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { # name nproc port args...
  name=$1; np=$2; port=$3; shift 3
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np "$@" \
      > gpurun_out/$name.log 2> gpurun_out/$name.err
  echo "$name rc=$?"; tail -c 400 gpurun_out/$name.log; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/$name.err | tail -3
}
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x -s -k "nccl" > gpurun_out/pytest_multirank8.log 2>&1; echo "pytest multirank rc=$?"; tail -12 gpurun_out/pytest_multirank8.log
run bench_cfg3_n8 8 29521 --steps 10 --warmup 3
run bench_cfg3_n4 4 29522 --steps 10 --warmup 3 --no-alt-dtype
run bench_cfg5_n8 8 29523 --workload cfg5 --steps 2 --warmup 3 --no-alt-dtype
run bench_cfg2_n8 8 29524 --workload cfg2 --steps 20 --warmup 3 --no-alt-dtype
