#!/bin/bash
# cfg5 (N = 4M, r = 512, 8 ports, 1M points) strong-scaled on the GPUs of this box: bash tools/gpu_cfg5.sh NGPUS
np=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
free -g | head -2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 2953$np bench.py --gpus $np \
    --workload cfg5 --steps 1 --warmup 3 --no-alt-dtype --no-secondary > gpurun_out/bench_cfg5_n$np.log 2> gpurun_out/bench_cfg5_n$np.err
echo "cfg5 N=$np rc=$?"; tail -c 3000 gpurun_out/bench_cfg5_n$np.log; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/bench_cfg5_n$np.err | tail -5
