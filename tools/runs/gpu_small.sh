#!/bin/bash
# GPU check of the r x r kernels (fused Cholesky + inverse, Jacobi variants): tests, then timings.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_path.py tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/pytest_small.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest_small.log)" | tee gpurun_out/small.log
for r in 128 256 512; do
  timeout 120 python tools/prof_small.py $r 2>&1 | tail -1
  MF_JACOBI_PAIR_KERNEL=1 timeout 120 python tools/prof_small.py $r 2>&1 | tail -1 | sed 's/^/pair kernel: /'
done | tee -a gpurun_out/small.log
timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-alt-dtype > gpurun_out/b3.log 2>&1; echo rc=$?
python - <<'P'
import json
d=json.loads(open("gpurun_out/b3.log").read().strip().splitlines()[-1]); print(d["ms_per_step"], d["basis_plus_projection"]["ms"], d["basis_plus_projection"]["frac"], d["kernels"]["jacobi_svd"])
P
