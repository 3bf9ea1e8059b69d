#!/bin/bash
# all GPU tests with the final choices, then the basis-size sweep of BASELINE configs[3] at N = 1M (r = 16 .. 512, stages 1+2 and the sweep)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
for r in 16 32 64 128 512; do
  pts=100000; if [ $r -ge 512 ]; then pts=20000; fi
  timeout 600 python bench.py --r $r --points $pts --steps 5 --warmup 3 --no-secondary --parity isolated > gpurun_out/bench_cfg4_r$r.log 2> gpurun_out/bench_cfg4_r$r.err; echo "cfg4 r=$r rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cfg4_r$r.log').read().strip().splitlines()[-1])
print('r=$r value %.4g ms/step %.2f sweep/gpu %.4g bp %.2f ms frac %.2f alt(f64) %.4g parity %s' % (d['value'], d['ms_per_step'], d['stages']['sweep_kernel_points_per_s_per_gpu'], d['basis_plus_projection']['ms'], d['basis_plus_projection']['frac'], d['other_dtype']['value'], d['parity']['ok']))
PY
done
