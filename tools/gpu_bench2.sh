#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2.log 2>&1; echo "bench cfg2 rc=$?"
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_n1.log 2>&1; echo "bench cfg3 rc=$?"
python - <<'P'
import json
for f in ['gpurun_out/bench_cfg2.log','gpurun_out/bench_cfg3_n1.log']:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value',round(d['value']), 'ms/step',round(d['ms_per_step'],3), 'e2e',round(d['e2e']['value']), d['stages'])
    for k,v in d['kernels'].items(): print('   ',k, v['calls_per_step'], round(v['ms_per_step'],3))
P
