#!/bin/bash
# First GPU call of the next round: the measurements DESIGN.md section 6a asks for (each line is one decision).
mkdir -p gpurun_out
: > gpurun_out/next.log
# (1) 64 < r <= 112: shared-memory kernel (one 8-warp CTA per SM) against the streamed kernel's three-CTA geometry
for r in 72 80 96 112; do
  timeout 120 python tools/prof_sweep.py $r 4 8880 3 3 2>&1 | tail -1 | sed "s/^/blocked /" | tee -a gpurun_out/next.log
  MF_SWEEP_FORCE_STREAM=1 timeout 120 python tools/prof_sweep.py $r 4 8880 3 3 2>&1 | tail -1 | sed "s/^/stream  /" | tee -a gpurun_out/next.log
done
# (2) is the e2e mean free of outliers after the allocator warm-up?  (ms_per_step vs ms_per_call_median / max)
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/next.log
python - <<'P' | tee -a gpurun_out/next.log
import json
d = json.loads(open("gpurun_out/bench_cfg2.log").read().strip().splitlines()[-1]); e = d["e2e"]
print("step ms", d["ms_per_step"], "e2e mean/median/max ms", e["ms_per_step"], e["ms_per_call_median"], e["ms_per_call_max"], "resident", e["operators_resident"]["ms_per_step"])
P
# (3) r x r kernels after this round's changes
for r in 64 128 256 512; do timeout 120 python tools/prof_small.py $r 2>&1 | tail -3; done | tee -a gpurun_out/next.log
