#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python examples/rom_sweep.py --points 40
timeout 300 python examples/rom_sweep.py --points 40 --replicate 3
python - <<'PY'
import numpy as np, os, sys
sys.path.insert(0, ".")
from morfem_b200 import data_io, synthetic
ct, tt = synthetic.waveguide_operators(9, 1, 379)
data_io.save_operators("/tmp/mf_data", ct, tt, synthetic.shipped_port_matrix())
print("saved", sorted(os.listdir("/tmp/mf_data")), sum(os.path.getsize(os.path.join("/tmp/mf_data", f)) for f in os.listdir("/tmp/mf_data")), "bytes")
PY
timeout 300 python examples/rom_sweep.py --points 40 --data /tmp/mf_data
} > gpurun_out/examples.log 2>&1
grep -v "^Done" gpurun_out/examples.log | cut -c1-700
