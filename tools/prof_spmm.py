"""Time the CSR and row-grouped SpMM kernels on the cfg2 operator: python tools/prof_spmm.py [r] [nz] [reps]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morfem_b200 import device as dv, synthetic
r = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ct, tt = synthetic.waveguide_operators(20, 10, nz)
n = ct.shape[0]
dev = dv.require_cuda()
q = torch.randn(n, r, dtype=torch.complex128, device=dev)
for grouped in (False, True):
    csr = dv.csr_of_transpose(ct)
    if grouped:
        dv.group_rows(csr, r)
    y = dv.spmm(csr, q); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = dv.spmm(csr, q); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    alg = csr.nnz * 12 + 4 * (n + 1) + 2 * 16 * n * r
    print(f"N={n} r={r} grouped={grouped}: min {min(ts) * 1e3:.1f} us  median {np.median(ts) * 1e3:.1f} us  {alg / np.median(ts) / 1e6:.0f} GB/s (CSR-algorithmic bytes)")
