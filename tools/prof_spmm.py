"""Time the CSR and row-grouped SpMM kernels on a waveguide operator: python tools/prof_spmm.py [r] [nz] [reps] [c128|f64] [nx] [ny]
(cfg2: r=64 nz=1000 nx=20 ny=10; cfg3: r=256 nz=2000 nx=25 ny=20).  MF_SPMM_SPLIT=0 selects the round-1 grouping policy."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morfem_b200 import device as dv, synthetic
r = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
real = len(sys.argv) > 4 and sys.argv[4] == "f64"
nx = int(sys.argv[5]) if len(sys.argv) > 5 else 20
ny = int(sys.argv[6]) if len(sys.argv) > 6 else 10
ct, tt = synthetic.waveguide_operators(nx, ny, nz)
n = ct.shape[0]
dev = dv.require_cuda()
q = torch.randn(n, r, dtype=torch.float64 if real else torch.complex128, device=dev)
w = 8 if real else 16
for grouped in (False, True):
    csr = dv.csr_of_transpose(ct)
    if grouped:
        dv.group_rows(csr, r)
    y = dv.spmm(csr, q); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = dv.spmm(csr, q); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    alg = csr.nnz * 12 + 4 * (n + 1) + 2 * w * n * r
    print(f"{'f64' if real else 'c128'} N={n} r={r} grouped={grouped} split={os.environ.get('MF_SPMM_SPLIT', '1')}: min {min(ts) * 1e3:.1f} us  median {np.median(ts) * 1e3:.1f} us  {alg / np.median(ts) / 1e6:.0f} GB/s (CSR-algorithmic bytes)")

win = dv.build_windows(ct, dev)
if win is not None:
    y2 = dv.spmm_window(win, q); torch.cuda.synchronize()
    print("window kernel max abs diff vs CSR kernels:", float((y2 - y).abs().max()), "wmax", win.wmax)
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y2 = dv.spmm_window(win, q); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"{'f64' if real else 'c128'} N={n} r={r} TMA window kernel: min {min(ts) * 1e3:.1f} us  median {np.median(ts) * 1e3:.1f} us  {alg / np.median(ts) / 1e6:.0f} GB/s (CSR-algorithmic bytes)")
