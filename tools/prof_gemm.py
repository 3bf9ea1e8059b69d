"""Time gemm_tn / gemm_nn: python tools/prof_gemm.py N R [f64]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morfem_b200 import device as dv
n, r = int(sys.argv[1]), int(sys.argv[2])
real = len(sys.argv) > 3 and sys.argv[3] == "f64"
dev = dv.require_cuda()
dt = torch.float64 if real else torch.complex128
a = torch.randn(n, r, dtype=dt, device=dev)
b = torch.randn(n, r, dtype=dt, device=dev)
w = torch.randn(r, r, dtype=dt, device=dev)
fl = (2.0 if real else 8.0) * n * r * r
for name, fn in (("gemm_tn(a,a)", lambda: dv.gemm_tn(a, a, conj=True)), ("gemm_tn(a,b)", lambda: dv.gemm_tn(a, b)), ("gemm_nn", lambda: dv.gemm_nn(a, w))):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"N={n} r={r} {'f64' if real else 'c128'} {name}: median {np.median(ts) * 1e3:.1f} us  {fl / np.median(ts) / 1e9:.1f} TFLOP/s")
