"""Time the r x r kernels: python tools/prof_small.py R"""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morfem_b200 import device as dv, _ffi
r = int(sys.argv[1])
lib = _ffi.load()
rng = np.random.default_rng(0)
a = rng.standard_normal((r + 50, r))
g = a.T @ a
P = lambda t: ctypes.c_void_p(t.data_ptr())
info = torch.zeros(1, dtype=torch.int32, device="cuda")
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return np.median(ts) * 1e3
gd0 = dv.to_device_c128(g)
def potrf():
    gd = gd0.clone(); dv._potrf_upper(gd, info); return gd
t_clone = timeit(lambda: gd0.clone())
print(f"r={r}: potrf {timeit(potrf) - t_clone:.1f} us", end="  ")
rr = potrf()
print(f"trtri {timeit(lambda: dv._trtri_upper(rr)):.1f} us")

ws = torch.empty(lib.mf_chol_inv_ws_bytes(r), dtype=torch.uint8, device="cuda")
rinv = torch.empty((r, r), dtype=torch.complex128, device="cuda")
def fused():
    gd = gd0.clone()
    _ffi.check(lib.mf_chol_inv_upper_c128(P(gd), gd.stride(0), r, P(rinv), r, P(info), P(ws), ws.numel(), None), "chol_inv")
print(f"r={r}: fused chol+inverse {timeit(fused) - t_clone:.1f} us")
u = torch.empty((r, r), dtype=torch.complex128, device="cuda"); sig = torch.empty(r, dtype=torch.float64, device="cuda")
sw = torch.zeros(1, dtype=torch.int32, device="cuda")
jws = torch.empty(lib.mf_jacobi_svd_ws_bytes(r), dtype=torch.uint8, device="cuda")
rt = potrf()
def jac():
    _ffi.check(lib.mf_jacobi_svd_c128(P(rt), rt.stride(0), r, P(u), r, P(sig), 40, 4 * np.finfo(float).eps, P(sw), P(jws), jws.numel(), None), "jacobi")
print(f"r={r}: jacobi svd {timeit(jac, 5):.1f} us, sweeps {int(sw.item())}")
