"""Debug: dump the LU-factored augmented matrix of point 0 from the blocked kernel (built with -DMF_BLOCKED_DEBUG) and compare
with a numpy re-computation of the same blocked algorithm's expected result (LAPACK getrf on the host)."""
import os, sys, ctypes
os.environ["MF_BLOCKED_DUMP"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morfem_b200 import device as dv, synthetic, _ffi
from scipy.linalg import lu_factor
np.set_printoptions(linewidth=250, precision=4, suppress=False)
r, m = (int(v) for v in sys.argv[1:3])
dev = dv.require_cuda()
lib = _ffi.load()
a0, a1, a2, b = synthetic.reduced_model(r, m, seed=5)
f = np.array([3.3e9])
up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
A0 = dv.symmetrize(dv.to_device_c128(a0)); A2 = dv.symmetrize(dv.to_device_c128(a2)); B = dv.to_device_c128(b)
R = (r + 7) // 8 * 8; NCB = R // 8 + (m + 7) // 8; LD = 8 * NCB
ws = torch.zeros(R * LD, dtype=torch.complex128, device=dev)
x = torch.empty((1, r, m), dtype=torch.complex128, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
c = [up(np.ones(1)), up(f), up(f ** 2), up(np.ones(1))]
P = lambda t: ctypes.c_void_p(t.data_ptr())
rc = lib.mf_sweep_lu_gsm_c128(P(A0), None, P(A2), r, P(B), m, r, m, P(c[0]), P(c[1]), P(c[2]), P(c[3]), None, 1, P(x), None, P(info), 3,
                              P(ws), ws.numel() * 16, None)
torch.cuda.synchronize()
print("rc", rc, "info", info.item())
Mg = ws.cpu().numpy().reshape(R, LD)
a = (a0 + a0.T) / 2 + f[0] ** 2 * (a2 + a2.T) / 2
ap = np.eye(R); ap[:r, :r] = a
lu, piv = lu_factor(ap)
U_ref = np.triu(lu)
U_gpu = np.triu(Mg[:, :R].real)
dg = np.diag(U_gpu).copy()
U_gpu[np.diag_indices(R)] = 1.0 / dg
print("U rel err:", np.abs(U_gpu - U_ref).max() / np.abs(U_ref).max())
L_ref = np.tril(lu, -1)
print("lapack piv:", piv.tolist())
if np.abs(U_gpu - U_ref).max() / np.abs(U_ref).max() > 1e-10:
    bad = np.argwhere(np.abs(U_gpu - U_ref) > 1e-8 * np.abs(U_ref).max())
    print("first bad U entries:", bad[:10].tolist())
    print("U_ref[:8,:8]\n", U_ref[:8, :8]); print("U_gpu[:8,:8]\n", U_gpu[:8, :8])
# first-panel L (negated, only rows of the first block are kept by later panels)
print("L11 ref (first 8x8, lapack order)\n", L_ref[:8, :8]); print("-L11 gpu\n", np.tril(Mg[:8, :8].real, -1))
xr = np.linalg.solve(a, b)
print("x err:", np.abs(x.cpu().numpy()[0].real - xr).max() / np.abs(xr).max())
