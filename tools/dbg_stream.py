"""Debug/timing helper for the streamed blocked sweep: parity vs scipy LU on seeded models, then timings."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morfem_b200 import device as dv, synthetic
from scipy.linalg import lu_factor, lu_solve
dev = dv.require_cuda()
up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
cases = [(128, 4, 3), (160, 2, 3), (200, 3, 2), (256, 4, 2), (130, 9, 2), (300, 2, 2), (512, 8, 2)]
for r, m, nf in cases:
    a0, a1, a2, b = synthetic.reduced_model(r, m, seed=5)
    f = np.linspace(3e9, 5e9, nf)
    ops = [dv.symmetrize(dv.to_device_c128(a0)), None, dv.symmetrize(dv.to_device_c128(a2))]
    args = (ops[0], ops[1], ops[2], dv.to_device_c128(b), up(np.ones_like(f)), up(f), up(f ** 2), up(np.ones_like(f)), up(2 * np.pi * f * 8.8541878128e-12))
    new = dv.sweep(*args, variant=3)
    torch.cuda.synchronize()
    xn = new.x.cpu().numpy()
    errs = []
    for i, t in enumerate(f):
        a = (a0 + a0.T) / 2 + t ** 2 * (a2 + a2.T) / 2
        xr = lu_solve(lu_factor(a), b)
        errs.append(np.abs(xn[i] - xr).max() / np.abs(xr).max())
    print(f"r={r} m={m}: x err vs LAPACK {max(errs):.2e}  info {new.info.cpu().numpy().tolist()}  finite {np.isfinite(xn).all()}", flush=True)
for r, m, nf in [(128, 4, 2000), (256, 4, 1000), (512, 8, 296)]:
    a0, a1, a2, b = synthetic.reduced_model(r, m, seed=5)
    f = np.linspace(3e9, 5e9, nf)
    ops = [dv.symmetrize(dv.to_device_c128(a0)), None, dv.symmetrize(dv.to_device_c128(a2))]
    args = (ops[0], ops[1], ops[2], dv.to_device_c128(b), up(np.ones_like(f)), up(f), up(f ** 2), up(np.ones_like(f)), up(2 * np.pi * f * 8.8541878128e-12))
    for variant in (3,):
        dv.sweep(*args, want_x=False, variant=variant); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dv.sweep(*args, want_x=False, variant=variant); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        fl = (8 / 3) * r ** 3 + 8 * r * r * m
        print(f"r={r} m={m} F={nf} variant={variant}: {ms:.2f} ms  {nf / ms * 1e3:.3e} pts/s  {fl * nf / ms / 1e9:.2f} TFLOP/s", flush=True)
