"""Debug helper: compare sweep variants on small seeded models (GPU)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morfem_b200 import device as dv, synthetic
dev = dv.require_cuda()
up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
cases = [(8, 2), (8, 4), (16, 2), (16, 4), (24, 4), (24, 2), (12, 3), (33, 3), (64, 2), (40, 9)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for r, m in cases:
    a0, a1, a2, b = synthetic.reduced_model(r, m, seed=5)
    f = np.linspace(3e9, 5e9, 4)
    ops = [dv.symmetrize(dv.to_device_c128(a0)), None, dv.symmetrize(dv.to_device_c128(a2))]
    args = (ops[0], ops[1], ops[2], dv.to_device_c128(b), up(np.ones_like(f)), up(f), up(f ** 2), up(np.ones_like(f)), up(2 * np.pi * f * 8.8541878128e-12))
    ref = dv.sweep(*args, variant=1)
    new = dv.sweep(*args, variant=3)
    torch.cuda.synchronize()
    xr, xn = ref.x.cpu().numpy(), new.x.cpu().numpy()
    sr, sn = ref.gsm.cpu().numpy(), new.gsm.cpu().numpy()
    ex = np.abs(xn - xr).max() / np.abs(xr).max()
    es = np.abs(sn - sr).max() / np.abs(sr).max()
    print(f"r={r} m={m}: x err {ex:.2e}  S err {es:.2e}  info {new.info.cpu().numpy().tolist()}")
    if ex > 1e-8 and r <= 24:
        d = np.abs(xn[0] - xr[0]) / np.abs(xr[0]).max()
        print("   bad rows (point 0):", np.nonzero(d.max(axis=1) > 1e-8)[0].tolist(), " bad cols:", np.nonzero(d.max(axis=0) > 1e-8)[0].tolist())
