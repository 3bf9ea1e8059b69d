"""Small sweeps through the kernels added in round 2, each against the generic kernel (1e-9 on S):
    python tools/crosscheck_sweeps.py
complex128 two-point CTA kernel (MF_LEFT_CFG=8), look-ahead body with the sub-partition remap (r = 128), plain body (r = 200),
float64 instances with the 16-byte FIFO copies (4-warp and 8-warp geometries).  Written for a compute-sanitizer memcheck pass; the
sanitizer is closed on this GPU pool, so it ran plain (gpurun_out/sanitize_plain.log: all six within 4e-12) -- odd sizes, ragged last
tiles and a batch that does not fill the grid are what it exercises."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morfem_b200 import device as dv, synthetic, implementation as impl, test_helpers as th


def run(r, m, nf, variant, real=False, env=None):
    for k, v in (env or {}).items():
        os.environ[k] = v
    a0, a1, a2, b = synthetic.reduced_model(r, m, seed=r)
    f = np.linspace(3e9, 5e9, nf)
    cb = impl.coefficient_array(th.b_coefficient, f)
    dev = dv.require_cuda()
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    conv = up if real else dv.to_device_c128
    args = (dv.symmetrize(conv(a0)), None, dv.symmetrize(conv(a2)), conv(b), up(np.ones_like(f)), up(f), up(f ** 2), up(cb), up(2 * np.pi * f * 8.8541878128e-12))
    res = dv.sweep(*args, want_x=True, want_gsm=True, variant=variant)
    cargs = (dv.symmetrize(dv.to_device_c128(a0)), None, dv.symmetrize(dv.to_device_c128(a2)), dv.to_device_c128(b)) + args[4:]
    ref = dv.sweep(*cargs, want_x=True, want_gsm=True, variant=1)
    torch.cuda.synchronize()
    err = float((res.gsm - ref.gsm).abs().max() / ref.gsm.abs().max())
    for k in (env or {}):
        os.environ.pop(k, None)
    print(f"r={r} m={m} F={nf} variant={variant} real={real} env={env}: max rel diff of S vs the generic kernel {err:.2e}")
    assert err < 1e-9


run(130, 5, 37, 5, env={"MF_LEFT_CFG": "8"})
run(48, 2, 7, 5, env={"MF_LEFT_CFG": "8"})
run(128, 4, 41, 5)
run(200, 3, 19, 5)
run(160, 4, 23, 0, real=True)
run(250, 2, 17, 0, real=True)
print("crosscheck_sweeps: done")
