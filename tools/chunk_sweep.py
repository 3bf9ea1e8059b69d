"""Left-looking sweep: throughput against the number of points a CTA takes per launch (MF_LEFT_CHUNK) and the body (MF_LEFT_VER),
one process, CUDA-event timing, S compared bit for bit with the single-launch result:
    python tools/chunk_sweep.py"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morfem_b200 import device as dv, synthetic, implementation as impl, test_helpers as th


def setup(r, m, nf, real):
    a0, a1, a2, b = synthetic.reduced_model(r, m, seed=11)
    f = np.linspace(3e9, 5e9, nf)
    cb = impl.coefficient_array(th.b_coefficient, f)
    dev = dv.require_cuda()
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    conv = up if real else dv.to_device_c128
    return (dv.symmetrize(conv(a0)), None, dv.symmetrize(conv(a2)), conv(b), up(np.ones_like(f)), up(f), up(f ** 2), up(cb),
            up(2 * np.pi * f * 8.8541878128e-12))


def run(args, nf, ver, chunk, reps=3, base=None):
    os.environ["MF_LEFT_VER"] = str(ver)
    os.environ["MF_LEFT_CHUNK"] = str(chunk)
    res = dv.sweep(*args, want_x=False, want_gsm=True, variant=5)
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = dv.sweep(*args, want_x=False, want_gsm=True, variant=5)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    same = "" if base is None else (" bit-equal" if torch.equal(res.gsm, base) else f" DIFF {float((res.gsm - base).abs().max()):.1e}")
    print(f"   ver {ver} chunk {chunk:3d}: {nf / min(times) * 1e3:.3e} (best) {nf / float(np.median(times)) * 1e3:.3e} (median) points/s{same}", flush=True)
    return res.gsm


cases = [  # r, m, F, real, [(ver, chunk), ...]; the first entry is the single-launch result everything is compared with
    (256, 4, 29600, False, [(2, 0), (2, 1), (2, 2), (2, 3), (3, 1), (3, 2)]),
    (240, 4, 29600, False, [(2, 0), (2, 2), (3, 1), (3, 2), (3, 3)]),
    (224, 4, 29600, False, [(2, 0), (2, 2), (3, 1), (3, 2)]),
    (208, 4, 29600, False, [(2, 0), (2, 2), (3, 0), (3, 1), (3, 2), (3, 3)]),
    (192, 4, 29600, False, [(3, 0), (3, 1), (3, 2)]),
    (176, 4, 29600, False, [(3, 0), (3, 1), (3, 2)]),
    (320, 4, 5920, False, [(3, 0), (3, 1), (3, 2), (2, 0), (2, 2)]),
    (256, 4, 44400, True, [(2, 0), (2, 1), (2, 2), (2, 3), (3, 1), (3, 2)]),
    (224, 4, 44400, True, [(2, 0), (2, 2), (2, 4), (3, 2)]),
    (192, 4, 44400, True, [(2, 0), (2, 2), (2, 4), (3, 2)]),
    (512, 8, 5920, True, [(2, 0), (2, 2), (3, 0), (3, 2)]),
]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for r, m, nf, real, runs in cases:
    print(f"{'float64' if real else 'complex128'} r={r} m={m} F={nf}", flush=True)
    args = setup(r, m, nf, real)
    base = None
    for ver, chunk in runs:
        g = run(args, nf, ver, chunk, base=base)
        if base is None:
            base = g
print("chunk_sweep: done")
