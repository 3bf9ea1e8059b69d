"""Run the batched sweep kernel alone (for ncu / timing): python tools/prof_sweep.py R M F VARIANT [REPS] [c128|f64]"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morfem_b200 import device as dv, synthetic, implementation as impl, test_helpers as th

r, m, nf, variant = (int(v) for v in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 15
real = len(sys.argv) > 6 and sys.argv[6] == "f64"
a0, a1, a2, b = synthetic.reduced_model(r, m, seed=11)
f = np.linspace(3e9, 5e9, nf)
cb = impl.coefficient_array(th.b_coefficient, f)
dev = dv.require_cuda()
up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
conv = up if real else dv.to_device_c128
ops = [dv.symmetrize(conv(a0)), None, dv.symmetrize(conv(a2))]
args = (ops[0], ops[1], ops[2], conv(b), up(np.ones_like(f)), up(f), up(f ** 2), up(cb), up(2 * np.pi * f * 8.8541878128e-12))
for _ in range(2):
    res = dv.sweep(*args, want_x=False, want_gsm=True, variant=variant)
torch.cuda.synchronize()
times = []
for _ in range(max(reps, 1)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = dv.sweep(*args, want_x=False, want_gsm=True, variant=variant)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
ms = float(np.median(times))
flops = ((8 / 3) * r ** 3 + 8 * r * r * m + 16 * r * r + 8 * r * m * m) * (0.25 if real else 1.0)
print(f"{'f64' if real else 'c128'} r={r} m={m} F={nf} variant={variant}: min {min(times):.3f} / median {ms:.3f} ms, {nf / ms * 1e3:.3e} pts/s, {flops * nf / ms / 1e9:.2f} TFLOP/s, info!=0: {int((res.info != 0).sum())}")
