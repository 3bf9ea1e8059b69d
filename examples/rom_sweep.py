#!/usr/bin/env python
"""Config 1 of BASELINE.json: the reference's example driver (main.py:18-68) on this package -- full-order S-parameter
sweep (SuperLU on the host, as the north star prescribes), reduced-order sweep on the GPU (greedy basis, projection, batched
solves, S-parameters), and the per-frequency difference the reference prints (`main.py:42-44`, `:67-68`).

    python examples/rom_sweep.py [--data DIR] [--points 100] [--replicate K]

``--data DIR`` loads ``Ct``, ``Tt``, ``WP`` like the reference (dense ``.npy`` arrays, ``main.py:21-23``) or from the sparse
``.npz`` / header-less ``.csv`` forms (``morfem_b200.data_io``).  The reference ships only ``WP.npy`` (``Ct``/``Tt`` are
missing large blobs), so without ``--data`` a synthetic N=3411 waveguide surrogate of the same size is paired with a port
matrix of the shipped ``WP.npy``'s structure.  ``--replicate K`` runs K uncoupled copies of the model on the block diagonal
(the scaling fixture of ``fake_interpolate_bigger_sample.py``).  Plots are replaced by a JSON summary.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
from numpy.linalg import norm

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morfem_b200 import data_io, synthetic                                                              # noqa: E402
from morfem_b200.test_helpers import finite_element_method_gsm, finite_element_method_model_order_reduction_gsm   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default=None, help="directory holding Ct.npy, Tt.npy, WP.npy (reference layout)")
    ap.add_argument("--points", type=int, default=100)
    ap.add_argument("--replicate", type=int, default=1, help="K uncoupled copies of the model (block-diagonal operators, stacked ports)")
    args = ap.parse_args()

    frequency_points = np.linspace(3e9, 5e9, args.points)           # main.py:18
    gate_count = 2                                                    # main.py:19
    if args.data:
        ct, tt, wp = data_io.load_operators(args.data)                  # main.py:21-23 (dense .npy), or sparse .npz / .csv
        source = args.data
    else:
        ct, tt = synthetic.waveguide_operators(9, 1, 379)               # 3411 DOFs, like the shipped WP.npy
        wp = synthetic.shipped_port_matrix()
        source = "synthetic N=3411 surrogate + port matrix with the structure of the shipped data/WP.npy"
    if args.replicate > 1:
        ct, tt, wp = data_io.replicate_block_diagonal(ct, tt, wp, args.replicate)
        source += f", replicated x{args.replicate} on the block diagonal"
    in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, wp)          # main.py:25-26

    t_init = time.time()
    from morfem_b200 import device as dv                             # CUDA context + library load, kept out of both timings
    import torch
    dv.require_cuda()
    torch.cuda.synchronize()
    t0 = time.time()
    gsm_ref = finite_element_method_gsm(frequency_points, gate_count, in_c, in_gamma, in_b)                           # main.py:28
    t1 = time.time()
    gsm_rom, _, qd = finite_element_method_model_order_reduction_gsm(frequency_points, gate_count, in_c, in_gamma, in_b,
                                                                     return_details=True)                             # main.py:40
    t2 = time.time()
    finite_element_method_model_order_reduction_gsm(frequency_points, gate_count, in_c, in_gamma, in_b)               # warm second call
    t3 = time.time()
    error = np.array([norm(gsm_rom[i] - gsm_ref[i]) for i in range(frequency_points.size)])                           # main.py:42-44
    print(json.dumps({"source": source, "N": int(in_c.shape[0]), "points": int(frequency_points.size),
                      "device_init_s": t0 - t_init, "full_order_s": t1 - t0, "reduced_order_s": t2 - t1,
                      "reduced_order_second_call_s": t3 - t2, "basis_size": int(qd.shape[1]),
                      "host_threads": len(os.sched_getaffinity(0)),
                      "error_mean": float(error.mean()), "error_max": float(error.max()),                              # main.py:67-68
                      "S11_dB_first_last": [float(20 * np.log10(abs(gsm_rom[0, 0, 0]))), float(20 * np.log10(abs(gsm_rom[-1, 0, 0])))],
                      "unitarity_max_dev": float(np.abs(np.einsum("fij,fkj->fik", gsm_rom, gsm_rom.conj()) - np.eye(gate_count)).max())}))
    print("Done")


if __name__ == "__main__":
    main()
