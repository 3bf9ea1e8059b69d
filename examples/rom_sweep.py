#!/usr/bin/env python
"""Config 1 of BASELINE.json: the reference's example driver (main.py:18-68) on this package -- full-order S-parameter
sweep (SuperLU on the host, as the north star prescribes), reduced-order sweep on the GPU (greedy basis, projection, batched
solves, S-parameters), and the per-frequency difference the reference prints (`main.py:42-44`, `:67-68`).

    python examples/rom_sweep.py [--data DIR] [--points 100]

``--data DIR`` loads ``Ct.npy``, ``Tt.npy``, ``WP.npy`` exactly like the reference (dense ``.npy`` arrays).  The reference
ships only ``WP.npy`` (``Ct``/``Tt`` are missing large blobs), so without ``--data`` a synthetic N=3411 waveguide surrogate of
the same size is paired with a port matrix of the shipped ``WP.npy``'s structure.  Plots are replaced by a JSON summary.
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np
from numpy.linalg import norm
from scipy.constants import pi, c as c_lightspeed
from scipy.sparse import csc_array

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morfem_b200 import synthetic                                                                       # noqa: E402
from morfem_b200.test_helpers import finite_element_method_gsm, finite_element_method_model_order_reduction_gsm   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default=None, help="directory holding Ct.npy, Tt.npy, WP.npy (reference layout)")
    ap.add_argument("--points", type=int, default=100)
    args = ap.parse_args()

    frequency_points = np.linspace(3e9, 5e9, args.points)           # main.py:18
    gate_count = 2                                                    # main.py:19
    if args.data and all(os.path.exists(os.path.join(args.data, f)) for f in ("Ct.npy", "Tt.npy", "WP.npy")):
        in_c = csc_array(np.load(os.path.join(args.data, "Ct.npy")))        # main.py:21-23
        in_gamma = csc_array(np.load(os.path.join(args.data, "Tt.npy")))
        in_b = csc_array(np.load(os.path.join(args.data, "WP.npy")))
        in_gamma = in_gamma * (-((2 * pi) / c_lightspeed) ** 2)           # main.py:25
        in_b = in_b * math.sqrt(1 / (8 * 1e-7 * pi ** 2))                 # main.py:26
        source = args.data
    else:
        ct, tt = synthetic.waveguide_operators(9, 1, 379)               # 3411 DOFs, like the shipped WP.npy
        in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, synthetic.shipped_port_matrix())
        source = "synthetic N=3411 surrogate + port matrix with the structure of the shipped data/WP.npy"

    t0 = time.time()
    gsm_ref = finite_element_method_gsm(frequency_points, gate_count, in_c, in_gamma, in_b)                           # main.py:28
    t1 = time.time()
    gsm_rom = finite_element_method_model_order_reduction_gsm(frequency_points, gate_count, in_c, in_gamma, in_b)     # main.py:40
    t2 = time.time()
    error = np.array([norm(gsm_rom[i] - gsm_ref[i]) for i in range(frequency_points.size)])                           # main.py:42-44
    print(json.dumps({"source": source, "N": int(in_c.shape[0]), "points": int(frequency_points.size),
                      "full_order_s": t1 - t0, "reduced_order_s": t2 - t1,
                      "error_mean": float(error.mean()), "error_max": float(error.max()),                              # main.py:67-68
                      "S11_dB_first_last": [float(20 * np.log10(abs(gsm_rom[0, 0, 0]))), float(20 * np.log10(abs(gsm_rom[-1, 0, 0])))],
                      "unitarity_max_dev": float(np.abs(np.einsum("fij,fkj->fik", gsm_rom, gsm_rom.conj()) - np.eye(gate_count)).max())}))
    print("Done")


if __name__ == "__main__":
    main()
