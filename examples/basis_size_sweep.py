#!/usr/bin/env python
"""Config 4 of BASELINE.json: the basis-size study of the reference's ``speed_and_error_of_no_points_in_q.py`` (time and mean
S-parameter error against the full-order sweep as a function of the number of snapshot points), re-expressed through the
current API as SURVEY.md D6 prescribes -- that script itself is stale (7-8 argument helpers, ``[:, :, i]`` indexing,
``kTE2.npy``): snapshots at equally spaced points of the frequency axis (``test_helpers.equally_distributed_points``,
implementation.py:197-214), basis size r = ports x points.

    python examples/basis_size_sweep.py [--grid 9 1 379] [--points 101] [--first 3] [--last 29]

Full-order solves (the yardstick sweep and the snapshots) are scipy SuperLU on the host -- outside the hot path by the north
star; each basis size then runs stages 1-4 on the GPU (``model_order_reduction_gsm_from_snapshots``).  One JSON line per size:
``{"snapshot_points", "r", "rom_s", "error_mean", "error_max"}`` (``speed_and_error_of_no_points_in_q.py:31-36``).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
from numpy.linalg import norm
from scipy.sparse import csc_array

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morfem_b200 import synthetic, implementation as impl                                                          # noqa: E402
from morfem_b200.test_helpers import (b_coefficient, equally_distributed_points, finite_element_method_gsm,        # noqa: E402
                                      model_order_reduction_gsm_from_snapshots)


def snapshot_block(points, in_c, in_gamma, in_b):
    """Full-order solutions at ``points`` side by side (N x (M * len(points))) -- implementation.py:204-208."""
    md = impl.ModelDefinition(np.asarray(points), in_c, csc_array(in_c.shape), in_gamma, in_b,
                              lambda t: 1., lambda t: t, lambda t: t ** 2, lambda t: b_coefficient(t))
    x = impl.solve_finite_element_method(md)                     # (P, N, M), SuperLU per point (implementation.py:474-475)
    return np.ascontiguousarray(np.concatenate(list(x), axis=1))


def basis_size_study(frequency_points, in_c, in_gamma, in_b, sizes, ref_gsm=None):
    """Returns a list of dicts, one per number of snapshot points in ``sizes``."""
    gate_count = in_b.shape[1]
    if ref_gsm is None:
        ref_gsm = finite_element_method_gsm(frequency_points, gate_count, in_c, in_gamma, in_b)      # :22
    # every size uses a subset-independent set of points (linspace indices), so solve each distinct point once
    cache = {}
    out = []
    for n_pts in sizes:
        pts = equally_distributed_points(frequency_points, n_pts)                                     # :27
        for t in pts:
            if float(t) not in cache:
                cache[float(t)] = snapshot_block([t], in_c, in_gamma, in_b)
        snaps = np.ascontiguousarray(np.concatenate([cache[float(t)] for t in pts], axis=1))
        start = time.time()
        gsm = model_order_reduction_gsm_from_snapshots(frequency_points, snaps, in_c, in_gamma, in_b)
        rom_s = time.time() - start
        error = np.array([norm(gsm[i] - ref_gsm[i]) for i in range(frequency_points.size)])          # :30-33
        out.append({"snapshot_points": int(n_pts), "r": int(snaps.shape[1]), "rom_s": rom_s,
                    "error_mean": float(error.mean()), "error_max": float(error.max())})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, nargs=3, default=[9, 1, 379], help="nx ny nz of the synthetic waveguide (default: 3411 DOFs)")
    ap.add_argument("--points", type=int, default=101)
    ap.add_argument("--first", type=int, default=3)
    ap.add_argument("--last", type=int, default=29)
    args = ap.parse_args()
    frequency_points = np.linspace(3e9, 5e9, args.points)                                            # :10
    ct, tt = synthetic.waveguide_operators(*args.grid)
    n = ct.shape[0]
    wp = synthetic.shipped_port_matrix() if n == 3411 else synthetic.port_matrix(n, 2, 19)
    in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, wp)
    t0 = time.time()
    ref = finite_element_method_gsm(frequency_points, 2, in_c, in_gamma, in_b)
    print(json.dumps({"N": int(n), "points": int(frequency_points.size), "full_order_s": time.time() - t0}))
    for row in basis_size_study(frequency_points, in_c, in_gamma, in_b, range(args.first, args.last + 1), ref):
        print(json.dumps(row))
    print("Done")


if __name__ == "__main__":
    main()
