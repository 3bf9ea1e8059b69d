"""On-disk formats of the driver inputs -- host side only, nothing here touches the GPU.

The reference keeps its operators as DENSE arrays: ``convert_csv_to_json.py:5-8`` turns header-less CSV files into ``data/*.npy``
and ``main.py:21-23`` wraps ``np.load`` of ``Ct.npy``, ``Tt.npy``, ``WP.npy`` in ``csc_array`` (3411 x 3411 float64 = 93 MB per
operator, which is why the shipped repository lacks them).  ``load_operators`` reads that layout and, next to it, a sparse
``.npz`` (``scipy.sparse.save_npz``) and the original ``.csv``; ``save_operators`` writes the sparse form (a few hundred kB for
the same model).  ``replicate_block_diagonal`` is the scaling fixture of ``fake_interpolate_bigger_sample.py:4-10``: ``k``
uncoupled copies of a model on the block diagonal, the port matrix stacked ``k`` times -- without ever forming a dense array.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np
import scipy.sparse as sp

NAMES = ("Ct", "Tt", "WP")          # main.py:21-23
_EXTENSIONS = (".npz", ".npy", ".csv")


def _load_one(directory: str, name: str) -> sp.csc_array:
    for ext in _EXTENSIONS:
        path = os.path.join(directory, name + ext)
        if not os.path.exists(path):
            continue
        if ext == ".npz":
            return sp.csc_array(sp.load_npz(path))
        if ext == ".npy":
            return sp.csc_array(np.load(path))                                    # main.py:21-23
        return sp.csc_array(np.loadtxt(path, delimiter=",", ndmin=2))             # convert_csv_to_json.py:6 (header-less CSV)
    raise FileNotFoundError(f"{name}{{{','.join(_EXTENSIONS)}}} not found in {directory!r}")


def load_operators(directory: str) -> Tuple[sp.csc_array, sp.csc_array, sp.csc_array]:
    """``(Ct, Tt, WP)`` as ``csc_array`` from ``directory`` (sparse ``.npz`` preferred, then the reference's dense ``.npy``, then
    ``.csv``), UNscaled -- apply ``synthetic.driver_scaled`` for the constants of ``main.py:25-26``."""
    ct, tt, wp = (_load_one(directory, n) for n in NAMES)
    if ct.shape[0] != ct.shape[1] or ct.shape != tt.shape or wp.shape[0] != ct.shape[0]:
        raise ValueError(f"inconsistent operator shapes: Ct {ct.shape}, Tt {tt.shape}, WP {wp.shape}")
    return ct, tt, wp


def save_operators(directory: str, ct, tt, wp, dense: bool = False) -> None:
    """Write the three inputs as sparse ``.npz`` (default) or in the reference's dense ``.npy`` layout."""
    os.makedirs(directory, exist_ok=True)
    for name, a in zip(NAMES, (ct, tt, wp)):
        a = sp.csc_array(a)
        if dense:
            np.save(os.path.join(directory, name + ".npy"), a.toarray())
        else:
            sp.save_npz(os.path.join(directory, name + ".npz"), sp.csc_matrix(a))


def replicate_block_diagonal(ct, tt, wp, k: int):
    """``k`` uncoupled copies of the model (fake_interpolate_bigger_sample.py:4-10, :27-30): operators on the block diagonal, the
    port matrix stacked.  The reference script fills the replicated ``Tt`` from ``Ct`` (its line 24 passes ``c`` again); here
    each operator is replicated from itself."""
    if k < 1:
        raise ValueError("k must be >= 1")
    eye = sp.identity(k, format="csc")
    big_ct = sp.csc_array(sp.kron(eye, sp.csc_array(ct), format="csc"))
    big_tt = sp.csc_array(sp.kron(eye, sp.csc_array(tt), format="csc"))
    big_wp = sp.csc_array(sp.vstack([sp.csc_array(wp)] * k, format="csc"))
    return big_ct, big_tt, big_wp
