"""Seeded synthetic inputs for the reduced-order sweep path.

The reference ships only ``data/WP.npy`` (3411x2) and the TE cutoff wavenumber; ``Ct.npy``/``Tt.npy`` are
absent (reference ``.MISSING_LARGE_BLOBS:1-2``).  Everything here is a structured-grid surrogate with the
same algebraic shape as the waveguide problem the reference's driver solves (``main.py:18-26``):

    (Ct + f^2 * Tt') x = beta(f) * WP'      Tt' = -(2*pi/c)^2 Tt,   WP' = sqrt(1/(8e-7 pi^2)) WP

``Ct`` (SPSD stiffness-like) and ``Tt`` (SPD mass-like) are Kronecker sums of 1-D linear-element
stiffness/mass matrices on an nx*ny*nz box, rows ordered with the long (z) axis slowest so the matrices
are banded with half-bandwidth nx*ny + nx + 1.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp

from scipy.constants import c as C_LIGHT, epsilon_0 as EPSILON_0  # same source as test_helpers.py:4

KTE = 54.5976295582387        # test_helpers.py:71, equals shipped data/kTE1.npy
WP_AMPLITUDE = 6.38461011     # peak |value| of the shipped data/WP.npy port profile
GAMMA_SCALE = -((2.0 * math.pi) / C_LIGHT) ** 2          # main.py:25
PORT_SCALE = math.sqrt(1.0 / (8.0 * 1e-7 * math.pi ** 2))  # main.py:26


def _stiffness_1d(n: int, h: float) -> sp.csr_array:
    return sp.diags_array([-np.ones(n - 1), 2.0 * np.ones(n), -np.ones(n - 1)], offsets=[-1, 0, 1], format="csr") / h


def _mass_1d(n: int, h: float) -> sp.csr_array:
    return sp.diags_array([np.ones(n - 1), 4.0 * np.ones(n), np.ones(n - 1)], offsets=[-1, 0, 1], format="csr") * (h / 6.0)


def waveguide_operators(nx: int, ny: int, nz: int, length: float | None = None):
    """Return ``(Ct, Tt)`` as ``csc_array`` (N x N, N = nx*ny*nz), real float64, symmetric.

    The cross-section is ``a x a/2`` with ``a = pi / KTE`` (so the surrogate's dominant cutoff equals the
    reference's hard-coded ``kte``); ``length`` defaults to a cubic cell size along z.
    """
    a = math.pi / KTE
    hx = a / (nx + 1)
    hy = (a / 2.0) / (ny + 1)
    hz = hx if length is None else length / (nz + 1)
    kx, ky, kz = _stiffness_1d(nx, hx), _stiffness_1d(ny, hy), _stiffness_1d(nz, hz)
    mx, my, mz = _mass_1d(nx, hx), _mass_1d(ny, hy), _mass_1d(nz, hz)
    myx = sp.kron(my, mx, format="csr")
    ct = sp.kron(kz, myx, format="csr") + sp.kron(mz, sp.kron(ky, mx, format="csr") + sp.kron(my, kx, format="csr"), format="csr")
    tt = sp.kron(mz, myx, format="csr")
    ct.sum_duplicates()
    tt.sum_duplicates()
    return sp.csc_array(ct), sp.csc_array(tt)


def port_matrix(n: int, ports: int, face: int) -> sp.csc_array:
    """Port matrix ``WP`` (n x ports): column p is a half-sine profile on ``face`` consecutive rows.

    Even ports sit on the last rows-of-first-face block, odd ports on the block before, mirroring the
    shipped ``data/WP.npy`` where column 0 occupies rows 19..37 and column 1 rows 0..18 with values
    ``-6.38461 * sin(pi j / 20)``.  Ports >= 2 are placed on the far end of the row range with higher
    harmonics so that all columns stay linearly independent.
    """
    rows, cols, vals = [], [], []
    for p in range(ports):
        harmonic = 1 + p // 4
        j = np.arange(1, face + 1)
        prof = -WP_AMPLITUDE * np.sin(harmonic * math.pi * j / (face + 1))
        slot = p % 4
        if slot == 0:
            start = face
        elif slot == 1:
            start = 0
        elif slot == 2:
            start = n - face
        else:
            start = n - 2 * face
        rows.append(start + j - 1)
        cols.append(np.full(face, p))
        vals.append(prof)
    return sp.csc_array((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, ports))


def shipped_port_matrix() -> sp.csc_array:
    """Analytic regeneration of the reference's ``data/WP.npy`` (3411 x 2, 19 nnz/column).

    ``tests/golden/make_golden.py`` checks this against the shipped file (max abs diff 1.1e-7: the file
    holds the profile to ~8 digits of its amplitude); the golden fixtures store the shipped values.
    """
    return port_matrix(3411, 2, 19)


def driver_scaled(ct, tt, wp):
    """Apply the reference driver's unit scaling (``main.py:25-26``): returns (in_c, in_gamma, in_b)."""
    return ct, tt * GAMMA_SCALE, wp * PORT_SCALE


def frequency_points(count: int, lo: float = 3e9, hi: float = 5e9) -> np.ndarray:
    """``main.py:18``: ``np.linspace(3e9, 5e9, 100)``."""
    return np.linspace(lo, hi, count)


def snapshot_matrix(n: int, r: int, seed: int = 0, decay_decades: float = 6.0, smooth_sweeps: int = 2,
                    dtype=np.float64) -> np.ndarray:
    """Synthetic snapshot block S (n x r) for basis/projection timing: smoothed Gaussian columns scaled
    by ``10**(-decay_decades * j / r)`` (singular-value decay, cond ~ 10**decay_decades)."""
    rng = np.random.default_rng(seed)
    s = rng.standard_normal((n, r))
    for _ in range(smooth_sweeps):
        s[1:-1] = 0.25 * s[:-2] + 0.5 * s[1:-1] + 0.25 * s[2:]
    s *= 10.0 ** (-decay_decades * np.arange(r) / max(r, 1))
    if np.issubdtype(dtype, np.complexfloating):
        s = s.astype(dtype)
    return np.ascontiguousarray(s)


def reduced_model(r: int, m: int, seed: int = 0, complex_valued: bool = False):
    """Seeded reduced operators ``(a0_r, a1_r, a2_r, b_r)`` with the scaling of a projected waveguide model.

    ``a0_r = X^T D0 X`` and ``a2_r = -(2 pi/c)^2 X^T D2 X`` for a random orthogonal X, with generalized
    eigenvalues ``d0/d2`` spread over k^2 in [2e3, 4e5] rad^2/m^2 so that a 3-5 GHz sweep
    (k^2 in [3.9e3, 1.1e4]) crosses a few resonances, as a projected cavity model does.  ``a1_r`` is zero
    (``test_helpers.py:57`` passes an empty ``csc_array``).  A small non-symmetric perturbation is added
    so that the symmetrisation ``(A + A^T)/2`` of ``implementation.py:528`` is exercised.
    """
    rng = np.random.default_rng(seed)
    x, _ = np.linalg.qr(rng.standard_normal((r, r)))
    d2 = np.exp(rng.uniform(np.log(1e-9), np.log(1e-7), r))
    k2 = np.exp(rng.uniform(np.log(2e3), np.log(4e5), r))
    a0 = x.T @ np.diag(d2 * k2) @ x
    a2 = GAMMA_SCALE * (x.T @ np.diag(d2) @ x)
    skew = rng.standard_normal((r, r))
    a0 = a0 + 1e-3 * np.abs(a0).max() * (skew - skew.T) / r
    b = PORT_SCALE * 1e-2 * rng.standard_normal((r, m))
    a1 = np.zeros((r, r))
    if complex_valued:
        # lossy variant (no reference semantics: SURVEY D2) -- small imaginary symmetric part
        loss = rng.standard_normal((r, r))
        a0 = a0 + 1j * 1e-2 * np.abs(a0).max() * (loss + loss.T) / (2 * r)
        a2 = a2.astype(complex)
        a1 = a1.astype(complex)
        b = b + 1j * 1e-1 * b * rng.standard_normal((r, m))
    return a0, a1, a2, b
