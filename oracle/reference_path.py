"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not shipped, not on the product path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this module.  ``morfem_b200`` never does: the product path fails loudly when the CUDA library
is missing instead of falling back to anything in here.

What this is: a numpy/scipy restatement of the four stages of morfem's reduced-order frequency sweep,
each function citing the reference ``file:line`` it follows (paths relative to ``/root/reference``).
The reference is pure Python and its arithmetic lives in third-party libraries that are not vendored and
not pinned by the reference (no requirements file): numpy (LAPACK ``gesdd`` via ``np.linalg.svd``, ``gesv``
via ``np.linalg.inv``), scipy (``_sparsetools.csr_matvecs`` behind ``dense @ csc_array``; LAPACK
``getrf``/``getrs`` via ``scipy.linalg.lu_factor``/``lu_solve``).  Versions in this image: numpy 2.3.5, scipy
1.18.1, OpenBLAS 0.3.30.  The restatement calls the same library entry points at the same call sites.

Parity pinning: the reference has no tests, assertions or golden vectors of its own (SURVEY.md section 4,
8c), so this oracle is pinned against OUTPUTS OF THE LIVE REFERENCE generated in the build container by
``tests/golden/make_golden.py`` (which imports ``/root/reference`` unmodified) and committed as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them.

Complex-valued operators have no reference semantics (the reference silently drops imaginary parts,
``implementation.py:190``).  ``complex_ok=True`` selects the documented restatement: result
dtype complex, everything else unchanged.  Those code paths are labelled "restatement, not reference".
"""
from __future__ import annotations

import math
from typing import Callable, Sequence

import numpy as np
from scipy.linalg import lu_factor, lu_solve

from scipy.constants import pi as PI, epsilon_0 as EPSILON_0, c as C_LIGHT  # test_helpers.py:4 (installed scipy's CODATA set)

KTE = 54.5976295582387           # test_helpers.py:71


# ----------------------------------------------------------------------------------------------- stage 1
def orthonormal_basis(snapshots: np.ndarray) -> np.ndarray:
    """Left singular vectors of the snapshot block, no truncation.

    implementation.py:226 (initial block), :298 (every greedy step), :210 (equally distributed mode):
    ``q = np.linalg.svd(q, full_matrices=False)[0]``.
    """
    return np.linalg.svd(snapshots, full_matrices=False)[0]


def singular_values(snapshots: np.ndarray) -> np.ndarray:
    """Singular values of the same factorisation (used to decide which principal angles are meaningful)."""
    return np.linalg.svd(snapshots, compute_uv=False)


# ----------------------------------------------------------------------------------------------- stage 2
def galerkin_projection(q: np.ndarray, a0, a1, a2, b, conj: bool = False):
    """Reduced operators ``(a0_r, a1_r, a2_r, b_r)``.

    implementation.py:180-184: ``q_t = q.T`` (plain transpose, no conjugate) then
    ``a_i_r = q_t @ a_i @ q`` evaluated left to right, ``b_r = q_t @ b``.
    ``conj=True`` is the north-star ``Q^H`` variant (restatement, not reference); identical for real q.
    """
    q_t = q.conj().T if conj else q.T
    out = []
    for op in (a0, a1, a2):
        out.append(np.asarray((q_t @ op) @ q))
    out.append(np.asarray(q_t @ b))
    return tuple(out)


# ----------------------------------------------------------------------------------------------- stage 3
def system_matrix(c0: float, c1: float, c2: float, a0, a1, a2) -> np.ndarray:
    """implementation.py:526-528: weighted sum of the three operators, then ``(a + a.T) / 2``."""
    a = c0 * a0 + c1 * a1 + c2 * a2
    return (a + a.T) / 2


def impulse_vector(cb: float, b) -> np.ndarray:
    """implementation.py:531-533 (dense branch): ``t_b(t) * b``."""
    return cb * b


def solve_point(c0, c1, c2, cb, a0, a1, a2, b) -> np.ndarray:
    """implementation.py:468-480, dense branch (:476-478): LU with partial pivoting, then solve."""
    a = system_matrix(c0, c1, c2, a0, a1, a2)
    rhs = impulse_vector(cb, b)
    return lu_solve(lu_factor(a), rhs)


def reduced_sweep(domain: Sequence[float], a0, a1, a2, b,
                  t_a0: Callable, t_a1: Callable, t_a2: Callable, t_b: Callable,
                  complex_ok: bool = False) -> np.ndarray:
    """implementation.py:189-194: ``x[i] = solve_fem_point(domain[i], md)``; result shape (F, r, M).

    The reference allocates a float64 result (:190) and so discards imaginary parts; with
    ``complex_ok=True`` the result takes the operands' dtype (restatement, not reference).
    """
    domain = np.asarray(domain)
    dtype = np.result_type(a0.dtype, a1.dtype, a2.dtype, b.dtype) if complex_ok else np.float64
    x = np.zeros((domain.size, b.shape[0], b.shape[1]), dtype=dtype)
    for i in range(domain.size):
        t = domain[i]
        x[i] = solve_point(t_a0(t), t_a1(t), t_a2(t), t_b(t), a0, a1, a2, b)
    return x


def full_order_sweep(domain: Sequence[float], a0, a1, a2, b,
                     t_a0: Callable, t_a1: Callable, t_a2: Callable, t_b: Callable) -> np.ndarray:
    """The full-order yardstick / snapshot solves: implementation.py:189-194 with the SPARSE branch of ``solve_fem_point``
    (:472-475, ``splu(a).solve(b)`` of the symmetrised system matrix :526-528 and the densified right-hand side :531-533).
    Outside the hot path (it stays scipy SuperLU in the product too); the tests use it for the BASELINE config-4 study."""
    from scipy.sparse import csc_matrix
    from scipy.sparse.linalg import splu
    domain = np.asarray(domain)
    x = np.zeros((domain.size, b.shape[0], b.shape[1]))
    for i in range(domain.size):
        t = domain[i]
        a = t_a0(t) * a0 + t_a1(t) * a1 + t_a2(t) * a2
        a = csc_matrix((a + a.T) / 2)
        x[i] = splu(a).solve(np.asarray((t_b(t) * b).todense()))
    return x


# ----------------------------------------------------------------------------------------------- stage 4
def b_coefficient(t: float) -> float:
    """test_helpers.py:70-72; ``math.sqrt`` raises ValueError below the TE cutoff (~2.605 GHz)."""
    return math.sqrt(math.sqrt(((2 * PI * t) / C_LIGHT) ** 2 - KTE ** 2) / t)


def scattering_matrix(frequency_point: float, e: np.ndarray, b: np.ndarray) -> np.ndarray:
    """test_helpers.py:9-14: impedance ``Z = j 2 pi f eps0 e^T b``, ``Y = Z^-1``, ``S = 2 (I + Y)^-1 - I``."""
    gim = 1j * 2 * PI * frequency_point * EPSILON_0 * e.T @ b
    gam = np.linalg.inv(gim)
    ident = np.eye(gam.shape[0])
    return 2 * np.linalg.inv(ident + gam) - ident


def scattering_sweep(frequency_points, x: np.ndarray, b_reduced: np.ndarray,
                     t_b: Callable = b_coefficient) -> np.ndarray:
    """test_helpers.py:60-65: GSM at every point with ``b = t_b(f) * b_reduced``; result (F, M, M) complex."""
    frequency_points = np.asarray(frequency_points)
    m = b_reduced.shape[1]
    gsm = np.zeros((frequency_points.size, m, m), dtype=complex)
    for i in range(frequency_points.size):
        f = frequency_points[i]
        gsm[i] = scattering_matrix(f, x[i], t_b(f) * b_reduced)
    return gsm


# ------------------------------------------------------------------------ greedy residual estimator (row N1)
def error_estimator(q: np.ndarray, domain, a0, a1, a2, b,
                    t_a0: Callable, t_a1: Callable, t_a2: Callable, t_b: Callable) -> np.ndarray:
    """implementation.py:348-452 with ``USE_OPM = False`` (the default, :16): per point the Frobenius norm of the 16-term
    expansion of ``(A(t) q x - t_b b)^H (A(t) q x - t_b b)`` built from the sparse products ``h(a_i) @ a_j`` (:370-385),
    their projections (:387-402) and the reduced solve (:404-415).  Sparse operands as in the reference."""
    def hh(m):
        return m.conj().T                                       # h(), :483-488

    ops = (a0, a1, a2)
    aha = [[hh(ops[i]) @ ops[j] for j in range(3)] for i in range(3)]          # :370-381
    ahb = [hh(ops[i]) @ b for i in range(3)]                                   # :373, :377, :381
    bha = [hh(b) @ ops[i] for i in range(3)]                                   # :382-384
    bh_b = hh(b) @ b                                                           # :385
    qh = hh(q)
    g = [[qh @ aha[i][j] @ q for j in range(3)] for i in range(3)]             # :387-398
    hb = [qh @ ahb[i] for i in range(3)]
    bq = [bha[i] @ q for i in range(3)]                                        # :399-401
    dense = lambda m: m.toarray() if hasattr(m, "toarray") else np.asarray(m)  # noqa: E731
    hb, bq, bh_b = [dense(m) for m in hb], [dense(m) for m in bq], dense(bh_b)
    a0_r, a1_r, a2_r, b_r = galerkin_projection(q, a0, a1, a2, b)              # :404-409
    domain = np.asarray(domain)
    err = np.empty(domain.size)
    for i in range(domain.size):
        t = domain[i]
        c = (t_a0(t), t_a1(t), t_a2(t))
        tb = t_b(t)
        x = solve_point(c[0], c[1], c[2], tb, a0_r, a1_r, a2_r, b_r)            # :415
        x_h = hh(x)
        e = tb * tb * bh_b                                                     # :424-441, same 16 terms
        for ia in range(3):
            for ib in range(3):
                e = e + c[ia] * c[ib] * x_h @ g[ia][ib] @ x
            e = e - c[ia] * tb * x_h @ hb[ia] - tb * c[ia] * bq[ia] @ x
        err[i] = np.linalg.norm(e)
    return err


# ------------------------------------------------------------------------------------------ whole path
def hot_path(snapshots, domain, a0, a1, a2, b,
             t_a0=lambda t: 1.0, t_a1=lambda t: t, t_a2=lambda t: t ** 2, t_b=b_coefficient,
             complex_ok: bool = False):
    """Stages 1-4 chained on a given snapshot block (the greedy point selection of
    implementation.py:217-328 is outside this path): returns ``(q, a0_r, a1_r, a2_r, b_r, x, gsm)``.

    Mirrors implementation.py:178-186 followed by test_helpers.py:60-65.
    """
    q = orthonormal_basis(snapshots)
    a0_r, a1_r, a2_r, b_r = galerkin_projection(q, a0, a1, a2, b)
    x = reduced_sweep(domain, a0_r, a1_r, a2_r, b_r, t_a0, t_a1, t_a2, t_b, complex_ok=complex_ok)
    gsm = scattering_sweep(domain, x, b_r, t_b)
    return q, a0_r, a1_r, a2_r, b_r, x, gsm


# ------------------------------------------------------------------------------------ comparison helpers
def principal_angles(q_ref: np.ndarray, q_new: np.ndarray) -> np.ndarray:
    """Principal angles (radians) between the column spaces of two orthonormal bases."""
    s = np.linalg.svd(q_ref.conj().T @ q_new, compute_uv=False)
    return np.arccos(np.clip(s, -1.0, 1.0))


def subspace_residual(q_ref: np.ndarray, q_new: np.ndarray) -> float:
    """``|| (I - Q_new Q_new^H) Q_ref ||_2``: sine of the largest principal angle, without the arccos
    cancellation that limits ``principal_angles`` to ~1e-8 resolution."""
    resid = q_ref - q_new @ (q_new.conj().T @ q_ref)
    return float(np.linalg.norm(resid, 2))


def rel_err(new: np.ndarray, ref: np.ndarray) -> float:
    """Relative Frobenius error."""
    denom = np.linalg.norm(ref)
    return float(np.linalg.norm(np.asarray(new) - np.asarray(ref)) / (denom if denom > 0 else 1.0))


def align_reduced(a_ref: np.ndarray, q_ref: np.ndarray, q_new: np.ndarray) -> np.ndarray:
    """Express a reference reduced operator in the new basis: ``W^T A_ref W`` with ``W = Q_ref^T Q_new``
    (SURVEY.md section 8c(ii)); valid when both bases span the same space."""
    w = q_ref.T @ q_new
    return w.T @ a_ref @ w
