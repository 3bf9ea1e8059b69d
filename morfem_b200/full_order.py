"""Full-order sparse solves: the snapshots of the greedy search and the full-order yardstick sweep.

Reference: ``solve_fem_point`` (implementation.py:468-480) builds ``t0*A0 + t1*A1 + t2*A2``, symmetrises it (:526-528, two sparse
additions and a transpose) and calls ``splu(a).solve(b)`` -- COLAMD ordering, symbolic and numeric factorisation from scratch at every
point, and ``solve_finite_element_method`` (:189-194) walks the points one after the other on one core.

The north star keeps these factorisations on the host (SuperLU); what this module removes is the work that is the SAME at every point:

* sparsity pattern: the union pattern of the three symmetrised operators is built once, in CSC, with the three value arrays aligned on
  it -- assembling ``A(t)`` is then one fused multiply-add over ``nnz`` values, no sparse addition, no transpose;
* column ordering (the symbolic half SuperLU lets a caller reuse): COLAMD runs on the first point (on the first wave of a thread pool); the pattern is then stored
  column-permuted and every later factorisation runs with ``permc_spec="NATURAL"``, which reproduces the ordering the reference's own
  call would have chosen (the pattern does not depend on ``t``);
* right-hand sides: all ports are solved in one ``solve`` call per point (the reference does the same, :475), and the densified port
  matrix is kept;
* points are independent: ``solve_many`` factorises them on a thread pool (scipy's SuperLU wrapper releases the GIL), one point per
  host core, the way the sweep kernels take one point per CTA.

Host code only -- nothing here touches the GPU, and nothing under ``oracle/`` is used; ``tests/test_full_order.py`` compares it with the
oracle's restatement of the reference loop.
"""
from __future__ import annotations

import os
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Optional, Sequence

import numpy as np
from scipy.sparse import csc_matrix, issparse
from scipy.sparse.linalg import splu


def _symmetrised_csc(a) -> csc_matrix:
    """(A + A^T)/2 of one operator, canonical CSC (sorted indices, duplicates summed)."""
    a = csc_matrix(a)
    s = csc_matrix((a + a.T) * 0.5)
    s.sum_duplicates()
    s.sort_indices()
    return s


class FullOrderSolver:
    """``A(t) x = t_b(t) B`` for a sparse model, pattern / ordering / right-hand side prepared once.

    ``solve(t)`` is ``implementation.solve_fem_point`` for a sparse model (implementation.py:468-480); ``solve_many(domain)`` is the sparse
    branch of ``solve_finite_element_method`` (:189-194)."""

    def __init__(self, md, reuse_ordering: bool = True):
        self.md = md
        n = next(a.shape[0] for a in (md.a0, md.a1, md.a2) if a is not None)
        ops = [_symmetrised_csc(csc_matrix((n, n)) if a is None else a) for a in (md.a0, md.a1, md.a2)]   # None = zero operator
        # union pattern: explicit ones on every stored position (an operator value that happens to be 0.0 keeps its slot)
        pat = None
        for s in ops:
            p = csc_matrix((np.ones(s.nnz), s.indices, s.indptr), shape=s.shape)
            pat = p if pat is None else pat + p
        pat = csc_matrix(pat)
        pat.sum_duplicates()
        pat.sort_indices()
        self.n = n
        self.indptr = pat.indptr.astype(np.int32)
        self.indices = pat.indices.astype(np.int32)
        self.values = [self._aligned(s) for s in ops]      # three arrays of length nnz(union), zeros where an operator has no entry
        self.dtype = np.result_type(*[v.dtype for v in self.values], np.float64)
        b = md.b
        self.b_dense = np.ascontiguousarray(np.asarray(b.todense()) if issparse(b) else np.asarray(b))
        self.reuse_ordering = reuse_ordering
        self._perm_c: Optional[np.ndarray] = None           # SuperLU's column permutation of the first factorisation
        self._p_indptr = self._p_indices = None
        self._p_values = None
        self._lock = threading.Lock()

    def _aligned(self, s: csc_matrix) -> np.ndarray:
        """Values of ``s`` scattered onto the union pattern (both have sorted row indices per column)."""
        out = np.zeros(self.indices.size, dtype=s.dtype)
        cols_u = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(self.indptr))
        cols_s = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(s.indptr))
        key_u = cols_u * self.n + self.indices
        key_s = cols_s * self.n + s.indices
        pos = np.searchsorted(key_u, key_s)                  # key_u is strictly increasing (canonical CSC)
        if pos.size and (pos.max() >= key_u.size or not np.array_equal(key_u[pos], key_s)):
            raise AssertionError("operator entry outside the union pattern")
        out[pos] = s.data
        return out

    # ------------------------------------------------------------------------------------------ assembly
    def coefficients(self, t: float):
        md = self.md
        return md.t_a0(t), md.t_a1(t), md.t_a2(t)

    def _combine(self, vals, t: float) -> np.ndarray:
        c0, c1, c2 = self.coefficients(t)
        data = np.multiply(vals[0], c0, dtype=self.dtype)
        data += c1 * vals[1]
        data += c2 * vals[2]
        return data

    def matrix(self, t: float) -> csc_matrix:
        """The symmetrised system matrix of implementation.py:526-528 on the union pattern."""
        return csc_matrix((self._combine(self.values, t), self.indices, self.indptr), shape=(self.n, self.n))

    def rhs(self, t: float) -> np.ndarray:
        """implementation.py:531-533 (densified port matrix times its coefficient)."""
        return self.md.t_b(t) * self.b_dense

    # ------------------------------------------------------------------------------------------ solves
    def _fix_ordering(self, lu):
        """Store the pattern with its columns in SuperLU's elimination order: column i of A sits at position perm_c[i]."""
        perm_c = np.asarray(lu.perm_c)
        order = np.argsort(perm_c)                           # permuted column j is original column order[j]
        counts = np.diff(self.indptr)[order]
        p_indptr = np.concatenate(([0], np.cumsum(counts))).astype(np.int32)
        gather = np.repeat(self.indptr[:-1][order].astype(np.int64) - p_indptr[:-1], counts) + np.arange(int(p_indptr[-1]), dtype=np.int64)
        self._p_indptr, self._p_indices = p_indptr, self.indices[gather]
        self._p_values = [v[gather] for v in self.values]
        self._perm_c = perm_c

    def solve(self, t: float) -> np.ndarray:
        """``splu(system_matrix(t)).solve(impulse_vector(t))`` -> (N, M)."""
        if self._perm_c is None or not self.reuse_ordering:
            lu = splu(self.matrix(t))                        # COLAMD, as the reference's call
            if self.reuse_ordering:
                with self._lock:                             # the first wave of a thread pool gets here concurrently
                    if self._perm_c is None:
                        self._fix_ordering(lu)
            return lu.solve(self.rhs(t))
        ap = csc_matrix((self._combine(self._p_values, t), self._p_indices, self._p_indptr), shape=(self.n, self.n))
        y = splu(ap, permc_spec="NATURAL").solve(self.rhs(t))   # A Pc y = b
        return y[self._perm_c]                                   # x = Pc y

    def solve_many(self, domain: Sequence[float], threads: Optional[int] = None) -> np.ndarray:
        """Every point of ``domain`` -> (P, N, M); points are factorised concurrently on ``threads`` host threads
        (default: the cores this process may run on, at most one per point)."""
        domain = np.asarray(domain, dtype=np.float64).ravel()
        out = np.zeros((domain.size, self.n, self.b_dense.shape[1]), dtype=np.result_type(self.dtype, self.b_dense.dtype))
        if domain.size == 0:
            return out
        if threads is None:
            threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        threads = max(1, min(int(threads), domain.size))
        if threads == 1:
            for i in range(domain.size):
                out[i] = self.solve(domain[i])
            return out

        def work(i):
            out[i] = self.solve(domain[i])

        with ThreadPoolExecutor(max_workers=threads) as pool:
            list(pool.map(work, range(domain.size)))
        return out


_solvers = {}      # one prepared solver per model, a few entries at most


def _fingerprint(a):
    """Cheap content check of one operator (the cache is keyed by object identity; an operator edited in place must miss)."""
    if a is None:
        return None
    data = a.data if issparse(a) else np.asarray(a)
    return (a.shape, int(getattr(a, "nnz", data.size)), complex(data.sum()) if data.size else 0j)


def solver_for(md) -> FullOrderSolver:
    """The prepared solver of ``md``: the greedy search asks for one snapshot per iteration on the same model."""
    key = (id(md.a0), id(md.a1), id(md.a2), id(md.b))
    marks = tuple(_fingerprint(a) for a in (md.a0, md.a1, md.a2, md.b))
    hit = _solvers.get(key)
    if hit is not None and hit[1] == marks and all(x is y for x, y in zip(hit[2], (md.a0, md.a1, md.a2, md.b))):
        hit[0].md = md                                       # the coefficient functions may differ between ModelDefinitions
        return hit[0]
    if len(_solvers) >= 4:
        _solvers.pop(next(iter(_solvers)))
    s = FullOrderSolver(md)
    _solvers[key] = (s, marks, (md.a0, md.a1, md.a2, md.b))
    return s
