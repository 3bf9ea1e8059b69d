"""Drop-in mirror of the reference's ``implementation.py`` public API, backed by the B200 kernels.

Same names, argument meaning, return shapes and array layouts as the reference (citations are into
``/root/reference/implementation.py``):

    morfem(domain, a0, a1, a2, b, t_a0, t_a1, t_a2, t_b) -> (x, q, a0_r, a1_r, a2_r, b_r)        :99-186
    ModelDefinition                                                                              :19-54
    solve_finite_element_method(md) -> x (F, rows(b), M)                                         :189-194
    solve_fem_point(t, md), system_matrix(t, md), impulse_vector(t, md), h(array)                :468-533
    projection_base(md), projection_base_equally_distributed(md)                                 :197-328
    error_estimator(md, q)                                                                       :348-452
    module flags ERROR_THRESHOLD, USE_EQUALLY_DISTRIBUTED, EQUALLY_DISTRIBUTED_REDUCTION_RATE     :12-16

What runs where: the four hot stages (basis orthonormalisation, Galerkin projection, reduced sweep, and -- via
``test_helpers`` -- S-parameters) run on the GPU through ``libmorfem_b200.so``; the full-order sparse
factorisations that produce snapshots stay on scipy's SuperLU exactly as the north star prescribes
(``implementation.py:475``) and are outside the hot path.  There is no CPU fallback for the hot stages.

Documented deviations from the reference
  * complex inputs: the reference allocates a float64 result and silently drops imaginary parts (:190);
    here complex inputs give complex results.  Real inputs give float64 results like the reference.
  * the symmetrisation ``(A + A.T)/2`` (:528) is applied once per reduced operator instead of once per point
    (linear, so mathematically identical).
  * ``TRUNCATION_TOL`` (new, default 0 = keep every column like the reference) drops basis directions whose
    singular value is below ``tol * sigma_max``.
"""
from __future__ import annotations

import math
import time
import warnings
from typing import Callable, Optional

import numpy as np
from scipy.sparse import csc_array, issparse

from . import full_order

ERROR_THRESHOLD = 1e-6
USE_EQUALLY_DISTRIBUTED = False
EQUALLY_DISTRIBUTED_REDUCTION_RATE = 0.97  # in range <0, 1)
PLOT_GREEDY_ITERATIONS = False             # accepted for compatibility; plotting is not part of this package
USE_OPM = False                            # True: incremental greedy search (only the new columns are orthonormalised, multiplied and projected)
TRUNCATION_TOL = 0.0
VERBOSE = False
FULL_ORDER_THREADS = None                  # host threads of the full-order SuperLU solves (None: every core this process may use)


class ModelDefinition:
    """Plain record with the reference's field names (implementation.py:19-54); carries full and reduced models."""

    def __init__(self, domain, a0, a1, a2, b, t_a0: Callable, t_a1: Callable, t_a2: Callable, t_b: Callable):
        self.domain = domain
        self.a0 = a0
        self.a1 = a1
        self.a2 = a2
        self.b = b
        self.t_a0 = t_a0
        self.t_a1 = t_a1
        self.t_a2 = t_a2
        self.t_b = t_b


# ------------------------------------------------------------------------------------------------ helpers
def coefficient_array(fn: Callable, domain: np.ndarray) -> np.ndarray:
    """Evaluate a ``float -> float`` coefficient callable over the whole domain (host side, once per sweep).

    A vectorised call is tried first and accepted only if it reproduces the scalar evaluation on a few probe points to
    within 2 ulp (numpy and libm may round ``**`` differently in the last bit); otherwise the callable is applied point
    by point (the reference's ``b_coefficient`` uses ``math.sqrt`` and only accepts scalars, test_helpers.py:72; this
    package's mirror also accepts arrays).  A ValueError raised by the vectorised call (a point outside the callable's
    domain) falls through to the scalar loop, which raises it for the offending point like the reference.
    """
    domain = np.asarray(domain, dtype=np.float64)
    n = domain.size
    if n == 0:
        return np.zeros(0)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            vec = np.asarray(fn(domain), dtype=np.float64)
        if vec.shape == ():
            vec = np.full(n, float(vec))
        if vec.shape == (n,):
            probes = sorted({0, n // 2, n - 1})
            eps = np.finfo(np.float64).eps
            if all(abs(float(fn(float(domain[i]))) - vec[i]) <= 2 * eps * abs(vec[i]) for i in probes):
                return np.ascontiguousarray(vec)
    except (TypeError, ValueError):
        pass
    return np.array([float(fn(float(t))) for t in domain], dtype=np.float64)


def _is_zero_operator(a) -> bool:
    if a is None:
        return True
    if issparse(a):
        return a.nnz == 0
    return not np.any(a)


def h(array):
    """hermitian conjugate (implementation.py:483-488)"""
    if array.ndim != 2:
        raise Exception("array has to be two-dimensional")
    return array.conj().T


def system_matrix(t: float, md: ModelDefinition):
    """implementation.py:526-528 (host helper kept for API compatibility; the sweep assembles on the device)."""
    a = md.t_a0(t) * md.a0 + md.t_a1(t) * md.a1 + md.t_a2(t) * md.a2
    return (a + a.T) / 2


def impulse_vector(t: float, md: ModelDefinition):
    """implementation.py:531-533"""
    b = md.t_b(t) * md.b
    return b.todense() if issparse(b) else b


def _real_inputs(*arrs) -> bool:
    return not any(np.iscomplexobj(a.data if issparse(a) else a) for a in arrs if a is not None)


# ------------------------------------------------------------------------------- device model (hot stages)
def _device_tensors(obj):
    """The CUDA tensors held by a device operand record (DeviceCSR / DeviceCSC, including the row-grouped arrays)."""
    import torch
    out = []
    for v in vars(obj).values():
        for t in (v if isinstance(v, tuple) else (v,)):
            if isinstance(t, torch.Tensor) and t.is_cuda:
                out.append(t)
    return out


class _DeviceOperators:
    """Full-order operators uploaded once: CSR views for the SpMMs, CSC port matrix."""

    def __init__(self, md: ModelDefinition, side_stream: bool = False, group_for_r: Optional[int] = None):
        """``side_stream=True`` queues the uploads (and, with ``group_for_r``, the row grouping) on a separate CUDA stream so
        that they overlap with whatever the caller runs on the current stream in the meantime (the Cholesky-QR passes only
        need the snapshot block).  Every operator gets its own ready event: the projection of the first operator runs while
        the next one is still crossing PCIe.  ``group_for_r=None`` keeps the plain CSR operands (building the row-grouped
        form costs about one SpMM, so it only pays when the operators stay on the device for many calls)."""
        from . import device as dv
        import torch
        self.dv = dv
        self.dev = dv.require_cuda()
        self.ops = [md.a0, md.a1, md.a2]
        self.zero = [_is_zero_operator(a) for a in self.ops]
        self.grouped = group_for_r is not None
        self._ready = {}
        stream = dv.upload_stream(self.dev) if side_stream else torch.cuda.current_stream()

        def mark(key):
            if side_stream:
                ev = torch.cuda.Event()
                ev.record(stream)
                self._ready[key] = ev

        with torch.cuda.stream(stream):
            self.b = dv.csc_to_device(md.b, self.dev)        # tiny, first
            mark("b")
            # CSR of a^T == CSC arrays of a: what `q_t @ a` multiplies by (implementation.py:181-183)
            self.at = []
            for i, (a, z) in enumerate(zip(self.ops, self.zero)):
                self.at.append(None if z else dv.csr_of_transpose(a, self.dev, group_for_r=group_for_r))
                if not z:
                    mark(i)
        self._a = [None, None, None]   # CSR of a itself, built lazily for the estimator (a_i @ q)
        self.b_host = csc_array(md.b)
        # operators kept for many calls: compare the sparsity patterns once; two operators with one pattern (Ct and Tt of a
        # FEM model) are then multiplied in one pass (device.spmm2).  One-shot uploads skip the host comparison.
        self.pair = None
        live = [i for i, z in enumerate(self.zero) if not z]
        if self.grouped and len(live) == 2 and all(issparse(self.ops[i]) for i in live):
            i0, i1 = live
            if dv.mark_same_pattern(self.at[i0], self.at[i1], csc_array(self.ops[i0]), csc_array(self.ops[i1])):
                self.pair = (i0, i1)

    def wait_ready(self, key=None):
        """Make the current stream wait for the side-stream upload of one operand (``key`` = operator index or "b") or of
        all of them (no-op once waited for, or without a side stream)."""
        import torch
        keys = list(self._ready) if key is None else ([key] if key in self._ready else [])
        cur = torch.cuda.current_stream()
        for k in keys:
            cur.wait_event(self._ready.pop(k))
            # the operand was allocated on the upload stream and is consumed on this one: tell the caching allocator
            obj = self.b if k == "b" else self.at[k]
            for t in _device_tensors(obj):
                t.record_stream(cur)

    def a_csr(self, i):
        if self._a[i] is None and not self.zero[i]:
            a = self.ops[i]
            self._a[i] = self.dv.csr_of_transpose(csc_array(a.T if issparse(a) else np.asarray(a).T), self.dev)
        return self._a[i]

    def project_block(self, x):
        """``x^T (a_i x)`` for every operator and ``x^T b`` -- the callback of ``device.basis_and_projection``."""
        dv = self.dv
        if self.pair is not None and x.shape[1] <= dv.SPMM2_MAX_R:
            i0, i1 = self.pair
            self.wait_ready(i0)
            self.wait_ready(i1)
            y0, y1 = dv.spmm2(self.at[i0], self.at[i1], x)
            g_list = [None, None, None]
            g_list[i0] = dv.gemm_tn(y0, x, conj=False)
            g_list[i1] = dv.gemm_tn(y1, x, conj=False)
            self.wait_ready("b")
            return g_list, dv.project_rhs(self.b, x, 0, conj=False)
        g_list = []
        for i, (at, z) in enumerate(zip(self.at, self.zero)):
            if z:
                g_list.append(None)
                continue
            self.wait_ready(i)
            if self.grouped:
                dv.group_rows(at, x.shape[1], real=x.dtype == dv.F64)
            g_list.append(dv.gemm_tn(dv.spmm(at, x), x, conj=False))
        self.wait_ready("b")
        return g_list, dv.project_rhs(self.b, x, 0, conj=False)

    def project(self, q):
        """Stage 2 (implementation.py:180-184): returns device (a0_r, a1_r, a2_r, b_r); zero operators give zeros."""
        dv = self.dv
        self.wait_ready()
        r = q.shape[1]
        out = []
        for at, z in zip(self.at, self.zero):
            if z:
                out.append(None)
                continue
            y = dv.spmm(at, q)                    # y = a^T q          == (q_t @ a)^T
            out.append(dv.gemm_tn(y, q, conj=False))  # y^T q        == (q_t @ a) @ q
        b_r = dv.project_rhs(self.b, q, 0, conj=False)
        return out[0], out[1], out[2], b_r


def _sweep_device(domain, ops_r, b_r, t_a0, t_a1, t_a2, t_b, want_x=True, want_gsm=False, variant=0):
    """Stage 3 (+4) on reduced operators already resident on the device."""
    from . import device as dv
    import torch
    from scipy.constants import pi, epsilon_0
    dev = b_r.device
    domain = np.asarray(domain, dtype=np.float64)
    coeffs = [coefficient_array(f, domain) for f in (t_a0, t_a1, t_a2, t_b)]
    zs = 2 * pi * domain * epsilon_0
    c0, c1, c2, cb, zsd = (dv.upload(np.ascontiguousarray(c, dtype=np.float64), dev) for c in (*coeffs, zs))
    sym = [None if o is None else dv.symmetrize(o) for o in ops_r]
    return dv.sweep(sym[0], sym[1], sym[2], b_r, c0, c1, c2, cb, zsd, want_x=want_x, want_gsm=want_gsm, variant=variant)


def _warn_singular(info_host: np.ndarray):
    bad = np.nonzero(info_host)[0]
    if bad.size:
        from scipy.linalg import LinAlgWarning
        warnings.warn(f"Diagonal number {int(info_host[bad[0]])} is exactly zero. Singular matrix. "
                      f"(first of {bad.size} affected domain points: index {int(bad[0])})", LinAlgWarning, stacklevel=3)


# -------------------------------------------------------------------------------------------- public API
def solve_fem_point(t: float, md: ModelDefinition):
    """solves (t0*A0 + t1*A1 + t2*A2)X = B for a specific point t in a domain (implementation.py:468-480)"""
    if issparse(md.a0) or issparse(md.a1) or issparse(md.a2):
        # full-order snapshot: stays scipy/SuperLU (north star); pattern, ordering and dense ports prepared once per model
        return full_order.solver_for(md).solve(t)
    one = ModelDefinition(np.array([t], dtype=np.float64), md.a0, md.a1, md.a2, md.b, md.t_a0, md.t_a1, md.t_a2, md.t_b)
    return solve_finite_element_method(one)[0]


def solve_finite_element_method(md: ModelDefinition):
    """implementation.py:189-194.  Dense (reduced) models run as one batched GPU sweep; sparse (full-order) models
    keep the reference's per-point SuperLU factorisations, with the point-independent work hoisted and the points spread
    over the host cores (``full_order.FullOrderSolver``)."""
    domain = np.asarray(md.domain)
    if issparse(md.a0) or issparse(md.a1) or issparse(md.a2):
        return full_order.solver_for(md).solve_many(domain, threads=FULL_ORDER_THREADS)
    from . import device as dv
    real = _real_inputs(md.a0, md.a1, md.a2, md.b)        # the reference's reduced models are real: float64 sweep kernel
    up = (lambda a: dv.real_or_complex_to_device(np.asarray(a))) if real else dv.to_device_c128
    ops = [None if _is_zero_operator(a) else up(a) for a in (md.a0, md.a1, md.a2)]
    b_r = up(md.b.toarray() if issparse(md.b) else md.b)             # impulse_vector densifies a sparse b (implementation.py:531-533)
    res = _sweep_device(domain, ops, b_r, md.t_a0, md.t_a1, md.t_a2, md.t_b, want_x=True, want_gsm=False)
    x = res.x.cpu().numpy()
    _warn_singular(res.info.cpu().numpy())
    return np.ascontiguousarray(x.real) if real else x


def _block_to_device(arr, md: ModelDefinition):
    """Tall host block -> device in the model's field: float64 for real models (the real twins of the kernels run, like
    the reference's own float64 arithmetic), complex128 otherwise."""
    from . import device as dv
    real = _real_inputs(md.a0, md.a1, md.a2, md.b) and not np.iscomplexobj(arr)
    return dv.real_or_complex_to_device(np.asarray(arr), widen=not real)


def _orthonormal_basis_device(snapshots: np.ndarray, md: Optional[ModelDefinition] = None):
    from . import device as dv
    s = dv.to_device_c128(snapshots) if md is None else _block_to_device(snapshots, md)
    q, info = dv.orthonormalize(s, truncation_tol=TRUNCATION_TOL)
    return q, info


def projection_base_equally_distributed(md: ModelDefinition):
    """implementation.py:197-214: snapshots at equally spaced domain indices, one orthonormalisation."""
    reduction_indices = np.linspace(0, md.domain.size - 1,
                                    math.floor(md.domain.size * (1 - EQUALLY_DISTRIBUTED_REDUCTION_RATE)), dtype=int)
    vector_count = md.b.shape[1]
    if issparse(md.a0) or issparse(md.a1) or issparse(md.a2):
        x = full_order.solver_for(md).solve_many(md.domain[reduction_indices], threads=FULL_ORDER_THREADS)   # (P, N, M)
        q = np.ascontiguousarray(x.transpose(1, 0, 2).reshape(md.b.shape[0], vector_count * reduction_indices.size))
    else:
        q = np.empty((md.b.shape[0], vector_count * reduction_indices.size))
        for i in range(reduction_indices.size):
            q[:, vector_count * i:vector_count * i + vector_count] = solve_fem_point(md.domain[reduction_indices[i]], md)
    qd, _ = _orthonormal_basis_device(q, md)
    return _basis_to_host(qd, md)


def _basis_to_host(qd, md) -> np.ndarray:
    q = qd.cpu().numpy()
    return np.ascontiguousarray(q.real) if _real_inputs(md.a0, md.a1, md.a2, md.b) else q


def error_estimator(md: ModelDefinition, q, opm=None, time_stats=None, _ops: Optional[_DeviceOperators] = None):
    """Residual estimator over the whole domain (implementation.py:348-452), device evaluated.

    The reference forms sparse products ``h(a_i) @ a_j`` and projects them; here the same blocks are Gram
    matrices of the SpMM outputs, ``(a_i q)^H (a_j q)`` -- no sparse-sparse product is needed."""
    from . import device as dv
    import torch
    ops = _ops or _DeviceOperators(md)
    qd = q if isinstance(q, torch.Tensor) else _block_to_device(q, md)
    c128 = lambda t: t if (t is None or t.is_complex()) else t.to(torch.complex128)      # noqa: E731  (r x r blocks for the estimator kernel)
    ys = [None if ops.zero[i] else dv.spmm(ops.a_csr(i), qd) for i in range(3)]
    g = [[None if (ys[a] is None or ys[b] is None) else c128(dv.gemm_tn(ys[a], ys[b], conj=True)) for b in range(3)] for a in range(3)]
    hb = [None if y is None else c128(dv.project_rhs(ops.b, y, 0, conj=True)) for y in ys]
    bb = dv.to_device_c128((h(ops.b_host) @ ops.b_host).toarray())
    a0_r, a1_r, a2_r, b_r = ops.project(qd)
    res = _sweep_device(md.domain, [a0_r, a1_r, a2_r], b_r, md.t_a0, md.t_a1, md.t_a2, md.t_b, want_x=True, want_gsm=False)
    dom = np.asarray(md.domain, dtype=np.float64)
    c = [torch.from_numpy(coefficient_array(f, dom)).to(qd.device) for f in (md.t_a0, md.t_a1, md.t_a2, md.t_b)]
    err = dv.estimator(res.x if res.x.is_complex() else res.x.to(torch.complex128), g, hb, bb, c[0], c[1], c[2], c[3])
    return err.cpu().numpy()


class _GreedyState:
    """Device-resident state of the incremental greedy search (``USE_OPM = True``, implementation.py:230-263, :275-295, :455-465).

    The reference keeps 16 projected matrices and grows them blockwise with ``expand_matrix`` while the new vectors are
    Gram-Schmidt-orthonormalised against the base (``orthonormalize_to_base``).  Here the base ``q``, ``a_i^T q`` (for the Galerkin
    projection) and ``a_i q`` (whose Gram matrices are the estimator blocks ``q^H a_i^H a_j q``) live in column-growing device
    buffers; a greedy step multiplies, projects and orthonormalises ONLY the new M columns: O(N r M) work instead of O(N r^2)."""

    def __init__(self, md: ModelDefinition, ops: "_DeviceOperators", first_block):
        from . import device as dv
        import torch
        self.dv, self.torch, self.md, self.ops = dv, torch, md, ops
        self.live = [i for i in range(3) if not ops.zero[i]]
        qd, _ = _orthonormal_basis_device(first_block, md)                  # implementation.py:222-226
        self.n, self.dtype, self.dev = qd.shape[0], qd.dtype, qd.device
        self.m = md.b.shape[1]
        self.cap, self.r = 0, 0
        self.q = None
        self.yt = {i: None for i in self.live}                               # a_i^T q
        self.y = {i: None for i in self.live}                                # a_i q
        self.a_r = {i: torch.zeros((0, 0), dtype=self.dtype, device=self.dev) for i in self.live}
        self.g = {(a, b): torch.zeros((0, 0), dtype=self.dtype, device=self.dev) for a in self.live for b in self.live}
        self.hb = {i: torch.zeros((0, self.m), dtype=self.dtype, device=self.dev) for i in self.live}
        self.b_r = torch.zeros((0, self.m), dtype=self.dtype, device=self.dev)
        self.bb = dv.to_device_c128((h(ops.b_host) @ ops.b_host).toarray())
        self._append_orthonormal(qd)

    def _reserve(self, cols: int):
        if cols <= self.cap:
            return
        torch = self.torch
        cap = max(cols, 2 * self.cap, 16)

        def grow(old):
            new = torch.empty((self.n, cap), dtype=self.dtype, device=self.dev)
            if old is not None and self.r:
                new[:, :self.r].copy_(old[:, :self.r])
            return new
        self.q = grow(self.q)
        for i in self.live:
            self.yt[i], self.y[i] = grow(self.yt[i]), grow(self.y[i])
        self.cap = cap

    def _append_orthonormal(self, v):
        """``v`` (N x k, orthonormal and orthogonal to the base): SpMMs, projections and Gram blocks of the new columns only."""
        dv, torch = self.dv, self.torch
        r, k = self.r, v.shape[1]
        self._reserve(r + k)
        self.q[:, r:r + k].copy_(v)
        v = self.q[:, r:r + k]
        q_old, q_all = self.q[:, :r], self.q[:, :r + k]
        for i in self.live:
            dv.spmm(self.ops.at[i], v.contiguous(), out=self.yt[i][:, r:r + k])
            dv.spmm(self.ops.a_csr(i), v.contiguous(), out=self.y[i][:, r:r + k])

        def grown(old, right, bottom_left):
            """[[old, right_top], [bottom_left, right_bottom]] -- expand_matrix (implementation.py:455-465)"""
            new = torch.empty((r + k, r + k), dtype=self.dtype, device=self.dev)
            new[:r, :r].copy_(old)
            new[:, r:].copy_(right)
            if r:
                new[r:, :r].copy_(bottom_left)
            return new
        vc = v.contiguous()
        for i in self.live:
            right = dv.gemm_tn(self.yt[i][:, :r + k], vc, conj=False)                          # q_all^T a_i v
            bl = dv.gemm_tn(self.yt[i][:, r:r + k].contiguous(), q_old, conj=False) if r else None   # v^T a_i q_old
            self.a_r[i] = grown(self.a_r[i], right, bl)
        for a in self.live:
            for b in self.live:
                right = dv.gemm_tn(self.y[a][:, :r + k], self.y[b][:, r:r + k].contiguous(), conj=True)
                bl = dv.gemm_tn(self.y[a][:, r:r + k].contiguous(), self.y[b][:, :r], conj=True) if r else None
                self.g[(a, b)] = grown(self.g[(a, b)], right, bl)
            self.hb[a] = torch.cat((self.hb[a], dv.project_rhs(self.ops.b, self.y[a][:, r:r + k].contiguous(), 0, conj=True)), dim=0)
        self.b_r = torch.cat((self.b_r, dv.project_rhs(self.ops.b, vc, 0, conj=False)), dim=0)
        self.r = r + k

    def append(self, vectors):
        """``orthonormalize_to_base`` (implementation.py:491-508) for a block of new full-order solutions: two block Gram-Schmidt
        passes against the base (the second removes what rounding left of the first), then the block itself is orthonormalised
        (CholeskyQR2 + SVD of R); then the state grows by the new columns."""
        dv = self.dv
        v = _block_to_device(vectors, self.md)
        if v.dtype != self.dtype:
            raise _ffi_error("incremental greedy search: a complex solution joined a real basis (set USE_OPM = False for such models)")
        q_old = self.q[:, :self.r]
        for _ in range(2):
            c = dv.gemm_tn(q_old, v, conj=True)                              # coefficients against the base
            v = v - dv.gemm_nn(q_old, c)
        v, _ = dv.orthonormalize(v.contiguous(), truncation_tol=TRUNCATION_TOL)
        self._append_orthonormal(v)

    def error_estimator(self):
        """implementation.py:348-452 with the blocks taken from the state (the reference's ``USE_OPM`` branch, :351-367)."""
        dv, torch, md = self.dv, self.torch, self.md
        ops_r = [self.a_r.get(i) for i in range(3)]
        res = _sweep_device(md.domain, ops_r, self.b_r, md.t_a0, md.t_a1, md.t_a2, md.t_b, want_x=True, want_gsm=False)
        dom = np.asarray(md.domain, dtype=np.float64)
        c = [torch.from_numpy(coefficient_array(f, dom)).to(self.dev) for f in (md.t_a0, md.t_a1, md.t_a2, md.t_b)]
        c128 = lambda t: t if (t is None or t.is_complex()) else t.to(torch.complex128)      # noqa: E731
        g = [[c128(self.g[(a, b)].contiguous()) if (a in self.live and b in self.live) else None for b in range(3)] for a in range(3)]
        hb = [c128(self.hb[a].contiguous()) if a in self.live else None for a in range(3)]
        err = dv.estimator(res.x if res.x.is_complex() else res.x.to(torch.complex128), g, hb, self.bb, c[0], c[1], c[2], c[3])
        return err.cpu().numpy()

    def basis(self):
        return self.q[:, :self.r].contiguous()


def _ffi_error(msg):
    from ._ffi import MorfemB200Error
    return MorfemB200Error(msg)


def new_solution_for_projection_base(md: ModelDefinition, q, opm=None, time_stats=None, _ops=None):
    """implementation.py:321-328"""
    error = error_estimator(md, q, opm, time_stats, _ops=_ops)
    idx_max = error.argmax()
    if error[idx_max] < ERROR_THRESHOLD:
        return None, error
    return solve_fem_point(md.domain[idx_max], md), error


def projection_base(md: ModelDefinition, _return_device: bool = False, _ops: Optional[_DeviceOperators] = None):
    """Greedy basis construction (implementation.py:217-328): start from the two end points of the domain, add
    the full-order solution at the arg-max of the residual estimator until it drops below ERROR_THRESHOLD,
    re-orthonormalising ``[q | q_new]`` each time (:297-298)."""
    from . import device as dv
    ops = _ops or _DeviceOperators(md)            # a caller that projects afterwards passes its own (one upload per call)
    if issparse(md.a0) or issparse(md.a1) or issparse(md.a2):   # the two end points, factorised side by side
        ends = full_order.solver_for(md).solve_many(md.domain[[0, -1]], threads=FULL_ORDER_THREADS)
        initial_vectors = np.hstack((ends[0], ends[1]))
    else:
        initial_vectors = np.hstack((solve_fem_point(md.domain[0], md), solve_fem_point(md.domain[-1], md)))
    if USE_OPM:                                   # incremental search: implementation.py:230-263 (set-up), :275-295 (growth)
        state = _GreedyState(md, ops, initial_vectors)
        projection_base.last_errors = []
        while True:
            error = state.error_estimator()
            projection_base.last_errors.append(error)
            idx_max = error.argmax()
            if error[idx_max] < ERROR_THRESHOLD:
                break
            state.append(solve_fem_point(md.domain[idx_max], md))
        qd = state.basis()
        return qd if _return_device else _basis_to_host(qd, md)
    qd, _ = _orthonormal_basis_device(initial_vectors, md)
    while True:
        q_new, _error = new_solution_for_projection_base(md, qd, _ops=ops)
        if q_new is None:
            break
        import torch
        new = _block_to_device(q_new, md)
        if new.dtype != qd.dtype:                 # a complex full-order solution joins a real basis: continue in complex128
            qd, new = qd.to(torch.complex128), new.to(torch.complex128)
        stacked = torch.cat((qd, new), dim=1).contiguous()
        qd, _ = dv.orthonormalize(stacked, truncation_tol=TRUNCATION_TOL)
    return qd if _return_device else _basis_to_host(qd, md)


# ---- host helpers of the reference kept for API completeness (none of them sits on the hot path) ---------------------------------------
class OfflinePhaseMatrices:
    """implementation.py:57-73: holder of the 16 projected estimator matrices of the ``USE_OPM`` mode.  The device path keeps them in
    ``_GreedyState``; this record only exists so that reference code constructing one keeps working."""
    qh_a0h_a0_q = qh_a0h_a1_q = qh_a0h_a2_q = qh_a0h_b = None
    qh_a1h_a0_q = qh_a1h_a1_q = qh_a1h_a2_q = qh_a1h_b = None
    qh_a2h_a0_q = qh_a2h_a1_q = qh_a2h_a2_q = qh_a2h_b = None
    bh_a0_q = bh_a1_q = bh_a2_q = bh_b = None


class TimeStatistics:
    """implementation.py:76-96: wall-clock buckets (the class-level dict of the reference, shared by all instances, included)."""
    times = {}

    def __init__(self):
        self.clock = time.time()

    def start_clock(self):
        self.clock = time.time()

    def add_time(self, name):
        now = time.time()
        self.times[name] = self.times.get(name, 0.0) + now - self.clock
        self.clock = now

    def add_custom_time(self, name, since):
        self.times[name] = self.times.get(name, 0.0) + time.time() - since

    def print_statistics(self):
        whole = self.times.get("Whole", sum(self.times.values())) or 1.0
        for name, t in self.times.items():
            print(f"{name}: {t:.3f} s ({100.0 * t / whole:.1f} %)")


def orthonormalize_vector_to_base(vector: np.ndarray, base: np.ndarray) -> np.ndarray:
    """implementation.py:511-523: one classical Gram-Schmidt pass of ``vector`` against the orthonormal columns of ``base``."""
    if vector.ndim != 1 or base.ndim != 2:
        raise Exception("vector has to be one-dimensional and base has to be two-dimensional")
    out = vector - base @ (base.T @ vector)          # sum of the projections on the (normalised) base vectors, no conjugate as in :518
    return out / np.linalg.norm(out)


def orthonormalize_to_base(vectors: np.ndarray, base: np.ndarray) -> np.ndarray:
    """implementation.py:491-508: the new vectors one after the other, each against the base extended by its predecessors."""
    if vectors.ndim != 2:
        raise Exception("new_vectors has to be two-dimensional")
    if base is not None and base.ndim != 2:
        raise Exception("base has to be two-dimensional or None")
    cols = []
    for i in range(vectors.shape[1]):
        col = orthonormalize_vector_to_base(vectors[:, i], base)
        base = np.hstack((base, col[:, None]))
        cols.append(col)
    return np.stack(cols, axis=1)


def expand_matrix(original: np.ndarray, old_q: np.ndarray, middle, new_part_q: np.ndarray):
    """implementation.py:455-465: ``[q | q_new]^H middle [q | q_new]`` from its leading block ``original = q^H middle q``."""
    return np.block([[original, h(old_q) @ middle @ new_part_q],
                     [h(new_part_q) @ middle @ old_q, h(new_part_q) @ middle @ new_part_q]])


def residual_norm(md: ModelDefinition, q: np.ndarray) -> np.ndarray:
    """implementation.py:331-345 (never called by the reference itself): true residual ``|| A(t) q x - t_b(t) b ||`` per domain point,
    with the reduced model solved on the device."""
    ops = _DeviceOperators(md)
    qd = _block_to_device(q, md)
    a0_r, a1_r, a2_r, b_r = ops.project(qd)
    res = _sweep_device(md.domain, [a0_r, a1_r, a2_r], b_r, md.t_a0, md.t_a1, md.t_a2, md.t_b, want_x=True, want_gsm=False)
    x = res.x.cpu().numpy()
    out = np.empty(np.asarray(md.domain).size)
    for i, t in enumerate(np.asarray(md.domain)):
        out[i] = np.linalg.norm(system_matrix(t, md) @ (np.asarray(q) @ x[i]) - impulse_vector(t, md))
    return out


def morfem_from_snapshots(snapshots: np.ndarray, domain: np.ndarray, a0: csc_array, a1: csc_array, a2: csc_array, b: csc_array,
                          t_a0: Callable[[float], float] = lambda t: 1.,
                          t_a1: Callable[[float], float] = lambda t: t,
                          t_a2: Callable[[float], float] = lambda t: t ** 2,
                          t_b: Callable[[float], float] = lambda t: t):
    """The four hot stages on a GIVEN snapshot block (N x r): ``q = svd(snapshots)[0]`` (implementation.py:226/298/210),
    then lines :178-186 verbatim.  Same 6-tuple as ``morfem``; this is ``morfem`` minus the greedy point selection and
    its full-order SuperLU solves, i.e. exactly the path the north star puts on the GPU."""
    from . import device as dv
    md = ModelDefinition(domain, a0, a1, a2, b, t_a0, t_a1, t_a2, t_b)
    ops = _DeviceOperators(md)
    all_real = _real_inputs(a0, a1, a2, b) and not np.iscomplexobj(snapshots)      # real data: float64 stage-1/2 kernels
    qd, (a0_r, a1_r, a2_r), b_r, _ = dv.basis_and_projection(dv.real_or_complex_to_device(snapshots, widen=not all_real),
                                                             ops.project_block, truncation_tol=TRUNCATION_TOL)
    res = _sweep_device(domain, [a0_r, a1_r, a2_r], b_r, t_a0, t_a1, t_a2, t_b, want_x=True, want_gsm=False)
    _warn_singular(res.info.cpu().numpy())
    real = _real_inputs(a0, a1, a2, b) and not np.iscomplexobj(snapshots)
    r = qd.shape[1]

    def host(t):
        if t is None:
            return np.zeros((r, r)) if real else np.zeros((r, r), dtype=complex)
        arr = t.cpu().numpy()
        return np.ascontiguousarray(arr.real) if real else arr

    return host(res.x), host(qd), host(a0_r), host(a1_r), host(a2_r), host(b_r)


def morfem(domain: np.ndarray, a0: csc_array, a1: csc_array, a2: csc_array, b: csc_array,
           t_a0: Callable[[float], float] = lambda t: 1.,
           t_a1: Callable[[float], float] = lambda t: t,
           t_a2: Callable[[float], float] = lambda t: t ** 2,
           t_b: Callable[[float], float] = lambda t: t):
    """Solve ``(t_a0 a0 + t_a1 a1 + t_a2 a2) x = t_b b`` over ``domain`` by model-order reduction.

    Same contract as the reference (implementation.py:99-186): returns ``(x, q, a0_r, a1_r, a2_r, b_r)`` with
    ``x`` of shape (I, Nr, M), ``q`` (N, Nr), reduced operators (Nr, Nr) and ``b_r`` (Nr, M), all C-contiguous
    ndarrays (float64 for real inputs)."""
    from . import device as dv
    import torch
    md = ModelDefinition(domain, a0, a1, a2, b, t_a0, t_a1, t_a2, t_b)
    start = time.time()
    ops = _DeviceOperators(md)                                                      # uploaded once: greedy search and projection share it
    if USE_EQUALLY_DISTRIBUTED:
        qd = _block_to_device(projection_base_equally_distributed(md), md)
    else:
        qd = projection_base(md, _return_device=True, _ops=ops)
    if VERBOSE:
        print("Projection base: ", time.time() - start, " s")
    a0_r, a1_r, a2_r, b_r = ops.project(qd)                                         # :178-184
    res = _sweep_device(domain, [a0_r, a1_r, a2_r], b_r, t_a0, t_a1, t_a2, t_b, want_x=True, want_gsm=False)  # :186
    _warn_singular(res.info.cpu().numpy())
    real = _real_inputs(a0, a1, a2, b)
    r = qd.shape[1]

    def host(t):
        if t is None:
            return np.zeros((r, r)) if real else np.zeros((r, r), dtype=complex)
        arr = t.cpu().numpy()
        return np.ascontiguousarray(arr.real) if real else arr

    return host(res.x), host(qd), host(a0_r), host(a1_r), host(a2_r), host(b_r)
