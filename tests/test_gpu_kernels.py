"""GPU parity tests of the individual kernels, through the C ABI (ctypes), against numpy/scipy and the oracle.

Tolerances: the north star asks for 1e-10 relative in complex128 on reduced matrices and S-parameters; dense
contractions are checked much tighter (they are plain sums of products)."""
import numpy as np
import pytest
import scipy.sparse as sp

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def crandn(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


@pytest.fixture(scope="module")
def dv():
    from morfem_b200 import device
    device.require_cuda()
    return device


@pytest.mark.parametrize("n,ra,rb,conj", [(1000, 24, 40, False), (4099, 64, 64, True), (777, 5, 3, True),
                                          (20000, 96, 130, False), (33, 256, 256, True), (1, 8, 8, False)])
def test_gemm_tn(dv, n, ra, rb, conj):
    rng = np.random.default_rng(n + ra)
    a, b = crandn(rng, n, ra), crandn(rng, n, rb)
    out = dv.gemm_tn(dv.to_device_c128(a), dv.to_device_c128(b), conj=conj).cpu().numpy()
    ref = (a.conj().T if conj else a.T) @ b
    assert rel(out, ref) < 1e-13


def test_gemm_tn_deterministic_and_strided(dv):
    rng = np.random.default_rng(5)
    big = dv.to_device_c128(crandn(rng, 3000, 80))
    a, b = big[:, :48], big[:, 16:80]          # strided views (lda = 80)
    o1 = dv.gemm_tn(a, b, conj=True).cpu().numpy()
    o2 = dv.gemm_tn(a, b, conj=True).cpu().numpy()
    assert np.array_equal(o1, o2)              # fixed-order reduction of the split partials
    h = big.cpu().numpy()
    assert rel(o1, h[:, :48].conj().T @ h[:, 16:80]) < 1e-13


@pytest.mark.parametrize("n,ra,rb", [(1000, 24, 40), (129, 64, 64), (5000, 100, 7), (130, 256, 256), (64, 3, 2)])
def test_gemm_nn(dv, n, ra, rb):
    rng = np.random.default_rng(n)
    a, w = crandn(rng, n, ra), crandn(rng, ra, rb)
    out = dv.gemm_nn(dv.to_device_c128(a), dv.to_device_c128(w)).cpu().numpy()
    assert rel(out, a @ w) < 1e-13


@pytest.mark.parametrize("n,r", [(300, 5), (129, 64), (1000, 65), (700, 200), (260, 256), (150, 512)])
def test_triangular_apply_skips_zero_tiles_and_equals_gemm_nn(dv, n, r):
    """X R^-1 with an upper-triangular factor (mf_trmm_nn_*): whatever sits below the diagonal tiles is never read, and
    the result equals the full product with a clean triangular factor -- bit for bit -- in both arithmetic types."""
    rng = np.random.default_rng(n + r)
    a, w = crandn(rng, n, r), np.triu(crandn(rng, r, r))
    dirty = w.copy()
    for j0 in range(0, r, 64):                                   # garbage below the 64-column diagonal tiles
        dirty[j0 + 64:, j0:j0 + 64] = 1e300
    full = dv.gemm_nn(dv.to_device_c128(a), dv.to_device_c128(w)).cpu().numpy()
    tri = dv.gemm_nn(dv.to_device_c128(a), dv.to_device_c128(dirty), w_upper=True).cpu().numpy()
    assert np.array_equal(full, tri)
    assert rel(tri, a @ w) < 1e-13
    up = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    full_r = dv.gemm_nn(up(a.real), up(w.real)).cpu().numpy()
    tri_r = dv.gemm_nn(up(a.real), up(dirty.real), w_upper=True).cpu().numpy()
    assert np.array_equal(full_r, tri_r) and rel(tri_r, a.real @ w.real) < 1e-14


@pytest.mark.parametrize("r", [1, 2, 7, 16, 33, 64, 150])
def test_cholesky_and_triangular_inverse(dv, r):
    from morfem_b200 import _ffi
    lib = _ffi.load()
    rng = np.random.default_rng(r)
    x = crandn(rng, 4 * r + 3, r)
    g = x.conj().T @ x
    gd = dv.to_device_c128(g)
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    _ffi.check(lib.mf_potrf_upper_c128(dv._ptr(gd), gd.stride(0), r, dv._ptr(info), dv._stream()))
    assert int(info.item()) == 0
    rr = gd.cpu().numpy()
    ref = np.linalg.cholesky(g).conj().T
    assert rel(rr, ref) < 1e-11
    rinv = torch.empty_like(gd)
    _ffi.check(lib.mf_trtri_upper_c128(dv._ptr(gd), gd.stride(0), r, dv._ptr(rinv), rinv.stride(0), dv._stream()))
    assert rel(rinv.cpu().numpy() @ rr, np.eye(r)) < 1e-10


def test_cholesky_reports_breakdown(dv):
    from morfem_b200 import _ffi
    lib = _ffi.load()
    g = np.diag([1.0, 2.0, -1.0, 3.0]).astype(complex)
    gd = dv.to_device_c128(g)
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    _ffi.check(lib.mf_potrf_upper_c128(dv._ptr(gd), gd.stride(0), 4, dv._ptr(info), dv._stream()))
    assert int(info.item()) == 3


@pytest.mark.parametrize("r,cmplx", [(1, False), (2, True), (5, True), (16, False), (31, True), (64, True), (65, True), (100, True),
                                     (128, False), (200, True), (256, True), (440, False), (512, True)])
def test_jacobi_svd(dv, r, cmplx):
    from morfem_b200 import _ffi
    lib = _ffi.load()
    rng = np.random.default_rng(100 + r)
    m = crandn(rng, r, r) if cmplx else rng.standard_normal((r, r)).astype(complex)
    m = np.triu(m) * (10.0 ** (-4.0 * np.arange(r) / max(r, 1)))[None, :]      # graded, upper triangular like an R factor
    md = dv.to_device_c128(m)
    u = torch.empty_like(md)
    sigma = torch.empty(r, dtype=torch.float64, device="cuda")
    sweeps = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws = torch.empty(lib.mf_jacobi_svd_ws_bytes(r), dtype=torch.uint8, device="cuda")
    _ffi.check(lib.mf_jacobi_svd_c128(dv._ptr(md), md.stride(0), r, dv._ptr(u), u.stride(0), dv._ptr(sigma), 40, 1e-15,
                                      dv._ptr(sweeps), dv._ptr(ws), ws.numel(), dv._stream()))
    uh, sh = u.cpu().numpy(), sigma.cpu().numpy()
    u_ref, s_ref, _ = np.linalg.svd(m)
    assert np.all(np.diff(sh) <= 0)
    assert np.max(np.abs(sh - s_ref)) < 1e-13 * s_ref[0]       # gesdd itself resolves sigma only to eps * sigma_max
    assert rel(uh.conj().T @ uh, np.eye(r)) < max(1e-13, 1e-15 * r)     # product of ~sweeps * r^2 / 2 plane rotations
    # U^H M must have orthogonal rows with norms sigma
    t = uh.conj().T @ m
    assert rel(t @ t.conj().T, np.diag(sh ** 2)) < 1e-12
    assert 1 <= int(sweeps.item()) <= 40


@pytest.mark.parametrize("r", [5, 64, 100, 256, 300])
@pytest.mark.parametrize("real_q", [False, True])
def test_two_operator_spmm_equals_two_single_passes(dv, r, real_q):
    """mf_spmm_csr2_*: two operators with one sparsity pattern in one pass == the single-operator kernel twice (bit for bit:
    same accumulation order) == scipy; operands with different patterns are refused."""
    from morfem_b200 import synthetic
    ct, tt = synthetic.waveguide_operators(6, 5, 30)
    n = ct.shape[0]
    rng = np.random.default_rng(r)
    q = rng.standard_normal((n, r)) if real_q else crandn(rng, n, r)
    qd = torch.from_numpy(np.ascontiguousarray(q)).cuda()
    a0, a1 = dv.csr_of_transpose(ct), dv.csr_of_transpose(tt)
    assert dv.mark_same_pattern(a0, a1, ct, tt) and dv.same_pattern(a0, a1)
    y0, y1 = dv.spmm2(a0, a1, qd)
    assert np.array_equal(y0.cpu().numpy(), dv.spmm(a0, qd).cpu().numpy())
    assert np.array_equal(y1.cpu().numpy(), dv.spmm(a1, qd).cpu().numpy())
    assert rel(y0.cpu().numpy(), ct.T @ q) < 1e-13 and rel(y1.cpu().numpy(), tt.T @ q) < 1e-13
    other = sp.csc_array(sp.random(n, n, density=0.01, random_state=1, format="csc"))
    b = dv.csr_of_transpose(other)
    assert not dv.mark_same_pattern(a0, b, ct, other)
    with pytest.raises(ValueError):
        dv.spmm2(a0, b, qd)


def test_symmetrize(dv):
    rng = np.random.default_rng(0)
    a = crandn(rng, 37, 37)
    out = dv.symmetrize(dv.to_device_c128(a)).cpu().numpy()
    assert np.array_equal(out, (a + a.T) / 2)


@pytest.mark.parametrize("r", [8, 24, 64, 100, 256, 300])
@pytest.mark.parametrize("real_vals", [True, False])
def test_spmm_matches_scipy(dv, r, real_vals):
    rng = np.random.default_rng(r)
    n = 1500
    a = sp.random(n, n, density=0.01, random_state=rng, format="csc")
    a = a + sp.diags_array(rng.standard_normal(n)).tocsc()
    if not real_vals:
        a = a + 1j * sp.random(n, n, density=0.005, random_state=rng, format="csc")
    a = sp.csc_array(a)
    q = crandn(rng, n, r)
    at = dv.csr_of_transpose(a)
    y = dv.spmm(at, dv.to_device_c128(q)).cpu().numpy()
    ref = (q.T @ a).T          # the reference's q_t @ a, transposed
    assert rel(y, ref) < 1e-13
    # row-sliced operand (what a row-sharded rank holds)
    lo, hi = 400, 1100
    ys = dv.spmm(dv.csr_of_transpose(a, row_range=(lo, hi)), dv.to_device_c128(q)).cpu().numpy()
    assert rel(ys, ref[lo:hi]) < 1e-13


def test_spmm_empty_rows_and_long_rows(dv):
    n, r = 300, 40
    rng = np.random.default_rng(3)
    dense = np.zeros((n, n))
    dense[:, 7] = rng.standard_normal(n)        # a^T has one row with n entries (> 32: several fetch rounds)
    dense[100, :50] = rng.standard_normal(50)
    a = sp.csc_array(dense)
    q = crandn(rng, n, r)
    y = dv.spmm(dv.csr_of_transpose(a), dv.to_device_c128(q)).cpu().numpy()
    assert rel(y, dense.T @ q) < 1e-13


@pytest.mark.parametrize("conj", [False, True])
def test_project_rhs(dv, conj):
    from morfem_b200 import synthetic
    rng = np.random.default_rng(1)
    n, r = 900, 48
    b = synthetic.port_matrix(n, 4, 19)
    q = crandn(rng, n, r)
    out = dv.project_rhs(dv.csc_to_device(b), dv.to_device_c128(q), 0, conj=conj).cpu().numpy()
    ref = (q.conj().T if conj else q.T) @ b.toarray()
    assert rel(out, ref) < 1e-14
    # partial sums over two row shards add up to the full projection
    bd = dv.csc_to_device(b)
    p0 = dv.project_rhs(bd, dv.to_device_c128(q[:500]), 0, conj=conj).cpu().numpy()
    p1 = dv.project_rhs(bd, dv.to_device_c128(q[500:]), 500, conj=conj).cpu().numpy()
    assert rel(p0 + p1, ref) < 1e-14


@pytest.mark.parametrize("r", [5, 32, 64, 100, 130, 300])
def test_spmm_row_grouped_matches_csr_kernel(dv, r):
    """The row-grouped SpMM (column unions of 4 / 2 consecutive rows) against scipy and against the CSR kernel, on a FEM
    operator, on a matrix with empty and very long rows, and on a row count that is not a multiple of the group size."""
    import scipy.sparse as sp
    from morfem_b200 import synthetic
    ct, _ = synthetic.waveguide_operators(5, 4, 37)
    rng = np.random.default_rng(r)
    rnd = sp.random(301, 301, density=0.03, random_state=3, format="csc")
    rnd = sp.csc_array(rnd + sp.csc_array((np.ones(301), (np.zeros(301, dtype=int), np.arange(301))), shape=(301, 301)))   # one full row of a^T
    rnd.sort_indices()
    for a in (ct, rnd):
        n = a.shape[0]
        q = rng.standard_normal((n, r)) + 1j * rng.standard_normal((n, r))
        qd = dv.to_device_c128(q)
        csr = dv.csr_of_transpose(a)
        y_csr = dv.spmm(csr, qd).cpu().numpy()
        dv.group_rows(csr, r)
        assert csr.grouped is not None and csr.grouped[0] == (4 if r <= 128 else 2)
        y_grp = dv.spmm(csr, qd).cpu().numpy()
        ref = (q.T @ a).T                                   # implementation.py:181: q_t @ a
        assert rel(y_grp, ref) < 1e-14
        assert rel(y_grp, y_csr) < 1e-14


# ----------------------------------------------------------------------------------- real float64 twins (row N2)
@pytest.mark.parametrize("n,ra,rb", [(1, 1, 1), (100, 3, 5), (4097, 64, 64), (30000, 33, 64), (9000, 130, 130), (2000, 256, 7)])
def test_real_gemm_twins_match_numpy_and_the_complex_kernels(dv, n, ra, rb):
    rng = np.random.default_rng(n + ra)
    a, b = rng.standard_normal((n, ra)), rng.standard_normal((n, rb))
    w = rng.standard_normal((ra, rb))
    up = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    c = dv.gemm_tn(up(a), up(b)).cpu().numpy()
    assert c.dtype == np.float64 and rel(c, a.T @ b) < 1e-14
    cc = dv.gemm_tn(dv.to_device_c128(a), dv.to_device_c128(b)).cpu().numpy()
    assert rel(c, cc.real) < 1e-14 and np.abs(cc.imag).max() == 0.0
    o = dv.gemm_nn(up(a), up(w)).cpu().numpy()
    assert o.dtype == np.float64 and rel(o, a @ w) < 1e-14
    # strided operands (column blocks of a wider matrix, odd leading dimension: the 8-byte copy path)
    wide = rng.standard_normal((n, ra + rb + 1))
    wd = up(wide)
    cs = dv.gemm_tn(wd[:, :ra], wd[:, ra:ra + rb]).cpu().numpy()
    assert rel(cs, wide[:, :ra].T @ wide[:, ra:ra + rb]) < 1e-14


@pytest.mark.parametrize("r", [7, 64, 100, 200, 256])
def test_real_spmm_and_rhs_projection_twins(dv, r):
    from morfem_b200 import synthetic
    ct, _ = synthetic.waveguide_operators(6, 5, 41)
    n = ct.shape[0]
    rng = np.random.default_rng(r)
    q = rng.standard_normal((n, r))
    qd = torch.from_numpy(q).cuda()
    csr = dv.csr_of_transpose(ct)
    y_csr = dv.spmm(csr, qd).cpu().numpy()
    dv.group_rows(csr, r, real=True)
    assert csr.grouped is not None and csr.grouped[0] == 4       # float64: four rows per group up to r = 256
    y_grp = dv.spmm(csr, qd).cpu().numpy()
    ref = (q.T @ ct).T
    assert y_csr.dtype == np.float64 and rel(y_csr, ref) < 1e-14 and rel(y_grp, ref) < 1e-14
    wp = synthetic.port_matrix(n, 3, 19)
    br = dv.project_rhs(dv.csc_to_device(wp), qd).cpu().numpy()
    assert br.dtype == np.float64 and rel(br, q.T @ wp) < 1e-14


@pytest.mark.parametrize("r", [1, 2, 9, 32, 63, 64])
def test_real_jacobi_svd_twin(dv, r):
    import ctypes
    from morfem_b200 import _ffi
    lib = _ffi.load()
    rng = np.random.default_rng(r)
    m = np.triu(rng.standard_normal((r, r))) * (10.0 ** (-4.0 * np.arange(r) / max(r - 1, 1)))[None, :]
    md = torch.from_numpy(np.ascontiguousarray(m)).cuda()
    u = torch.empty((r, r), dtype=torch.float64, device="cuda")
    sigma = torch.empty(r, dtype=torch.float64, device="cuda")
    sweeps = torch.zeros(1, dtype=torch.int32, device="cuda")
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    assert lib.mf_jacobi_svd_f64_supported(r) == 1
    _ffi.check(lib.mf_jacobi_svd_f64(P(md), r, r, P(u), r, P(sigma), 40, 4 * np.finfo(float).eps, P(sweeps), None), "mf_jacobi_svd_f64")
    torch.cuda.synchronize()
    uh, sh = u.cpu().numpy(), sigma.cpu().numpy()
    s_ref = np.linalg.svd(m, compute_uv=False)
    assert np.all(np.diff(sh) <= 0) and np.max(np.abs(sh - s_ref)) < 1e-13 * s_ref[0]
    assert rel(uh.T @ uh, np.eye(r)) < 1e-13
    t = uh.T @ m                                            # rows orthogonal with norms sigma
    gram = t @ t.T
    assert np.max(np.abs(gram - np.diag(np.diag(gram)))) < 1e-12 * s_ref[0] ** 2
    assert 0 < int(sweeps.item()) < 40


@pytest.mark.parametrize("n,r", [(5000, 64), (3000, 130), (2500, 256), (700, 300)])
def test_gram_matrices_use_the_symmetric_tile_path(dv, n, r):
    """gemm_tn(a, a, conj=True) computes only the upper 64 x 64 tiles and mirrors the rest: result exactly Hermitian."""
    rng = np.random.default_rng(n)
    a = crandn(rng, n, r)
    ad = dv.to_device_c128(a)
    g = dv.gemm_tn(ad, ad, conj=True).cpu().numpy()
    assert rel(g, a.conj().T @ a) < 1e-13
    lower_tiles = (np.arange(r)[:, None] // 64) > (np.arange(r)[None, :] // 64)
    assert np.array_equal(g[lower_tiles], g.conj().T[lower_tiles])
    ar = torch.from_numpy(np.ascontiguousarray(a.real)).cuda()
    gr = dv.gemm_tn(ar, ar).cpu().numpy()
    assert rel(gr, a.real.T @ a.real) < 1e-13 and np.array_equal(gr[lower_tiles], gr.T[lower_tiles])
    # a plain transpose of the same operand (conj=False, complex) is NOT Hermitian and must take the general path
    gt = dv.gemm_tn(ad, ad, conj=False).cpu().numpy()
    assert rel(gt, a.T @ a) < 1e-13


@pytest.mark.parametrize("real", [False, True])
@pytest.mark.parametrize("grid,r", [((5, 4, 30), 6), ((6, 5, 40), 64), ((7, 3, 50), 100), ((4, 4, 33), 256)])
def test_tma_window_spmm_matches_scipy(real, grid, r):
    """mf_spmm_window_*: Q row windows staged in shared memory by bulk asynchronous copies (the north star's TMA variant of the
    stage-2 SpMM) against scipy's own product (implementation.py:181-183), including a row range with window-relative
    column indices (the multi-GPU halo layout) and a ragged last block."""
    from morfem_b200 import device as dv, synthetic
    dev = dv.require_cuda()
    ct, tt = synthetic.waveguide_operators(*grid)
    n = ct.shape[0]
    rng = np.random.default_rng(r)
    q = rng.standard_normal((n, r)) if real else rng.standard_normal((n, r)) + 1j * rng.standard_normal((n, r))
    qd = torch.from_numpy(q).to(dev)
    for a in (ct, tt):
        win = dv.build_windows(a, dev)
        assert win is not None and win.wmax <= 9 * (8 + 2)
        y = dv.spmm_window(win, qd).cpu().numpy()
        ref = a.T @ q
        assert np.linalg.norm(y - ref) <= 1e-13 * np.linalg.norm(ref)
        lo, hi = n // 3 + 1, n - 5                                   # a row shard with a halo window, not a multiple of 8 rows
        from morfem_b200 import dist as mfd
        w0, w1 = mfd.column_window(np.asarray(a.indptr), np.asarray(a.indices), lo, hi, n)
        win = dv.build_windows(a, dev, row_range=(lo, hi), col_offset=w0)
        y = dv.spmm_window(win, qd[w0:w1].contiguous()).cpu().numpy()
        assert np.linalg.norm(y - ref[lo:hi]) <= 1e-13 * np.linalg.norm(ref[lo:hi])
