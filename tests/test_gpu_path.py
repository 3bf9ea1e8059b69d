"""GPU parity of stages 1 + 2 and of the chained path, through the Python mirror of the reference API, against the
committed outputs of the live reference and against the CPU oracle.

Basis parity is by subspace distance (principal angles): column signs of singular vectors are arbitrary and
directions with sigma_j/sigma_0 below ~1e-10 are rounding noise in the reference's own SVD."""
import os
import warnings

import numpy as np
import pytest
from scipy.sparse import csc_array

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import reference_path as orc   # noqa: E402  (checker only)
from morfem_b200 import synthetic           # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def dv():
    from morfem_b200 import device
    device.require_cuda()
    return device


def operators_from(g):
    nx, ny, nz = (int(v) for v in g["grid"])
    ct, tt = synthetic.waveguide_operators(nx, ny, nz)
    wp = synthetic.port_matrix(ct.shape[0], int(g["ports"]), int(g["face"]))
    return synthetic.driver_scaled(ct, tt, wp)


# ------------------------------------------------------------------------------------------------ stage 1
@pytest.mark.parametrize("n,r,decades,cmplx", [(5000, 16, 3.0, False), (20000, 64, 6.0, False), (3000, 40, 6.0, True),
                                               (40000, 128, 5.0, False), (700, 1, 0.0, False), (9000, 33, 11.0, False)])
def test_orthonormalize_matches_svd(dv, n, r, decades, cmplx):
    s = synthetic.snapshot_matrix(n, r, seed=r, decay_decades=decades)
    if cmplx:
        s = s + 1j * synthetic.snapshot_matrix(n, r, seed=r + 1, decay_decades=decades)
    qd, info = dv.orthonormalize(dv.to_device_c128(s))
    q = qd.cpu().numpy()
    u, sig, _ = np.linalg.svd(s, full_matrices=False)      # implementation.py:226
    assert np.linalg.norm(q.conj().T @ q - np.eye(r)) < 1e-13 * r
    assert np.max(np.abs(info.sigma - sig) / sig[0]) < 1e-13
    if not cmplx:
        assert np.abs(q.imag).max() == 0.0
    # leading k-dimensional singular subspaces agree to the perturbation bound eps*sigma_0/gap_k (Davis-Kahan / Wedin):
    # no two backward-stable algorithms can agree better than that, LAPACK gesdd included
    eps = np.finfo(float).eps
    for k in sorted({1, max(1, r // 4), max(1, r // 2), r}):
        gap = sig[k - 1] - (sig[k] if k < r else 0.0)
        tol = max(1e-10, 1e3 * eps * sig[0] / gap)
        if tol > 1e-3:
            continue
        resid = orc.subspace_residual(u[:, :k], q[:, :k])
        assert resid < tol, (k, resid, tol)


def test_orthonormalize_live_reference_svd_pair(dv):
    """The last np.linalg.svd call of the reference's greedy loop on the N=3411 driver case (input, output)."""
    g = np.load(os.path.join(GOLDEN, "cfg1_rom3411.npz"))
    s = g["svd_in"]
    qd, info = dv.orthonormalize(dv.to_device_c128(s))
    q = qd.cpu().numpy().real
    keep = int(np.count_nonzero(info.sigma > 1e-10 * info.sigma[0]))
    assert keep >= 4
    gap = info.sigma[keep - 1] - (info.sigma[keep] if keep < s.shape[1] else 0.0)
    tol = max(1e-10, 1e3 * np.finfo(float).eps * info.sigma[0] / gap)
    assert orc.subspace_residual(g["svd_out"][:, :keep], q[:, :keep]) < tol
    assert np.max(orc.principal_angles(g["svd_out"][:, :keep], q[:, :keep])) < max(1e-6, tol)


def test_orthonormalize_truncation(dv):
    rng = np.random.default_rng(0)
    base = np.linalg.qr(rng.standard_normal((4000, 12)))[0]
    s = np.hstack([base, base[:, :4] @ rng.standard_normal((4, 4))])     # rank 12 block with 16 columns
    qd, info = dv.orthonormalize(dv.to_device_c128(s), truncation_tol=1e-10)
    assert info.kept == 12 and qd.shape == (4000, 12)
    q = qd.cpu().numpy().real
    assert orc.subspace_residual(base, q) < 1e-10
    qd_all, info_all = dv.orthonormalize(dv.to_device_c128(s))           # default: keep every column like the reference
    assert qd_all.shape == (4000, 16) and info_all.kept == 16


# ------------------------------------------------------------------------------------------------ stage 2
@pytest.mark.parametrize("fixture", ["stages_n600", "equidist_n600"])
def test_projection_matches_live_reference(dv, fixture):
    from morfem_b200 import implementation as impl, test_helpers as th
    g = np.load(os.path.join(GOLDEN, fixture + ".npz"))
    in_c, in_gamma, in_b = operators_from(g)
    md = impl.ModelDefinition(g["f"], in_c, csc_array(in_c.shape), in_gamma, in_b, lambda t: 1.0, lambda t: t, lambda t: t ** 2,
                              th.b_coefficient)
    ops = impl._DeviceOperators(md)
    a0_r, a1_r, a2_r, b_r = ops.project(dv.to_device_c128(g["q"]))       # stage isolated: the reference's own q
    assert a1_r is None                                                  # empty operator -> exact zeros (D3)
    assert not np.any(g["a1_r"])
    for new, key in ((a0_r, "a0_r"), (a2_r, "a2_r"), (b_r, "b_r")):
        out = new.cpu().numpy()
        assert np.abs(out.imag).max() == 0.0
        assert orc.rel_err(out.real, g[key]) < 1e-12, key


def test_projection_transpose_vs_hermitian(dv):
    """implementation.py:180 uses q.T (no conjugate); conj=True is the north star's Q^H variant."""
    rng = np.random.default_rng(2)
    ct, tt = synthetic.waveguide_operators(4, 3, 40)
    n, r = ct.shape[0], 24
    q = np.linalg.qr(rng.standard_normal((n, r)) + 1j * rng.standard_normal((n, r)))[0]
    at = dv.csr_of_transpose(ct)
    qd = dv.to_device_c128(q)
    y = dv.spmm(at, qd)
    plain = dv.gemm_tn(y, qd, conj=False).cpu().numpy()
    herm = dv.gemm_tn(qd, dv.spmm(dv.csr_of_transpose(csc_array(ct.T)), qd), conj=True).cpu().numpy()
    assert orc.rel_err(plain, (q.T @ ct) @ q) < 1e-13
    assert orc.rel_err(herm, q.conj().T @ (ct @ q)) < 1e-13


# ------------------------------------------------------------------------------------------- chained path
def test_chained_path_stages_fixture(dv):
    from morfem_b200 import implementation as impl, test_helpers as th
    g = np.load(os.path.join(GOLDEN, "stages_n600.npz"))
    in_c, in_gamma, in_b = operators_from(g)
    keep = 10          # directions above 1e-10 relative singular value
    qd, info = dv.orthonormalize(dv.to_device_c128(g["snapshots"][:, :keep]))
    q = qd.cpu().numpy().real
    q_ref = orc.orthonormal_basis(g["snapshots"][:, :keep])
    # any rounding-level perturbation of S moves span(S) by eps * cond(S) (cond ~ 7e8 for these true snapshots)
    span_tol = max(1e-10, 50 * np.finfo(float).eps * info.sigma[0] / info.sigma[-1])
    assert orc.subspace_residual(q_ref, q) < span_tol
    md = impl.ModelDefinition(g["f"], in_c, csc_array(in_c.shape), in_gamma, in_b, lambda t: 1.0, lambda t: t, lambda t: t ** 2,
                              th.b_coefficient)
    ops = impl._DeviceOperators(md)
    a0_r, a1_r, a2_r, b_r = ops.project(qd)
    ref = orc.galerkin_projection(q_ref, in_c, md.a1, in_gamma, in_b)
    # reduced matrices: express the reference's in the new basis
    assert orc.rel_err(a0_r.cpu().numpy().real, orc.align_reduced(ref[0], q_ref, q)) < max(1e-10, 10 * span_tol)
    assert orc.rel_err(a2_r.cpu().numpy().real, orc.align_reduced(ref[2], q_ref, q)) < max(1e-10, 10 * span_tol)
    # stage isolated (the device's own q on both sides): the north star's 1e-10 on reduced matrices
    iso = orc.galerkin_projection(q, in_c, md.a1, in_gamma, in_b)
    assert orc.rel_err(a0_r.cpu().numpy().real, iso[0]) < 1e-10 and orc.rel_err(a2_r.cpu().numpy().real, iso[2]) < 1e-10
    res = impl._sweep_device(g["f"], [a0_r, a1_r, a2_r], b_r, md.t_a0, md.t_a1, md.t_a2, md.t_b, want_x=True, want_gsm=True)
    x_ref = orc.reduced_sweep(g["f"], *ref, md.t_a0, md.t_a1, md.t_a2, md.t_b)
    s_ref = orc.scattering_sweep(g["f"], x_ref, ref[3])
    assert orc.rel_err(res.gsm.cpu().numpy(), s_ref) < 1e-9       # basis-invariant output of the whole chain
    lifted = np.einsum("nr,frm->fnm", q, res.x.cpu().numpy().real)
    lifted_ref = np.einsum("nr,frm->fnm", q_ref, x_ref)
    assert orc.rel_err(lifted, lifted_ref) < 1e-8


def test_public_api_equally_distributed_matches_live_reference(dv):
    """morfem() end to end in USE_EQUALLY_DISTRIBUTED mode (implementation.py:197-214): same 6-tuple contract."""
    from morfem_b200 import implementation as impl, test_helpers as th
    g = np.load(os.path.join(GOLDEN, "equidist_n600.npz"))
    in_c, in_gamma, in_b = operators_from(g)
    impl.USE_EQUALLY_DISTRIBUTED = True
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            x, q, a0_r, a1_r, a2_r, b_r = impl.morfem(g["f"], in_c, csc_array(in_c.shape), in_gamma, in_b, t_b=th.b_coefficient)
    finally:
        impl.USE_EQUALLY_DISTRIBUTED = False
    assert x.shape == g["x"].shape and x.dtype == np.float64 and x.flags.c_contiguous
    assert q.shape == g["q"].shape and q.dtype == np.float64 and q.flags.c_contiguous
    assert a1_r.shape == g["a1_r"].shape and not np.any(a1_r)
    assert b_r.shape == g["b_r"].shape
    assert np.linalg.norm(q.T @ q - np.eye(q.shape[1])) < 1e-12
    cb = np.array([th.b_coefficient(t) for t in g["f"]])
    s = th.scattering_sweep(g["f"], x, b_r, cb)
    # the 12-column snapshot block has near-null directions, so compare the basis-invariant S-parameters
    assert orc.rel_err(s, g["gsm"]) < 1e-6


def test_public_api_greedy_driver_matches_live_reference(dv):
    """finite_element_method_model_order_reduction_gsm on the N=3411 surrogate with the shipped WP values
    (test_helpers.py:53-67, main.py:40): greedy basis on the device estimator, fused sweep + S-parameters."""
    from morfem_b200 import test_helpers as th
    g = np.load(os.path.join(GOLDEN, "cfg1_rom3411.npz"))
    nx, ny, nz = (int(v) for v in g["grid"])
    ct, tt = synthetic.waveguide_operators(nx, ny, nz)
    wp = np.zeros((ct.shape[0], 2))
    wp[g["wp_shipped_nz_rows"], g["wp_shipped_nz_cols"]] = g["wp_shipped_nz_vals"]
    in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, csc_array(wp))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gsm = th.finite_element_method_model_order_reduction_gsm(g["f"], 2, in_c, in_gamma, in_b)
    assert gsm.shape == (100, 2, 2) and gsm.dtype == np.complex128
    err_rom = np.linalg.norm((gsm - g["gsm_rom"]).reshape(100, -1), axis=1)
    err_full = np.linalg.norm((gsm - g["gsm_full"]).reshape(100, -1), axis=1)
    # both ROMs stop at the same 1e-6 residual threshold; they agree with each other and with the full-order sweep
    # to the accuracy the reference itself reports against its full solve (main.py:42-44, ~1e-7)
    assert err_rom.max() < 1e-5 and err_full.max() < 1e-5


# ------------------------------------------------------------- overlapped stages 1 + 2 and the sharded driver
def test_basis_and_projection_equals_sequential_stages(dv):
    """``q^T A q = w^T (x^T A x) w``: the overlapped evaluation (SVD rotation on a side stream) must reproduce the
    sequential orthonormalize -> project chain and the oracle, stage isolated on the same basis."""
    from morfem_b200 import implementation as impl, test_helpers as th
    g = np.load(os.path.join(GOLDEN, "stages_n600.npz"))
    in_c, in_gamma, in_b = operators_from(g)
    keep = 10
    md = impl.ModelDefinition(g["f"], in_c, csc_array(in_c.shape), in_gamma, in_b, lambda t: 1.0, lambda t: t, lambda t: t ** 2, th.b_coefficient)
    ops = impl._DeviceOperators(md)
    sd = dv.to_device_c128(g["snapshots"][:, :keep])
    q, reduced, b_r, info = dv.basis_and_projection(sd, ops.project_block)
    torch.cuda.synchronize()
    qh = q.cpu().numpy().real
    assert np.linalg.norm(qh.T @ qh - np.eye(keep)) < 1e-12
    assert reduced[1] is None and info.kept == keep and info.passes >= 2
    ref = orc.galerkin_projection(qh, in_c, md.a1, in_gamma, in_b)          # implementation.py:180-184 on the SAME q
    for new, old in ((reduced[0], ref[0]), (reduced[2], ref[2]), (b_r, ref[3])):
        assert orc.rel_err(new.cpu().numpy(), old) < 1e-10
    seq = ops.project(q)
    for new, old in ((reduced[0], seq[0]), (reduced[2], seq[2]), (b_r, seq[3])):
        assert orc.rel_err(new.cpu().numpy(), old.cpu().numpy()) < 1e-10
    sig = np.linalg.svd(g["snapshots"][:, :keep], compute_uv=False)
    assert np.max(np.abs(info.sigma - sig)) < 1e-12 * sig[0]


def test_sharded_hot_path_single_rank_matches_oracle(dv):
    """``dist.ShardedHotPath`` (the call bench.py times) on one rank: S-parameters against the oracle's chained path."""
    from morfem_b200 import dist as mfd
    from scipy.constants import pi, epsilon_0
    nx, ny, nz = 6, 5, 60
    ct, tt = synthetic.waveguide_operators(nx, ny, nz)
    n = ct.shape[0]
    wp = synthetic.port_matrix(n, 2, 19)
    in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, wp)
    f = synthetic.frequency_points(64)
    s = synthetic.snapshot_matrix(n, 12, seed=3, decay_decades=4.0)
    cb = np.array([orc.b_coefficient(t) for t in f])
    path = mfd.ShardedHotPath([in_c, csc_array(in_c.shape), in_gamma], in_b, n, f.size, [np.ones_like(f), f, f ** 2, cb, 2 * pi * f * epsilon_0])
    gsm, q, (a0_r, a1_r, a2_r, b_r), res = path.step(dv.to_device_c128(s), want_x=True)
    torch.cuda.synchronize()
    qh = q.cpu().numpy().real
    ref = orc.galerkin_projection(qh, in_c, csc_array(in_c.shape), in_gamma, in_b)
    x_ref = orc.reduced_sweep(f, ref[0], ref[1], ref[2], ref[3], lambda t: 1.0, lambda t: t, lambda t: t ** 2, orc.b_coefficient)
    s_ref = orc.scattering_sweep(f, x_ref, ref[3])
    cond = np.array([np.linalg.cond(orc.system_matrix(1.0, t, t ** 2, ref[0], ref[1], ref[2])) for t in f])
    tol = np.maximum(1e-10, 50 * np.finfo(float).eps * cond)
    err = np.linalg.norm((gsm.cpu().numpy() - s_ref).reshape(f.size, -1), axis=1) / np.linalg.norm(s_ref.reshape(f.size, -1), axis=1)
    assert np.all(err < tol), (err.max(), tol.max())
    assert not np.any(res.info.cpu().numpy())


def test_graph_replay_equals_eager_step_and_falls_back(dv):
    """The CUDA-graph replay (optimistic CholeskyQR2, flags verified afterwards) gives the eager step's S-parameters;
    an ill-conditioned block (needs a shifted third pass) fails verification and is re-run on the adaptive path."""
    from morfem_b200 import dist as mfd
    from scipy.constants import pi, epsilon_0
    nx, ny, nz = 6, 5, 60
    ct, tt = synthetic.waveguide_operators(nx, ny, nz)
    n = ct.shape[0]
    in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, synthetic.port_matrix(n, 2, 19))
    f = synthetic.frequency_points(48)
    cb = np.array([orc.b_coefficient(t) for t in f])
    path = mfd.ShardedHotPath([in_c, csc_array(in_c.shape), in_gamma], in_b, n, f.size, [np.ones_like(f), f, f ** 2, cb, 2 * pi * f * epsilon_0])
    sd = dv.to_device_c128(synthetic.snapshot_matrix(n, 12, seed=3, decay_decades=4.0))
    eager = path.step(sd, gather=False)[0].clone()
    for _ in range(3):
        replay = path.step_graph(sd)[0]
    assert path.verify() is None and path.launches_per_graph > 20
    assert orc.rel_err(replay.cpu().numpy(), eager.cpu().numpy()) < 1e-10
    rng = np.random.default_rng(4)      # nearly collinear columns (cond ~ 1e10 even after column scaling): CholeskyQR2 is not enough
    bad = dv.to_device_c128(np.outer(rng.standard_normal(n), np.ones(12)) + 1e-10 * rng.standard_normal((n, 12)))
    path.step_graph(bad)
    redo = path.verify()
    assert redo is not None
    assert orc.rel_err(redo[0].cpu().numpy(), path.step(bad, gather=False)[0].cpu().numpy()) < 1e-9


def test_real_float64_stages_equal_the_complex_path(dv):
    """Row N2: for real snapshots and operators the float64 twins of stages 1 + 2 give the complex128 path's reduced model
    (same algorithm on 8-byte elements; every imaginary part of the complex run is an exact zero)."""
    from morfem_b200 import implementation as impl, test_helpers as th
    g = np.load(os.path.join(GOLDEN, "stages_n600.npz"))
    in_c, in_gamma, in_b = operators_from(g)
    keep = 10
    md = impl.ModelDefinition(g["f"], in_c, csc_array(in_c.shape), in_gamma, in_b, lambda t: 1.0, lambda t: t, lambda t: t ** 2, th.b_coefficient)
    ops = impl._DeviceOperators(md)
    snaps = np.ascontiguousarray(g["snapshots"][:, :keep])
    qc, rc, bc, _ = dv.basis_and_projection(dv.to_device_c128(snaps), ops.project_block)
    qr, rr, br, info = dv.basis_and_projection(dv.real_or_complex_to_device(snaps), ops.project_block)
    torch.cuda.synchronize()
    assert qr.dtype == torch.float64 and rr[0].dtype == torch.float64 and br.dtype == torch.float64
    assert orc.subspace_residual(qc.cpu().numpy().real, qr.cpu().numpy()) < 1e-9
    # reduced operators live in each run's own basis; the S-parameters are basis invariant
    f = g["f"]
    sc = impl._sweep_device(f, list(rc), bc, md.t_a0, md.t_a1, md.t_a2, md.t_b, want_x=False, want_gsm=True).gsm.cpu().numpy()
    sr = impl._sweep_device(f, list(rr), br, md.t_a0, md.t_a1, md.t_a2, md.t_b, want_x=False, want_gsm=True).gsm.cpu().numpy()
    assert orc.rel_err(sr, sc) < 1e-8
    ref = orc.galerkin_projection(qr.cpu().numpy(), in_c, md.a1, in_gamma, in_b)
    for new, old in ((rr[0], ref[0]), (rr[2], ref[2]), (br, ref[3])):
        assert orc.rel_err(new.cpu().numpy().real, old) < 1e-10
    # end-to-end helper: automatic real path == forced complex path
    s_auto = th.model_order_reduction_gsm_from_snapshots(f, snaps, in_c, in_gamma, in_b)
    s_cplx = th.model_order_reduction_gsm_from_snapshots(f, snaps, in_c, in_gamma, in_b, real_path=False)
    assert orc.rel_err(s_auto, s_cplx) < 1e-8


@pytest.mark.parametrize("r", [113, 130, 200, 256, 300])
def test_blocked_cholesky_and_inverse_for_large_r(dv, r):
    """r > 112: the r x r Cholesky and triangular inverse run as blocked drivers over the shared-memory kernels."""
    rng = np.random.default_rng(r)
    a = rng.standard_normal((r + 40, r)) + 1j * rng.standard_normal((r + 40, r))
    g = a.conj().T @ a
    gd = dv.to_device_c128(g)
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    dv._potrf_upper(gd, info)
    rinv = dv._trtri_upper(gd)
    torch.cuda.synchronize()
    rr = gd.cpu().numpy()
    assert int(info.item()) == 0 and np.allclose(np.tril(rr, -1), 0)
    assert orc.rel_err(rr.conj().T @ rr, g) < 1e-13
    assert orc.rel_err(rinv.cpu().numpy() @ rr, np.eye(r)) < 1e-11
    bad = g.copy()
    bad[150 if r > 150 else 100, :] = 0; bad[:, 150 if r > 150 else 100] = 0           # not positive definite from that column on
    bd = dv.to_device_c128(bad)
    dv._potrf_upper(bd, info)
    assert int(info.item()) == (151 if r > 150 else 101)


@pytest.mark.parametrize("r", [1, 5, 31, 32, 33, 64, 100, 150, 256, 300, 512])
def test_fused_cholesky_inverse_kernel(dv, r):
    """mf_chol_inv_upper_c128 (one cooperative launch): G = R^H R, R^-1, zeros below the diagonal of both, LAPACK-style
    info for a matrix that stops being positive definite, ragged last block, strided operands."""
    from morfem_b200 import _ffi
    lib = _ffi.load()
    rng = np.random.default_rng(1000 + r)
    a = rng.standard_normal((r + 40, r)) + 1j * rng.standard_normal((r + 40, r))
    g = a.conj().T @ a
    wide = torch.zeros((r, r + 3), dtype=torch.complex128, device="cuda")          # leading dimension r + 3
    gd = wide[:, :r]
    gd.copy_(dv.to_device_c128(g))
    rinv = torch.full((r, r), 7.0, dtype=torch.complex128, device="cuda")          # stale content must be overwritten
    info = torch.full((1,), -1, dtype=torch.int32, device="cuda")
    ws = torch.empty(lib.mf_chol_inv_ws_bytes(r), dtype=torch.uint8, device="cuda")
    run = lambda m: _ffi.check(lib.mf_chol_inv_upper_c128(dv._ptr(m), m.stride(0), r, dv._ptr(rinv), rinv.stride(0), dv._ptr(info),
                                                          dv._ptr(ws), ws.numel(), dv._stream()), "mf_chol_inv_upper_c128")
    run(gd)
    torch.cuda.synchronize()
    rr, ri = gd.cpu().numpy(), rinv.cpu().numpy()
    assert int(info.item()) == 0
    assert np.all(np.tril(rr, -1) == 0) and np.all(np.tril(ri, -1) == 0)
    assert np.all(np.diag(rr).real > 0) and np.abs(np.diag(rr).imag).max() == 0
    assert orc.rel_err(rr.conj().T @ rr, g) < 1e-13
    ref = np.linalg.cholesky(g).conj().T                                           # LAPACK potrf: the same factor
    assert orc.rel_err(rr, ref) < 1e-11
    assert orc.rel_err(ri @ rr, np.eye(r)) < 1e-10
    first = rr.copy()
    gd.copy_(dv.to_device_c128(g)); run(gd); torch.cuda.synchronize()
    assert np.array_equal(gd.cpu().numpy(), first)                                 # deterministic (replicated across ranks)
    if r >= 5:
        k = (2 * r) // 3
        bad = g.copy(); bad[k, :] = 0; bad[:, k] = 0
        bd = dv.to_device_c128(bad)
        run(bd)
        assert int(info.item()) == k + 1


def test_e2e_helper_falls_back_when_choleskyqr2_is_not_enough(dv):
    """``model_order_reduction_gsm_from_snapshots`` runs the optimistic CholeskyQR2 and verifies its flags with the result
    download; a nearly collinear snapshot block must take the adaptive (shifted, three-pass) path and still match the oracle."""
    from morfem_b200 import test_helpers as th
    nx, ny, nz = 6, 5, 60
    ct, tt = synthetic.waveguide_operators(nx, ny, nz)
    n = ct.shape[0]
    in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, synthetic.port_matrix(n, 2, 19))
    f = synthetic.frequency_points(32)
    rng = np.random.default_rng(5)
    base = synthetic.snapshot_matrix(n, 6, seed=2, decay_decades=2.0)
    snaps = np.hstack([base, base[:, :3] + 1e-9 * rng.standard_normal((n, 3))])        # cond ~ 1e9 after column scaling
    gsm = th.model_order_reduction_gsm_from_snapshots(f, snaps, in_c, in_gamma, in_b)
    q_ref, a0, a1, a2, b_r, x_ref, s_ref = orc.hot_path(snaps, f, in_c, csc_array(in_c.shape), in_gamma, in_b)
    cond = np.array([np.linalg.cond(orc.system_matrix(1.0, t, t ** 2, a0, a1, a2)) for t in f])
    err = np.linalg.norm((gsm - s_ref).reshape(f.size, -1), axis=1) / np.linalg.norm(s_ref.reshape(f.size, -1), axis=1)
    assert np.all(err < np.maximum(1e-7, 1e3 * np.finfo(float).eps * cond)), err.max()


def test_basis_size_study_config4(dv):
    """BASELINE config 4 (the reference's speed_and_error_of_no_points_in_q experiment through the current API):
    S-parameter error of the reduced sweep against the full-order sweep as the number of equally spaced snapshot points
    grows.  The GPU path must reproduce the CPU restatement of the same experiment size by size and converge like it."""
    import importlib.util, os
    from morfem_b200 import synthetic
    spec = importlib.util.spec_from_file_location("basis_size_sweep", os.path.join(os.path.dirname(__file__), "..", "examples", "basis_size_sweep.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ct, tt = synthetic.waveguide_operators(6, 4, 40)
    n = ct.shape[0]
    in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, synthetic.port_matrix(n, 2, 19))
    f = np.linspace(3e9, 5e9, 41)
    sizes = [2, 3, 4, 5, 6]          # cond(snapshot block) 1e2 .. 1e9: the last sizes take the shifted Cholesky-QR path
    rows = mod.basis_size_study(f, in_c, in_gamma, in_b, sizes)
    assert [r["snapshot_points"] for r in rows] == sizes and [r["r"] for r in rows] == [2 * k for k in sizes]
    # the oracle's version of the experiment: same snapshots -> svd basis -> projection -> sweep -> S-parameters (all CPU)
    from scipy.sparse import csc_array
    full = orc.full_order_sweep(f, in_c, csc_array(in_c.shape), in_gamma, in_b, lambda t: 1.0, lambda t: t, lambda t: t ** 2, orc.b_coefficient)
    ref_gsm = orc.scattering_sweep(f, full, in_b.toarray())
    errs_cpu = []
    for k in sizes:
        idx = np.linspace(0, f.size - 1, k, dtype=int)
        snaps = np.concatenate([full[i] for i in idx], axis=1)
        q = orc.orthonormal_basis(snaps)
        a0_r, a1_r, a2_r, b_r = orc.galerkin_projection(q, in_c, csc_array(in_c.shape), in_gamma, in_b)
        x = orc.reduced_sweep(f, a0_r, a1_r, a2_r, b_r, lambda t: 1.0, lambda t: t, lambda t: t ** 2, orc.b_coefficient)
        gsm = orc.scattering_sweep(f, x, b_r)
        errs_cpu.append(np.mean([np.linalg.norm(gsm[i] - ref_gsm[i]) for i in range(f.size)]))
    errs_gpu = [r["error_mean"] for r in rows]
    for eg, ec in zip(errs_gpu, errs_cpu):
        assert abs(eg - ec) <= 0.05 * ec + 5e-10, (errs_gpu, errs_cpu)      # same error curve (7e-6, 2e-8, 7e-11, ... on the CPU)
    assert errs_gpu[0] > errs_gpu[1] > errs_gpu[2] and errs_gpu[-1] < 1e-9


def test_error_estimator_matches_live_reference_values():
    """``error_estimator`` (implementation.py:348-452) evaluated on the device -- Gram matrices of the SpMM
    outputs instead of the reference's sparse-sparse products, the reduced sweep, ``mf_estimator_c128`` -- against the
    per-point estimates of the LIVE reference on the bases of a greedy run (tests/golden/estimator_n600.npz).  1e-8 relative
    where the estimate is above its own cancellation floor (a few eps times the size of the largest of its 16 terms), and
    the greedy search's arg-max point must agree."""
    from scipy.sparse import csc_array
    from morfem_b200 import implementation as impl, test_helpers as th, synthetic
    g = np.load(os.path.join(GOLDEN, "estimator_n600.npz"))
    eps = np.finfo(float).eps
    for tag, ports, iters in (("p2", 2, 4), ("p3", 3, 2)):
        ct, tt = synthetic.waveguide_operators(*(int(v) for v in g["grid"]))
        wp = synthetic.port_matrix(ct.shape[0], ports, int(g["face"]))
        in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, wp)
        f = g[tag + "_f"]
        md = impl.ModelDefinition(f, in_c, csc_array(in_c.shape), in_gamma, in_b, lambda t: 1, lambda t: t, lambda t: t ** 2,
                                  lambda t: th.b_coefficient(t))
        for it in range(iters):
            err = impl.error_estimator(md, g[f"{tag}_q{it}"])
            ref, scale = g[f"{tag}_err{it}"], g[f"{tag}_scale{it}"]
            assert err.shape == ref.shape
            assert np.all(np.abs(err - ref) <= 1e-8 * ref + 200 * eps * scale), (tag, it, float(np.abs(err - ref).max()))
            assert int(err.argmax()) == int(ref.argmax())


def test_incremental_greedy_search_matches_live_reference_opm_mode(monkeypatch):
    """``USE_OPM = True`` (implementation.py:16, :230-295): the greedy search that orthonormalises, multiplies and projects only the
    new columns and grows the estimator blocks, against the live reference run in the same mode (tests/golden/opm_n600.npz): the
    estimator curve of every iteration (1e-8 above its cancellation floor), the same points picked, the same final basis size,
    and ROM S-parameters as close to the full-order ones as the reference's."""
    from scipy.sparse import csc_array
    from morfem_b200 import implementation as impl, test_helpers as th, synthetic
    g = np.load(os.path.join(GOLDEN, "opm_n600.npz"))
    ct, tt = synthetic.waveguide_operators(*(int(v) for v in g["grid"]))
    wp = synthetic.port_matrix(ct.shape[0], int(g["ports"]), int(g["face"]))
    in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, wp)
    f = g["f"]
    monkeypatch.setattr(impl, "USE_OPM", True)
    gsm = th.finite_element_method_model_order_reduction_gsm(f, 2, in_c, in_gamma, in_b)
    errs = impl.projection_base.last_errors
    eps = np.finfo(float).eps
    assert len(errs) == g["errors"].shape[0]
    for it, (err, ref) in enumerate(zip(errs, g["errors"])):
        assert np.all(np.abs(err - ref) <= 1e-8 * ref + 200 * eps * g["scale"]), (it, float(np.abs(err - ref).max()))
        assert int(err.argmax()) == int(ref.argmax())
    d_new = np.linalg.norm((gsm - g["gsm_full"]).reshape(f.size, -1), axis=1)
    d_ref = np.linalg.norm((g["gsm_rom"] - g["gsm_full"]).reshape(f.size, -1), axis=1)
    assert d_new.max() <= max(2 * d_ref.max(), 1e-9)
    # the non-incremental default walks through the same points (same estimator, whole-basis re-orthonormalisation instead)
    monkeypatch.setattr(impl, "USE_OPM", False)
    gsm2 = th.finite_element_method_model_order_reduction_gsm(f, 2, in_c, in_gamma, in_b)
    assert np.abs(gsm2 - gsm).max() < 1e-6
