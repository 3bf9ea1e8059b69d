"""Pin the CPU oracle (oracle/reference_path.py) against outputs of the LIVE reference.

The fixtures in tests/golden/*.npz were produced by tests/golden/make_golden.py, which imports the unmodified
reference from /root/reference (the reference has no golden vectors of its own: SURVEY.md section 8c).  The oracle
calls the same numpy/scipy entry points at the same call sites, so on the machine that generated the fixtures the
agreement is to rounding; tolerances leave room for a different BLAS on another host."""
import glob
import os

import numpy as np
import pytest
from scipy.sparse import csc_array

from oracle import reference_path as orc
from morfem_b200 import synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


REDUCED = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "reduced_r*.npz")))


def sweep_tol(cond):
    """Forward-error allowance of an LU solve: a few hundred ulps times the condition number of the worst point."""
    return max(1e-10, 200 * np.finfo(float).eps * float(np.max(cond)))


@pytest.mark.parametrize("name", REDUCED)
def test_reduced_sweep_and_gsm_match_live_reference(name):
    g = load(name)
    f = g["f"]
    x = orc.reduced_sweep(f, g["a0"], g["a1"], g["a2"], g["b"], lambda t: 1.0, lambda t: t, lambda t: t ** 2, orc.b_coefficient)
    assert x.dtype == np.float64 and x.shape == g["x"].shape
    assert orc.rel_err(x, g["x"]) < sweep_tol(g["cond"])
    gsm = orc.scattering_sweep(f, g["x"], g["b"])          # stage 4 isolated: fed the reference's own x
    assert gsm.dtype == np.complex128
    assert orc.rel_err(gsm, g["gsm"]) < 1e-12


def operators_from(g):
    nx, ny, nz = (int(v) for v in g["grid"])
    ct, tt = synthetic.waveguide_operators(nx, ny, nz)
    wp = synthetic.port_matrix(ct.shape[0], int(g["ports"]), int(g["face"]))
    return synthetic.driver_scaled(ct, tt, wp)


def test_stage_chain_matches_live_reference():
    g = load("stages_n600")
    in_c, in_gamma, in_b = operators_from(g)
    q, a0_r, a1_r, a2_r, b_r, x, gsm = orc.hot_path(g["snapshots"], g["f"], in_c, csc_array(in_c.shape), in_gamma, in_b)
    # identical LAPACK call on identical input: columns agree to rounding where the singular values are resolved
    keep = 10
    assert np.max(np.abs(np.abs(np.sum(q[:, :keep] * g["q"][:, :keep], axis=0)) - 1.0)) < 1e-6
    assert orc.subspace_residual(g["q"][:, :keep], q[:, :keep]) < 1e-6
    # stage 2 isolated (fed the reference's q)
    p = orc.galerkin_projection(g["q"], in_c, csc_array(in_c.shape), in_gamma, in_b)
    for new, key in zip(p, ("a0_r", "a1_r", "a2_r", "b_r")):
        assert orc.rel_err(new, g[key]) < 1e-13, key
    assert not np.any(p[1])
    # stage 3 isolated (fed the reference's reduced operators)
    xs = orc.reduced_sweep(g["f"], g["a0_r"], g["a1_r"], g["a2_r"], g["b_r"], lambda t: 1.0, lambda t: t, lambda t: t ** 2,
                           orc.b_coefficient)
    lifted_new = np.einsum("nr,frm->fnm", g["q"], xs)
    lifted_ref = np.einsum("nr,frm->fnm", g["q"], g["x"])
    assert orc.rel_err(lifted_new, lifted_ref) < 1e-8
    # stage 4 isolated
    assert orc.rel_err(orc.scattering_sweep(g["f"], g["x"], g["b_r"]), g["gsm"]) < 1e-12
    # chained S-parameters are basis invariant
    assert orc.rel_err(gsm, g["gsm"]) < 1e-7


def test_equally_distributed_fixture_is_consistent():
    g = load("equidist_n600")
    in_c, in_gamma, in_b = operators_from(g)
    p = orc.galerkin_projection(g["q"], in_c, csc_array(in_c.shape), in_gamma, in_b)
    for new, key in zip(p, ("a0_r", "a1_r", "a2_r", "b_r")):
        assert orc.rel_err(new, g[key]) < 1e-13, key
    assert orc.rel_err(orc.scattering_sweep(g["f"], g["x"], g["b_r"]), g["gsm"]) < 1e-12


def test_cfg1_fixture_shipped_port_matrix_and_svd_pair():
    g = load("cfg1_rom3411")
    wp = np.zeros((3411, 2))
    wp[g["wp_shipped_nz_rows"], g["wp_shipped_nz_cols"]] = g["wp_shipped_nz_vals"]
    assert np.abs(wp - synthetic.shipped_port_matrix().toarray()).max() < 2e-7
    u = orc.orthonormal_basis(g["svd_in"])
    s = orc.singular_values(g["svd_in"])
    keep = int(np.count_nonzero(s > 1e-10 * s[0]))
    assert orc.subspace_residual(g["svd_out"][:, :keep], u[:, :keep]) < 1e-6
    # ROM and full-order S-parameters of the live reference agree to the greedy threshold's order (main.py:42-44)
    err = np.linalg.norm((g["gsm_rom"] - g["gsm_full"]).reshape(100, -1), axis=1)
    assert err.max() < 1e-5
    # lossless two-port: |S11|^2 + |S21|^2 = 1
    assert np.abs(np.abs(g["gsm_rom"][:, 0, 0]) ** 2 + np.abs(g["gsm_rom"][:, 1, 0]) ** 2 - 1.0).max() < 1e-6


def test_full_order_sweep_matches_live_reference():
    """``oracle.full_order_sweep`` (the sparse SuperLU branch, implementation.py:472-475) against the live reference's
    ``finite_element_method_gsm`` (test_helpers.py:25-50) on the config-1 model with the shipped port matrix."""
    from scipy.sparse import csc_array
    g = load("cfg1_rom3411")
    nx, ny, nz = (int(v) for v in g["grid"])
    ct, tt = synthetic.waveguide_operators(nx, ny, nz)
    wp = np.zeros((ct.shape[0], 2))
    wp[g["wp_shipped_nz_rows"], g["wp_shipped_nz_cols"]] = g["wp_shipped_nz_vals"]
    in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, csc_array(wp))
    f = g["f"]
    x = orc.full_order_sweep(f, in_c, csc_array(in_c.shape), in_gamma, in_b, lambda t: 1.0, lambda t: t, lambda t: t ** 2, orc.b_coefficient)
    gsm = orc.scattering_sweep(f, x, in_b.toarray())
    assert x.shape == (f.size, ct.shape[0], 2)
    assert orc.rel_err(gsm, g["gsm_full"]) < 1e-9


def test_b_coefficient_raises_below_cutoff():
    with pytest.raises(ValueError):
        orc.b_coefficient(2.0e9)         # test_helpers.py:72: math.sqrt of a negative number
    assert orc.b_coefficient(3.0e9) > 0


def test_comparison_helpers():
    rng = np.random.default_rng(0)
    q, _ = np.linalg.qr(rng.standard_normal((50, 6)))
    signs = np.array([1, -1, 1, -1, -1, 1.0])
    assert orc.subspace_residual(q, q * signs) < 1e-14
    assert np.max(orc.principal_angles(q, q * signs)) < 1e-7
    a = rng.standard_normal((50, 50))
    a_r = q.T @ a @ q
    q2 = q @ np.linalg.qr(rng.standard_normal((6, 6)))[0]
    assert orc.rel_err(orc.align_reduced(a_r, q, q2), q2.T @ a @ q2) < 1e-13


def test_error_estimator_matches_live_reference(golden_dir):
    """The oracle's restatement of implementation.py:348-452 against the per-point estimates of the live reference, on
    the bases a greedy run passes through.  The estimate is a difference of terms of size ``scale``: it is defined to a few
    eps * scale, which is all that is left at the snapshot points themselves."""
    from scipy.sparse import csc_array
    from morfem_b200 import synthetic
    g = np.load(os.path.join(golden_dir, "estimator_n600.npz"))
    eps = np.finfo(float).eps
    for tag, ports, iters in (("p2", 2, 4), ("p3", 3, 2)):
        ct, tt = synthetic.waveguide_operators(*(int(v) for v in g["grid"]))
        wp = synthetic.port_matrix(ct.shape[0], ports, int(g["face"]))
        in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, wp)
        f = g[tag + "_f"]
        for it in range(iters):
            err = orc.error_estimator(g[f"{tag}_q{it}"], f, in_c, csc_array(in_c.shape), in_gamma, in_b,
                                      lambda t: 1, lambda t: t, lambda t: t ** 2, orc.b_coefficient)
            ref, scale = g[f"{tag}_err{it}"], g[f"{tag}_scale{it}"]
            assert np.all(np.abs(err - ref) <= 1e-10 * ref + 50 * eps * scale), (tag, it, np.abs(err - ref).max())
            assert int(err.argmax()) == int(ref.argmax())
