"""GPU parity of stages 3 + 4 (batched reduced solves + S-parameters) through the C ABI.

Checked against (a) the committed outputs of the live reference (tests/golden/reduced_r*.npz) and (b) the CPU
oracle on seeded inputs.  Tolerance: the north star's 1e-10 relative on S-parameters and solutions, widened only
by the conditioning of the point's system matrix (an LU solve cannot be reproduced below ~eps*cond(A) by *any*
other evaluation order; the fixtures record cond(A(t)) for every point)."""
import glob
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import reference_path as orc   # noqa: E402  (checker only)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REDUCED = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "reduced_r*.npz")))
EPS = np.finfo(float).eps


@pytest.fixture(scope="module")
def dv():
    from morfem_b200 import device
    device.require_cuda()
    return device


def run_sweep(dv, f, a0, a1, a2, b, cb, variant=0, want_x=True, want_gsm=True):
    from scipy.constants import pi, epsilon_0
    dev = dv.require_cuda()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)  # noqa: E731
    ops = [None if (a is None or not np.any(a)) else dv.symmetrize(dv.to_device_c128(a)) for a in (a0, a1, a2)]
    res = dv.sweep(ops[0], ops[1], ops[2], dv.to_device_c128(b), t(np.ones_like(f)), t(f), t(f ** 2), t(cb), t(2 * pi * f * epsilon_0),
                   want_x=want_x, want_gsm=want_gsm, variant=variant)
    torch.cuda.synchronize()
    return res


def per_point_rel(new, ref):
    new = np.asarray(new).reshape(new.shape[0], -1)
    ref = np.asarray(ref).reshape(ref.shape[0], -1)
    return np.linalg.norm(new - ref, axis=1) / np.linalg.norm(ref, axis=1)


def variants_for(dv, r, m):
    from morfem_b200 import _ffi
    lib = _ffi.load()
    return [v for v in (1, 2, 3, 4, 5) if lib.mf_sweep_variant_supported(r, m, v)]


@pytest.mark.parametrize("name", REDUCED)
def test_sweep_matches_live_reference_fixture(dv, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    f = g["f"]
    cb = np.array([orc.b_coefficient(t) for t in f])
    r, m = g["b"].shape
    for variant in [0] + variants_for(dv, r, m):
        res = run_sweep(dv, f, g["a0"], g["a1"], g["a2"], g["b"], cb, variant=variant)
        x = res.x.cpu().numpy()
        s = res.gsm.cpu().numpy()
        assert not np.any(res.info.cpu().numpy())
        assert np.abs(x.imag).max() == 0.0                       # real data stays exactly real
        tol = np.maximum(1e-10, 20 * EPS * g["cond"])
        ex, es = per_point_rel(x.real, g["x"]), per_point_rel(s, g["gsm"])
        assert np.all(ex < tol), (name, variant, ex.max(), tol.max())
        assert np.all(es < tol), (name, variant, es.max(), tol.max())


@pytest.mark.parametrize("r,m,nf", [(1, 1, 3), (2, 2, 5), (7, 3, 33), (16, 16, 9), (31, 5, 40), (48, 2, 300), (100, 8, 12), (112, 16, 5),
                                    (113, 1, 4), (128, 4, 20), (160, 3, 500), (200, 2, 6), (256, 4, 10), (257, 9, 3), (384, 16, 2), (512, 4, 150)])
def test_sweep_matches_oracle_on_seeded_models(dv, r, m, nf):
    from morfem_b200 import synthetic
    a0, a1, a2, b = synthetic.reduced_model(r, m, seed=100 + r)
    rng = np.random.default_rng(r)
    a1 = 1e-12 * np.abs(a0).max() * rng.standard_normal((r, r))      # exercise the a1 term as well
    f = np.linspace(3e9, 5e9, nf)
    tb = orc.b_coefficient
    x_ref = orc.reduced_sweep(f, a0, a1, a2, b, lambda t: 1.0, lambda t: t, lambda t: t ** 2, tb)
    s_ref = orc.scattering_sweep(f, x_ref, b)
    cond = np.array([np.linalg.cond(orc.system_matrix(1.0, t, t ** 2, a0, a1, a2)) for t in f])
    cb = np.array([tb(t) for t in f])
    tol = np.maximum(1e-10, 20 * EPS * cond)
    for variant in [0] + variants_for(dv, r, m):
        res = run_sweep(dv, f, a0, a1, a2, b, cb, variant=variant)
        ex = per_point_rel(res.x.cpu().numpy().real, x_ref)
        es = per_point_rel(res.gsm.cpu().numpy(), s_ref)
        assert np.all(ex < tol), (variant, ex.max())
        assert np.all(es < tol), (variant, es.max())


def test_sweep_complex_operators_match_restatement(dv):
    """Complex (lossy) operators have no reference semantics (implementation.py:190 drops imaginary parts); the
    oracle's complex_ok restatement is the checker."""
    from morfem_b200 import synthetic
    r, m, nf = 40, 3, 25
    a0, a1, a2, b = synthetic.reduced_model(r, m, seed=9, complex_valued=True)
    f = np.linspace(3e9, 5e9, nf)
    tb = orc.b_coefficient
    x_ref = orc.reduced_sweep(f, a0, a1, a2, b, lambda t: 1.0, lambda t: t, lambda t: t ** 2, tb, complex_ok=True)
    s_ref = orc.scattering_sweep(f, x_ref, b)
    cond = np.array([np.linalg.cond(orc.system_matrix(1.0, t, t ** 2, a0, a1, a2)) for t in f])
    tol = np.maximum(1e-10, 20 * EPS * cond)
    cb = np.array([tb(t) for t in f])
    for variant in [0] + variants_for(dv, r, m):
        res = run_sweep(dv, f, a0, a1, a2, b, cb, variant=variant)
        assert np.all(per_point_rel(res.x.cpu().numpy(), x_ref) < tol)
        assert np.all(per_point_rel(res.gsm.cpu().numpy(), s_ref) < tol)


def test_sweep_pivoting_follows_lapack(dv):
    """A matrix that needs row exchanges at every step; first-maximum tie-breaking like idamax."""
    r, m = 12, 2
    rng = np.random.default_rng(4)
    a0 = np.fliplr(np.eye(r)) * 3.0 + 1e-3 * rng.standard_normal((r, r))
    a0 = (a0 + a0.T) / 2
    b = rng.standard_normal((r, m))
    f = np.array([3e9, 4e9])
    zero = np.zeros((r, r))
    x_ref = orc.reduced_sweep(f, a0, zero, zero, b, lambda t: 1.0, lambda t: t, lambda t: t ** 2, lambda t: 1.0)
    for variant in [0] + variants_for(dv, r, m):
        res = run_sweep(dv, f, a0, None, None, b, np.ones(2), variant=variant)
        assert np.all(per_point_rel(res.x.cpu().numpy().real, x_ref) < 1e-12)


def test_sweep_reports_singular_points_like_lu_factor(dv):
    """lu_factor only warns on an exactly singular U (implementation.py:477); the ABI reports the 1-based column."""
    r, m = 6, 2
    a0 = np.eye(r)
    a0[3, 3] = 0.0
    b = np.ones((r, m))
    f = np.array([3e9, 4e9, 5e9])
    for variant in [0] + variants_for(dv, r, m):
        res = run_sweep(dv, f, a0, None, None, b, np.ones(3), variant=variant, want_gsm=False)
        assert list(res.info.cpu().numpy()) == [4, 4, 4]


def test_sweep_singular_and_pivoting_in_the_streamed_kernel(dv):
    """r > 112 runs the streamed two-level kernel: a permutation-like matrix needs exchanges in every outer panel, and an
    exactly singular matrix must report the first zero pivot (here deep inside the third outer panel)."""
    r, m = 150, 2
    rng = np.random.default_rng(8)
    a0 = np.fliplr(np.eye(r)) * 2.0 + 1e-3 * rng.standard_normal((r, r))
    a0 = (a0 + a0.T) / 2
    b = rng.standard_normal((r, m))
    f = np.array([3e9, 4e9, 5e9])
    zero = np.zeros((r, r))
    x_ref = orc.reduced_sweep(f, a0, zero, zero, b, lambda t: 1.0, lambda t: t, lambda t: t ** 2, lambda t: 1.0)
    for variant in [0] + variants_for(dv, r, m):
        res = run_sweep(dv, f, a0, None, None, b, np.ones(3), variant=variant)
        assert np.all(per_point_rel(res.x.cpu().numpy().real, x_ref) < 1e-11), variant
    sing = np.eye(r)
    sing[77, 77] = 0.0
    for variant in [0] + variants_for(dv, r, m):
        res = run_sweep(dv, f, sing, None, None, np.ones((r, m)), np.ones(3), variant=variant, want_gsm=False)
        assert list(res.info.cpu().numpy()) == [78, 78, 78], variant


@pytest.mark.parametrize("r,m,nf", [(40, 3, 7), (130, 5, 301), (256, 4, 297), (250, 8, 2)])
def test_two_point_cta_kernel_matches_oracle(dv, r, m, nf, monkeypatch):
    """sweep_left4_kernel (MF_LEFT_CFG=8: two points per CTA in anti-phase, panel warps and DMMA warps on different SM
    sub-partitions) against the oracle: odd point counts (the second slot of a CTA runs out first), a single CTA, row exchanges in
    every panel and the report of an exactly singular point."""
    from morfem_b200 import synthetic
    monkeypatch.setenv("MF_LEFT_CFG", "8")
    a0, a1, a2, b = synthetic.reduced_model(r, m, seed=300 + r)
    f = np.linspace(3e9, 5e9, nf)
    tb = orc.b_coefficient
    x_ref = orc.reduced_sweep(f, a0, a1, a2, b, lambda t: 1.0, lambda t: t, lambda t: t ** 2, tb)
    s_ref = orc.scattering_sweep(f, x_ref, b)
    cond = np.array([np.linalg.cond(orc.system_matrix(1.0, t, t ** 2, a0, a1, a2)) for t in f])
    cb = np.array([tb(t) for t in f])
    tol = np.maximum(1e-10, 20 * EPS * cond)
    res = run_sweep(dv, f, a0, a1, a2, b, cb, variant=5)
    assert not np.any(res.info.cpu().numpy())
    assert np.all(per_point_rel(res.x.cpu().numpy().real, x_ref) < tol)
    assert np.all(per_point_rel(res.gsm.cpu().numpy(), s_ref) < tol)
    # a permutation-like matrix (exchanges in every panel) and an exactly singular one
    rng = np.random.default_rng(r)
    p0 = np.fliplr(np.eye(r)) * 2.0 + 1e-3 * rng.standard_normal((r, r))
    p0 = (p0 + p0.T) / 2
    f3 = np.array([3e9, 4e9, 5e9])
    zero = np.zeros((r, r))
    xp = orc.reduced_sweep(f3, p0, zero, zero, b, lambda t: 1.0, lambda t: t, lambda t: t ** 2, lambda t: 1.0)
    res = run_sweep(dv, f3, p0, None, None, b, np.ones(3), variant=5)
    assert np.all(per_point_rel(res.x.cpu().numpy().real, xp) < 1e-11)
    sing = np.eye(r)
    sing[r // 2 + 3, r // 2 + 3] = 0.0
    res = run_sweep(dv, f3, sing, None, None, np.ones((r, m)), np.ones(3), variant=5, want_gsm=False)
    assert list(res.info.cpu().numpy()) == [r // 2 + 4] * 3


def test_sweep_outputs_are_optional_and_idempotent(dv):
    from morfem_b200 import synthetic
    r, m, nf = 24, 2, 50
    a0, a1, a2, b = synthetic.reduced_model(r, m, seed=3)
    f = np.linspace(3e9, 5e9, nf)
    cb = np.array([orc.b_coefficient(t) for t in f])
    both = run_sweep(dv, f, a0, a1, a2, b, cb)
    only_s = run_sweep(dv, f, a0, a1, a2, b, cb, want_x=False)
    only_x = run_sweep(dv, f, a0, a1, a2, b, cb, want_gsm=False)
    assert only_s.x is None and only_x.gsm is None
    assert torch.equal(both.gsm, only_s.gsm) and torch.equal(both.x, only_x.x)     # bit-identical, run to run
    # stage 4 alone (mf_gsm_c128) on the sweep's own x reproduces the fused epilogue
    s4 = dv.gsm(both.x, dv.to_device_c128(b), torch.from_numpy(cb).cuda(), torch.from_numpy(2 * np.pi * f * 8.8541878128e-12).cuda())
    assert orc.rel_err(s4.cpu().numpy(), both.gsm.cpu().numpy()) < 1e-9


def test_sweep_large_batch_property(dv):
    """At bench scale (r=64, 10k points) check the size-independent property A(t) x = b(t) on a sample of points."""
    from morfem_b200 import synthetic
    r, m, nf = 64, 2, 10000
    a0, a1, a2, b = synthetic.reduced_model(r, m, seed=11)
    f = np.linspace(3e9, 5e9, nf)
    cb = np.array([orc.b_coefficient(t) for t in f])
    res = run_sweep(dv, f, a0, a1, a2, b, cb)
    x = res.x.cpu().numpy()
    assert not np.any(res.info.cpu().numpy())
    for i in range(0, nf, 397):
        a = orc.system_matrix(1.0, f[i], f[i] ** 2, a0, a1, a2)
        resid = np.linalg.norm(a @ x[i] - cb[i] * b) / (np.linalg.norm(a) * np.linalg.norm(x[i]))
        assert resid < 1e-14, (i, resid)
    # energy conservation of the lossless model: S is unitary
    s = res.gsm.cpu().numpy()
    unit = np.einsum("fij,fkj->fik", s, s.conj())
    assert np.abs(unit - np.eye(m)).max() < 1e-6


# ---------------------------------------------------------------------------------- real float64 twin (row N2)
def run_sweep_real(dv, f, a0, a1, a2, b, cb, want_x=True, want_gsm=True, variant=0):
    from scipy.constants import pi, epsilon_0
    dev = dv.require_cuda()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)  # noqa: E731
    ops = [None if (a is None or not np.any(a)) else dv.symmetrize(t(a)) for a in (a0, a1, a2)]
    res = dv.sweep(ops[0], ops[1], ops[2], t(b), t(np.ones_like(f)), t(f), t(f ** 2), t(cb), t(2 * pi * f * epsilon_0), want_x=want_x, want_gsm=want_gsm,
                   variant=variant)
    torch.cuda.synchronize()
    return res


def real_variants_for(r, m):
    from morfem_b200 import _ffi
    lib = _ffi.load()
    return [v for v in (3, 5) if lib.mf_sweep_f64_variant_supported(r, m, v)]


@pytest.mark.parametrize("name", REDUCED)
def test_real_sweep_matches_live_reference_fixture(dv, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    f = g["f"]
    cb = np.array([orc.b_coefficient(t) for t in f])
    r, m = g["b"].shape
    tol = np.maximum(1e-10, 20 * EPS * g["cond"])
    for variant in [0] + real_variants_for(r, m):
        res = run_sweep_real(dv, f, g["a0"], g["a1"], g["a2"], g["b"], cb, variant=variant)
        assert res.x.dtype == torch.float64                        # every size up to r = 512 has a real kernel (the reference's dtype)
        x = res.x.cpu().numpy()
        assert not np.any(res.info.cpu().numpy())
        assert np.all(per_point_rel(x, g["x"]) < tol) and np.all(per_point_rel(res.gsm.cpu().numpy(), g["gsm"]) < tol), variant


@pytest.mark.parametrize("r,m,nf", [(1, 1, 3), (7, 3, 33), (16, 16, 9), (31, 5, 40), (64, 2, 500), (100, 8, 12), (128, 4, 6),
                                    (129, 1, 5), (160, 4, 300), (200, 9, 7), (256, 4, 40), (300, 2, 3), (512, 8, 150)])
def test_real_sweep_equals_complex_sweep_and_oracle(dv, r, m, nf):
    from morfem_b200 import synthetic
    a0, a1, a2, b = synthetic.reduced_model(r, m, seed=100 + r)
    f = np.linspace(3e9, 5e9, nf)
    tb = orc.b_coefficient
    cb = np.array([tb(t) for t in f])
    x_ref = orc.reduced_sweep(f, a0, a1, a2, b, lambda t: 1.0, lambda t: t, lambda t: t ** 2, tb)
    s_ref = orc.scattering_sweep(f, x_ref, b)
    cond = np.array([np.linalg.cond(orc.system_matrix(1.0, t, t ** 2, a0, a1, a2)) for t in f])
    tol = np.maximum(1e-10, 20 * EPS * cond)
    for variant in real_variants_for(r, m):
        real = run_sweep_real(dv, f, a0, a1, a2, b, cb, variant=variant)
        cplx = run_sweep(dv, f, a0, a1, a2, b, cb, variant=variant)          # the complex128 instance of the same kernel
        assert real.x.dtype == torch.float64
        assert not np.any(real.info.cpu().numpy())
        # same algorithm and pivots (the two element types may run different bodies / CTA geometries of the left-looking kernel,
        # so the agreement is to rounding times the conditioning of the point, not bit for bit)
        same = np.maximum(1e-12, 20 * EPS * cond)
        assert np.all(per_point_rel(real.x.cpu().numpy(), cplx.x.cpu().numpy().real) < same), variant
        assert np.all(per_point_rel(real.gsm.cpu().numpy(), cplx.gsm.cpu().numpy()) < same), variant
        assert np.all(per_point_rel(real.x.cpu().numpy(), x_ref) < tol), variant
        assert np.all(per_point_rel(real.gsm.cpu().numpy(), s_ref) < tol), variant


def test_real_left_looking_sweep_pivots_and_reports_singular_points(dv):
    """The float64 left-looking kernel on a permutation-like matrix (an exchange in every column of every panel) and on an
    exactly singular one (first zero pivot deep inside the fifth panel)."""
    r, m = 150, 2
    rng = np.random.default_rng(8)
    a0 = np.fliplr(np.eye(r)) * 2.0 + 1e-3 * rng.standard_normal((r, r))
    a0 = (a0 + a0.T) / 2
    b = rng.standard_normal((r, m))
    f = np.array([3e9, 4e9, 5e9])
    zero = np.zeros((r, r))
    x_ref = orc.reduced_sweep(f, a0, zero, zero, b, lambda t: 1.0, lambda t: t, lambda t: t ** 2, lambda t: 1.0)
    res = run_sweep_real(dv, f, a0, None, None, b, np.ones(3), variant=5)
    assert np.all(per_point_rel(res.x.cpu().numpy(), x_ref) < 1e-11)
    sing = np.eye(r)
    sing[77, 77] = 0.0
    res = run_sweep_real(dv, f, sing, None, None, np.ones((r, m)), np.ones(3), variant=5, want_gsm=False)
    assert list(res.info.cpu().numpy()) == [78, 78, 78]


def test_real_sweep_reports_singular_points(dv):
    r, m = 20, 2
    a0 = np.eye(r)
    a0[11, 11] = 0.0
    res = run_sweep_real(dv, np.array([3e9, 4e9]), a0, None, None, np.ones((r, m)), np.ones(2), want_gsm=False)
    assert list(res.info.cpu().numpy()) == [12, 12]


def test_public_stage3_and_stage4_api_match_live_reference(dv):
    """The reference-facing calls themselves: ``solve_finite_element_method`` / ``solve_fem_point`` (implementation.py:189-194,
    :468-480) on a dense reduced ModelDefinition and ``generalized_scattering_matrix`` (test_helpers.py:9-14), against the
    committed outputs of the live reference; real input -> float64 result like the reference (:190)."""
    from morfem_b200 import implementation as impl, test_helpers as th
    g = np.load(os.path.join(GOLDEN, "reduced_r33_m3.npz"))
    f = g["f"]
    md = impl.ModelDefinition(f, g["a0"], g["a1"], g["a2"], g["b"], lambda t: 1.0, lambda t: t, lambda t: t ** 2, th.b_coefficient)
    x = impl.solve_finite_element_method(md)
    assert x.dtype == np.float64 and x.shape == g["x"].shape and x.flags.c_contiguous
    tol = np.maximum(1e-10, 20 * EPS * g["cond"])
    assert np.all(per_point_rel(x, g["x"]) < tol)
    i = f.size // 2
    xi = impl.solve_fem_point(f[i], md)
    assert xi.shape == g["x"][i].shape and np.linalg.norm(xi - g["x"][i]) / np.linalg.norm(g["x"][i]) < tol[i]
    s_i = th.generalized_scattering_matrix(f[i], g["x"][i], th.b_coefficient(f[i]) * g["b"])
    assert np.linalg.norm(s_i - g["gsm"][i]) / np.linalg.norm(g["gsm"][i]) < tol[i]
    # complex operands: complex result (documented deviation D2), checked against the oracle's restatement
    a0c = g["a0"] * (1.0 + 0.01j)
    mdc = impl.ModelDefinition(f[:5], a0c, g["a1"], g["a2"], g["b"], lambda t: 1.0, lambda t: t, lambda t: t ** 2, th.b_coefficient)
    xc = impl.solve_finite_element_method(mdc)
    xr = orc.reduced_sweep(f[:5], a0c, g["a1"], g["a2"], g["b"], lambda t: 1.0, lambda t: t, lambda t: t ** 2, orc.b_coefficient, complex_ok=True)
    assert np.iscomplexobj(xc) and np.all(per_point_rel(xc, xr) < 1e-8)


@pytest.mark.parametrize("r,m,real", [(208, 3, False), (256, 4, False), (224, 2, True), (200, 5, True)])
def test_left_looking_bodies_and_chunked_launches(dv, r, m, real, monkeypatch):
    """The left-looking kernel cuts a long batch into launches of a few points per CTA (``left_chunk``; MF_LEFT_CHUNK overrides) and
    has two bodies (MF_LEFT_VER: 2 plain, 3 look-ahead).  A batch that spans several launches with a ragged last one: the same body
    gives the same bits whatever the launch length (x, S and info land at the right points), both bodies and the library's own choice
    match the oracle, and a singular system at one point deep inside the batch is reported at that point only."""
    from scipy.constants import pi, epsilon_0
    from morfem_b200 import synthetic
    nf = 700                                           # 296 resident CTAs: launches of 296 + 296 + 108 points at one point per CTA
    a0, a1, a2, b = synthetic.reduced_model(r, m, seed=500 + r)
    f = np.linspace(3e9, 5e9, nf)
    tb = orc.b_coefficient
    cb = np.array([tb(t) for t in f])
    dev = dv.require_cuda()
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)  # noqa: E731
    conv = up if real else dv.to_device_c128
    ops = (dv.symmetrize(conv(a0)), None, dv.symmetrize(conv(a2)), conv(b))
    coef = (up(np.ones_like(f)), up(f), up(f ** 2), up(cb), up(2 * pi * f * epsilon_0))

    def go(ver=None, chunk=None, operands=ops, coefficients=coef, want_gsm=True):
        for key, val in (("MF_LEFT_VER", ver), ("MF_LEFT_CHUNK", chunk)):
            if val is None:
                monkeypatch.delenv(key, raising=False)
            else:
                monkeypatch.setenv(key, str(val))
        res = dv.sweep(*operands, *coefficients, want_x=True, want_gsm=want_gsm, variant=5)
        torch.cuda.synchronize()
        return res

    sample = np.array([0, 1, 147, 295, 296, 297, 591, 592, 593, 650, 698, 699])
    x_ref = orc.reduced_sweep(f[sample], a0, a1, a2, b, lambda t: 1.0, lambda t: t, lambda t: t ** 2, tb)
    s_ref = orc.scattering_sweep(f[sample], x_ref, b)
    cond = np.array([np.linalg.cond(orc.system_matrix(1.0, t, t ** 2, a0, a1, a2)) for t in f[sample]])
    tol = np.maximum(1e-10, 20 * EPS * cond)
    for ver in (2, 3):
        one = go(ver, 0)
        assert not np.any(one.info.cpu().numpy())
        for chunk in (1, 2):
            cut = go(ver, chunk)
            assert torch.equal(cut.x, one.x) and torch.equal(cut.gsm, one.gsm) and torch.equal(cut.info, one.info), (ver, chunk)
        assert np.all(per_point_rel(one.x.cpu().numpy()[sample].real, x_ref) < tol), ver
        assert np.all(per_point_rel(one.gsm.cpu().numpy()[sample], s_ref) < tol), ver
    own = go()                                         # the library's choice of body and launch length
    assert np.all(per_point_rel(own.x.cpu().numpy()[sample].real, x_ref) < tol)
    assert np.all(per_point_rel(own.gsm.cpu().numpy()[sample], s_ref) < tol)
    # A(t) = I - c2(t) E with E = e_k e_k^T and c2 = 1 at point 600 only: exactly singular there (LAPACK info k + 1), identity elsewhere
    k = r // 2 + 5
    e = np.zeros((r, r))
    e[k, k] = -1.0
    c2 = np.zeros(nf)
    c2[600] = 1.0
    sing_ops = (conv(np.eye(r)), None, conv(e), conv(np.ones((r, m))))
    sing_coef = (up(np.ones(nf)), up(np.zeros(nf)), up(c2), up(np.ones(nf)), up(np.ones(nf)))
    for ver, chunk in ((2, 1), (3, 2), (None, None)):
        res = go(ver, chunk, sing_ops, sing_coef, want_gsm=False)
        info = res.info.cpu().numpy()
        assert info[600] == k + 1 and not np.any(np.delete(info, 600)), (ver, chunk)
        assert np.all(res.x.cpu().numpy()[:600].real == 1.0)
