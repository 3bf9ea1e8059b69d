"""Full-order sparse solves: ``morfem_b200.full_order`` against the oracle's restatement of the reference loop
(``splu`` of the symmetrised system matrix at every point, implementation.py:189-194, :468-480, :526-533) -- host logic, no GPU.
The prepared solver keeps SuperLU's own column ordering, so the comparison is bit for bit."""
import numpy as np
import pytest
import scipy.sparse as sp

from morfem_b200 import full_order, implementation as impl, synthetic
from morfem_b200.test_helpers import b_coefficient
from oracle import reference_path as rp


def _model(nx=4, ny=3, nz=14, ports=2, with_a1=False, unsymmetric=False, complex_valued=False, points=7):
    ct, tt = synthetic.waveguide_operators(nx, ny, nz)
    n = ct.shape[0]
    c, g, b = synthetic.driver_scaled(ct, tt, synthetic.port_matrix(n, ports, nx * ny))
    a1 = sp.csc_array(c.shape)
    rng = np.random.default_rng(5)
    if with_a1:                      # a damping-like term whose pattern is NOT contained in the other two
        rows, cols = rng.integers(0, n, 3 * n), rng.integers(0, n, 3 * n)
        a1 = sp.csc_array((rng.standard_normal(3 * n) * 1e-12, (rows, cols)), shape=c.shape)
    if unsymmetric:                  # the symmetrisation of implementation.py:528 has something to do
        c = sp.csc_array(c + sp.triu(c, 1) * 0.25)
    if complex_valued:
        g = sp.csc_array(g * (1.0 - 0.02j))
    f = synthetic.frequency_points(points)
    md = impl.ModelDefinition(f, sp.csc_array(c), a1, sp.csc_array(g), sp.csc_array(b),
                              lambda t: 1., lambda t: t, lambda t: t ** 2, b_coefficient)
    return md


def _oracle(md, dtype=float):
    if dtype is float:
        return rp.full_order_sweep(md.domain, md.a0, md.a1, md.a2, md.b, md.t_a0, md.t_a1, md.t_a2, md.t_b)
    from scipy.sparse.linalg import splu           # the oracle allocates a float64 result like the reference (:190); complex case restated here
    out = []
    for t in md.domain:
        a = md.t_a0(t) * md.a0 + md.t_a1(t) * md.a1 + md.t_a2(t) * md.a2
        out.append(splu(sp.csc_matrix((a + a.T) / 2)).solve(np.asarray((md.t_b(t) * md.b).todense()).astype(complex)))
    return np.array(out)


@pytest.mark.parametrize("kw", [dict(), dict(with_a1=True), dict(unsymmetric=True), dict(ports=4, with_a1=True, unsymmetric=True)])
@pytest.mark.parametrize("threads", [1, 3])
def test_solve_many_is_the_reference_loop(kw, threads):
    md = _model(**kw)
    ref = _oracle(md)
    got = full_order.FullOrderSolver(md).solve_many(md.domain, threads=threads)
    assert got.shape == ref.shape and got.dtype == np.float64
    scale = np.abs(ref).max(axis=(1, 2), keepdims=True)
    assert (np.abs(got - ref) / scale).max() <= 1e-12            # assembled on the union pattern: same values up to summation order
    if not kw:                                                    # symmetric operators, two of them: the same roundings, bit for bit
        assert np.array_equal(got, ref)


def test_ordering_reuse_matches_fresh_colamd():
    md = _model(with_a1=True)
    fresh = full_order.FullOrderSolver(md, reuse_ordering=False).solve_many(md.domain, threads=1)
    s = full_order.FullOrderSolver(md)
    reused = s.solve_many(md.domain, threads=1)
    assert s._perm_c is not None and sorted(s._perm_c.tolist()) == list(range(s.n))
    assert np.array_equal(fresh, reused)                          # same ordering -> the same elimination, bit for bit
    a = s.matrix(md.domain[2])                                    # the assembled matrix is the reference's system_matrix
    assert abs(a - sp.csc_matrix(impl.system_matrix(md.domain[2], md))).max() <= 1e-15 * abs(a).max()


def test_complex_model():
    md = _model(complex_valued=True, points=4)
    ref = _oracle(md, dtype=complex)
    got = full_order.FullOrderSolver(md).solve_many(md.domain, threads=2)
    assert np.iscomplexobj(got)
    assert (np.abs(got - ref) / np.abs(ref).max()).max() <= 1e-12


def test_public_entry_points_use_the_prepared_solver():
    md = _model(points=5)
    ref = _oracle(md)
    x = impl.solve_finite_element_method(md)                       # sparse branch of implementation.py:189-194
    assert np.array_equal(x, ref)
    one = impl.solve_fem_point(md.domain[3], md)                   # implementation.py:468-480
    assert np.array_equal(np.asarray(one), ref[3])
    assert full_order.solver_for(md) is full_order.solver_for(md)
    md.a0.data[0] *= 2.0                                           # edited in place: the prepared values must not be reused
    again = impl.solve_fem_point(md.domain[3], md)
    assert np.array_equal(np.asarray(again), _oracle(md)[3])
    assert not np.array_equal(again, one)


def test_empty_domain_and_single_point():
    md = _model(points=3)
    ref0 = _oracle(md)[0]
    s = full_order.FullOrderSolver(md)
    assert s.solve_many(np.zeros(0)).shape == (0, s.n, 2)
    assert np.array_equal(s.solve_many(md.domain[:1])[0], ref0)
    md.a1 = None                                                   # "no damping term" spelled as None instead of an empty matrix
    assert np.array_equal(full_order.solver_for(md).solve(md.domain[0]), ref0)
