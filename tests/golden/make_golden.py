"""Generate golden input/output vectors by running the LIVE reference (``/root/reference``, unmodified).

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

The reference imports matplotlib at module top (``implementation.py:10``) and matplotlib is not installed,
so two empty stub modules are registered first; plotting is behind ``PLOT_GREEDY_ITERATIONS = False``.
Outputs are written next to this script as compressed ``.npz`` files and committed.

Cases
  cfg1_rom3411      whole ROM driver (``test_helpers.finite_element_method_model_order_reduction_gsm``) on an
                    N=3411 surrogate paired with the shipped ``data/WP.npy``; also records the last
                    ``np.linalg.svd`` input/output inside the greedy loop (a real stage-1 pair).
  stages_n600       ``implementation.morfem`` with ``projection_base`` pinned to the SVD of a seeded snapshot
                    block, so lines 178-186 (projection + reduced sweep) run verbatim; then the GSM loop.
  equidist_n600     ``USE_EQUALLY_DISTRIBUTED = True`` path (``implementation.py:197-214``) end to end.
  reduced_rXX_mY    ``solve_finite_element_method`` + ``generalized_scattering_matrix`` on seeded reduced models.
  opm_n600          the greedy search with ``USE_OPM = True`` (incremental Gram-Schmidt + blockwise growth of the estimator matrices).
  estimator_n600    ``implementation.error_estimator`` (:348-452) itself, called on the orthonormalised bases a greedy run
                    passes through (r = 4, 6, 8, 10; 2 ports) and on a 3-port model: the per-point residual estimate the
                    greedy search takes its arg-max of, plus the size of the largest term of the 16-term sum (the estimate
                    is a difference of terms of that size, so it is only defined to eps times it).
"""
from __future__ import annotations

import io
import contextlib
import os
import sys
import types
import warnings

import numpy as np
from scipy.sparse import csc_array

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, "/root/reference")
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    import implementation as ref_impl   # noqa: E402
    import test_helpers as ref_help     # noqa: E402

from morfem_b200 import synthetic  # noqa: E402


def quiet(fn, *args, **kwargs):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*args, **kwargs)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB  " + ", ".join(f"{k}{tuple(np.shape(v))}" for k, v in arrays.items()))


def csc_parts(prefix, m):
    m = csc_array(m)
    return {prefix + "_data": m.data, prefix + "_indices": m.indices, prefix + "_indptr": m.indptr,
            prefix + "_shape": np.array(m.shape)}


def case_cfg1():
    ct, tt = synthetic.waveguide_operators(9, 1, 379)
    wp_shipped = np.load("/root/reference/data/WP.npy")
    wp_regen = synthetic.shipped_port_matrix().toarray()
    assert np.abs(wp_shipped - wp_regen).max() < 2e-7, np.abs(wp_shipped - wp_regen).max()
    in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, csc_array(wp_shipped))
    f = np.linspace(3e9, 5e9, 100)  # main.py:18

    svd_calls = []
    real_svd = np.linalg.svd

    def recording_svd(a, *args, **kwargs):
        out = real_svd(a, *args, **kwargs)
        svd_calls.append((np.array(a), np.array(out[0])))
        return out

    np.linalg.svd = recording_svd
    try:
        x, q, a0_r, a1_r, a2_r, b_r = quiet(ref_impl.morfem, f, in_c, csc_array(in_c.shape), in_gamma, in_b,
                                            t_b=lambda t: ref_help.b_coefficient(t))
    finally:
        np.linalg.svd = real_svd
    gsm_rom = quiet(ref_help.finite_element_method_model_order_reduction_gsm, f, 2, in_c, in_gamma, in_b)
    gsm_full = quiet(ref_help.finite_element_method_gsm, f, 2, in_c, in_gamma, in_b)
    gsm_from_x = np.stack([ref_help.generalized_scattering_matrix(f[i], x[i], ref_help.b_coefficient(f[i]) * b_r)
                           for i in range(f.size)])
    assert np.allclose(gsm_from_x, gsm_rom, rtol=0, atol=1e-9)
    last_s, last_u = svd_calls[-1]
    save("cfg1_rom3411", grid=np.array([9, 1, 379]), f=f, wp_shipped_nz_rows=np.nonzero(wp_shipped)[0],
         wp_shipped_nz_cols=np.nonzero(wp_shipped)[1], wp_shipped_nz_vals=wp_shipped[np.nonzero(wp_shipped)],
         x=x, q=q, a0_r=a0_r, a1_r=a1_r, a2_r=a2_r, b_r=b_r, gsm_rom=gsm_from_x, gsm_full=gsm_full,
         svd_in=last_s, svd_out=last_u)


def pinned_basis_morfem(snapshots, f, in_c, in_gamma, in_b):
    """Run implementation.morfem with the greedy basis search replaced by the SVD of ``snapshots`` so that
    lines 178-186 execute verbatim on a known basis."""
    saved = ref_impl.projection_base
    ref_impl.projection_base = lambda md: np.linalg.svd(snapshots, full_matrices=False)[0]
    try:
        return quiet(ref_impl.morfem, f, in_c, csc_array(in_c.shape), in_gamma, in_b,
                     t_b=lambda t: ref_help.b_coefficient(t))
    finally:
        ref_impl.projection_base = saved


def case_stages():
    ct, tt = synthetic.waveguide_operators(5, 4, 30)
    n = ct.shape[0]
    wp = synthetic.port_matrix(n, 2, 19)
    in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, wp)
    f = np.linspace(3e9, 5e9, 40)
    # true snapshots at 8 frequencies (full-order solves, implementation.py:475) -> 16 columns
    md = ref_impl.ModelDefinition(f, in_c, csc_array(in_c.shape), in_gamma, in_b, lambda t: 1, lambda t: t,
                                  lambda t: t ** 2, lambda t: ref_help.b_coefficient(t))
    idx = np.linspace(0, f.size - 1, 8, dtype=int)
    snaps = np.hstack([ref_impl.solve_fem_point(f[i], md) for i in idx])
    x, q, a0_r, a1_r, a2_r, b_r = pinned_basis_morfem(snaps, f, in_c, in_gamma, in_b)
    gsm = np.stack([ref_help.generalized_scattering_matrix(f[i], x[i], ref_help.b_coefficient(f[i]) * b_r)
                    for i in range(f.size)])
    save("stages_n600", grid=np.array([5, 4, 30]), ports=np.array(2), face=np.array(19), f=f, snapshots=snaps,
         q=q, a0_r=a0_r, a1_r=a1_r, a2_r=a2_r, b_r=b_r, x=x, gsm=gsm)


def case_equidist():
    ct, tt = synthetic.waveguide_operators(5, 4, 30)
    n = ct.shape[0]
    wp = synthetic.port_matrix(n, 2, 19)
    in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, wp)
    f = np.linspace(3e9, 5e9, 200)
    ref_impl.USE_EQUALLY_DISTRIBUTED = True
    try:
        x, q, a0_r, a1_r, a2_r, b_r = quiet(ref_impl.morfem, f, in_c, csc_array(in_c.shape), in_gamma, in_b,
                                            t_b=lambda t: ref_help.b_coefficient(t))
    finally:
        ref_impl.USE_EQUALLY_DISTRIBUTED = False
    gsm = np.stack([ref_help.generalized_scattering_matrix(f[i], x[i], ref_help.b_coefficient(f[i]) * b_r)
                    for i in range(f.size)])
    save("equidist_n600", grid=np.array([5, 4, 30]), ports=np.array(2), face=np.array(19), f=f,
         q=q, a0_r=a0_r, a1_r=a1_r, a2_r=a2_r, b_r=b_r, x=x, gsm=gsm)


def case_reduced(r, m, npts, seed):
    a0, a1, a2, b = synthetic.reduced_model(r, m, seed=seed)
    f = np.linspace(3e9, 5e9, npts)
    md = ref_impl.ModelDefinition(f, a0, a1, a2, b, lambda t: 1., lambda t: t, lambda t: t ** 2,
                                  lambda t: ref_help.b_coefficient(t))
    x = ref_impl.solve_finite_element_method(md)
    gsm = np.stack([ref_help.generalized_scattering_matrix(f[i], x[i], ref_help.b_coefficient(f[i]) * b)
                    for i in range(f.size)])
    cond = np.array([np.linalg.cond(ref_impl.system_matrix(t, md)) for t in f])
    save(f"reduced_r{r}_m{m}", r=np.array(r), m=np.array(m), seed=np.array(seed), f=f, a0=a0, a1=a1, a2=a2, b=b,
         x=x, gsm=gsm, cond=cond)


def case_estimator():
    out = {}
    for tag, ports in (("p2", 2), ("p3", 3)):
        ct, tt = synthetic.waveguide_operators(5, 4, 30)
        n = ct.shape[0]
        wp = synthetic.port_matrix(n, ports, 19)
        in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, wp)
        f = np.linspace(3e9, 5e9, 60)
        md = ref_impl.ModelDefinition(f, in_c, csc_array(in_c.shape), in_gamma, in_b, lambda t: 1, lambda t: t,
                                      lambda t: t ** 2, lambda t: ref_help.b_coefficient(t))
        # the bases of a greedy run (implementation.py:222-226, :297-298): end points first, then the arg-max points
        q = np.linalg.svd(np.hstack((ref_impl.solve_fem_point(f[0], md), ref_impl.solve_fem_point(f[-1], md))),
                          full_matrices=False)[0]
        for it in range(4 if ports == 2 else 2):
            err = quiet(ref_impl.error_estimator, md, q, ref_impl.OfflinePhaseMatrices(), ref_impl.TimeStatistics())
            bh_b = (in_b.T @ in_b).toarray()
            scale = np.array([ref_help.b_coefficient(t) ** 2 for t in f]) * np.linalg.norm(bh_b)
            out[f"{tag}_q{it}"] = q
            out[f"{tag}_err{it}"] = err
            out[f"{tag}_scale{it}"] = scale
            q_new = ref_impl.solve_fem_point(f[int(err.argmax())], md)
            q = np.linalg.svd(np.hstack((q, q_new)), full_matrices=False)[0]
        out[f"{tag}_f"] = f
    save("estimator_n600", grid=np.array([5, 4, 30]), face=np.array(19), **out)


def case_opm():
    """The greedy search with USE_OPM = True (implementation.py:16, :230-295): new vectors Gram-Schmidt-orthonormalised against the base,
    the 16 estimator matrices grown blockwise.  Records the estimator curve of every iteration, the final basis size and the ROM S-parameters."""
    ct, tt = synthetic.waveguide_operators(5, 4, 30)
    n = ct.shape[0]
    wp = synthetic.port_matrix(n, 2, 19)
    in_c, in_gamma, in_b = synthetic.driver_scaled(ct, tt, wp)
    f = np.linspace(3e9, 5e9, 60)
    md = ref_impl.ModelDefinition(f, in_c, csc_array(in_c.shape), in_gamma, in_b, lambda t: 1, lambda t: t,
                                  lambda t: t ** 2, lambda t: ref_help.b_coefficient(t))
    errors = []
    real_new = ref_impl.new_solution_for_projection_base

    def recording(md_, q_, opm_, ts_):
        q_new, err = real_new(md_, q_, opm_, ts_)
        errors.append(np.array(err))
        return q_new, err

    ref_impl.USE_OPM = True
    ref_impl.new_solution_for_projection_base = recording
    try:
        x, q, a0_r, a1_r, a2_r, b_r = quiet(ref_impl.morfem, f, in_c, csc_array(in_c.shape), in_gamma, in_b,
                                            t_b=lambda t: ref_help.b_coefficient(t))
    finally:
        ref_impl.USE_OPM = False
        ref_impl.new_solution_for_projection_base = real_new
    gsm = np.stack([ref_help.generalized_scattering_matrix(f[i], x[i], ref_help.b_coefficient(f[i]) * b_r) for i in range(f.size)])
    gsm_full = quiet(ref_help.finite_element_method_gsm, f, 2, in_c, in_gamma, in_b)
    save("opm_n600", grid=np.array([5, 4, 30]), ports=np.array(2), face=np.array(19), f=f, basis_size=np.array(q.shape[1]),
         errors=np.stack(errors), gsm_rom=gsm, gsm_full=gsm_full,
         scale=np.array([ref_help.b_coefficient(t) ** 2 for t in f]) * np.linalg.norm((in_b.T @ in_b).toarray()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "estimator":
        case_estimator()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "opm":
        case_opm()
        sys.exit(0)
    case_cfg1()
    case_stages()
    case_equidist()
    for r, m, npts, seed in [(8, 2, 64, 1), (24, 4, 48, 2), (33, 3, 24, 6), (64, 2, 64, 3), (96, 3, 16, 4), (160, 4, 8, 5),
                             (256, 4, 4, 7)]:
        case_reduced(r, m, npts, seed)
    case_estimator()
    case_opm()
