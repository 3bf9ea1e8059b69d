"""ctypes binding of ``libmorfem_b200.so`` (the C ABI declared in ``include/morfem_b200.h``).

The library is built in-tree by ``morfem_b200/csrc/Makefile`` (``__graft_entry__.build()``).  There is no
CPU fallback: if the shared object is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmorfem_b200.so")


class MorfemB200Error(RuntimeError):
    """A C-ABI call returned a non-zero status."""


# name -> (restype, argtypes); mirrors include/morfem_b200.h one to one
SIGNATURES = {
    "mf_version": (c_int, []),
    "mf_last_error": (c_char_p, []),
    "mf_launch_count": (c_int64, []),
    "mf_gemm_tn_ws_bytes": (c_size_t, [c_int, c_int, c_int64]),
    "mf_gemm_tn_c128": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int64, c_int, c_void_p, c_int64,
                                c_void_p, c_size_t, c_void_p]),
    "mf_gemm_nn_c128": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p]),
    "mf_trmm_nn_c128": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "mf_trmm_nn_f64": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "mf_equilibrate_c128": (c_int, [c_void_p, c_int64, c_int, c_double, c_void_p, c_void_p, c_void_p]),
    "mf_potrf_upper_c128": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "mf_trtri_upper_c128": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p]),
    "mf_chol_inv_ws_bytes": (c_size_t, [c_int]),
    "mf_chol_inv_upper_c128": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mf_scale_cols_c128": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_void_p]),
    "mf_scale_rows_c128": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_void_p]),
    "mf_jacobi_svd_ws_bytes": (c_size_t, [c_int]),
    "mf_jacobi_svd_c128": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p, c_int, c_double, c_void_p,
                                   c_void_p, c_size_t, c_void_p]),
    "mf_spmm_csr_c128": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_int64, c_int, c_void_p, c_int64,
                                 c_void_p]),
    "mf_spmm_csr2_c128": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_int64, c_int, c_void_p, c_int64,
                                  c_void_p, c_int64, c_void_p]),
    "mf_spmm_csr2_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_int64,
                                 c_void_p, c_int64, c_void_p]),
    "mf_spmm_group_size": (c_int, [c_int]),
    "mf_spmm_group_size_f64": (c_int, [c_int]),
    "mf_spmm_group_count": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "mf_spmm_group_fill": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mf_spmm_grouped_c128": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64,
                                     c_void_p]),
    "mf_spmm_window_rows_per_block": (c_int, []),
    "mf_spmm_window_max_nnz_per_row": (c_int, []),
    "mf_spmm_window_max_rows": (c_int, []),
    "mf_spmm_window_c128": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64,
                                    c_void_p]),
    "mf_spmm_window_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64,
                                   c_void_p]),
    "mf_project_rhs_c128": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int64, c_int, c_int64, c_int64,
                                    c_int, c_void_p, c_int64, c_void_p]),
    "mf_gemm_tn_f64_ws_bytes": (c_size_t, [c_int, c_int, c_int64]),
    "mf_gemm_tn_f64": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int64, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "mf_gemm_nn_f64": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p]),
    "mf_spmm_csr_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p]),
    "mf_spmm_grouped_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p]),
    "mf_sweep_f64_supported": (c_int, [c_int, c_int]),
    "mf_sweep_f64_variant_supported": (c_int, [c_int, c_int, c_int]),
    "mf_sweep_f64_ws_bytes": (c_size_t, [c_int, c_int, c_int64, c_int]),
    "mf_sweep_lu_gsm_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                    c_int, c_void_p, c_size_t, c_void_p]),
    "mf_jacobi_svd_f64_supported": (c_int, [c_int]),
    "mf_jacobi_svd_f64": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p, c_int, c_double, c_void_p, c_void_p]),
    "mf_project_rhs_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int, c_int64, c_int64, c_void_p, c_int64,
                                   c_void_p]),
    "mf_symmetrize_c128": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p]),
    "mf_sweep_ws_bytes": (c_size_t, [c_int, c_int, c_int64, c_int]),
    "mf_sweep_variant_supported": (c_int, [c_int, c_int, c_int]),
    "mf_sweep_lu_gsm_c128": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                     c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "mf_gsm_c128": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "mf_peak_dmma_tflops": (c_int, [c_int, c_void_p, c_void_p]),
    "mf_estimator_c128": (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once) and attach argument types.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MorfemB200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C morfem_b200/csrc`).  morfem_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library ever diverge
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().mf_last_error()
        raise MorfemB200Error(f"{what or 'morfem_b200 call'} failed with status {status}: "
                              f"{msg.decode(errors='replace') if msg else ''}")


def launch_count() -> int:
    return int(load().mf_launch_count())
