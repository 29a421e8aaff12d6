"""ctypes binding of libgmmvi_b200.so (C ABI in include/gmmvi_b200.h).

There is deliberately no fallback: if the shared library is missing the import of any op fails
loudly with the build command to run.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgmmvi_b200.so")

_lib = None

c_f = C.c_void_p      # device float*
c_i = C.c_void_p      # device int32*
c_vp = C.c_void_p

_SIGNATURES = {
    "gvi_version": (C.c_int, []),
    "gvi_last_error": (C.c_char_p, []),
    "gvi_prepare_full_workspace": (C.c_size_t, [C.c_int, C.c_int]),
    "gvi_prepare_full_f32": (C.c_int, [c_f, C.c_int, C.c_int, c_f, c_f, c_f, c_i, c_vp, C.c_size_t, c_vp]),
    "gvi_logdens_full_f32": (C.c_int, [c_f, C.c_int, C.c_int, c_f, c_f, c_f, C.c_int, c_f, c_vp]),
    "gvi_split_tf32_f32": (C.c_int, [c_f, C.c_longlong, c_f, c_f, c_vp]),
    "gvi_logdens_full_tc_supported": (C.c_int, [C.c_int]),
    "gvi_logdens_full_tc_f32": (C.c_int, [c_f, C.c_int, C.c_int, c_f, c_f, c_f, c_f, C.c_int, c_f, c_vp]),
    "gvi_logdens_full_h16_supported": (C.c_int, [C.c_int]),
    "gvi_h16_padded_dim": (C.c_int, [C.c_int]),
    "gvi_split_h16_f32": (C.c_int, [c_f, C.c_int, C.c_int, c_vp, c_vp, c_f, c_vp]),
    "gvi_group_absmax_f32": (C.c_int, [c_f, C.c_longlong, C.c_int, C.c_int, c_f, c_vp]),
    "gvi_logdens_full_h16_f32": (C.c_int, [c_f, c_f, C.c_int, C.c_int, c_f, c_f, c_vp, c_vp, c_f, c_f, C.c_int, c_f,
                                           c_vp]),
    "gvi_logdens_diag_f32": (C.c_int, [c_f, C.c_int, C.c_int, c_f, c_f, C.c_int, c_f, c_vp]),
    "gvi_mixture_lse_f32": (C.c_int, [c_f, c_f, C.c_int, C.c_int, c_f, c_vp]),
    "gvi_mixture_grad_full_workspace": (C.c_size_t, [C.c_int, C.c_int]),
    "gvi_mixture_grad_full_f32": (C.c_int, [c_f, C.c_int, C.c_int, c_f, c_f, c_f, c_f, c_f, C.c_int, c_f, c_vp,
                                            C.c_size_t, c_vp]),
    "gvi_mixture_grad_diag_f32": (C.c_int, [c_f, C.c_int, C.c_int, c_f, c_f, c_f, c_f, c_f, C.c_int, c_f, c_vp]),
    "gvi_importance_weights_f32": (C.c_int, [c_f, c_f, c_i, C.c_int, C.c_int, C.c_int, c_f, c_f, c_f, c_f, c_vp, c_vp]),
    "gvi_row_max_f32": (C.c_int, [c_f, c_f, C.c_int, C.c_int, c_f, c_vp]),
    "gvi_row_sumexp_f32": (C.c_int, [c_f, c_f, C.c_int, C.c_int, c_f, c_f, c_vp]),
    "gvi_importance_weights_ext_f32": (C.c_int, [c_f, c_f, C.c_int, C.c_int, c_f, c_f, c_f, c_f, c_f, c_f, c_vp, c_vp]),
    "gvi_stein_full_workspace": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "gvi_stein_full_f32": (C.c_int, [c_f, C.c_int, C.c_int, c_f, c_f, c_f, c_vp, c_f, C.c_int, C.c_int, c_f, c_f,
                                     c_vp, C.c_size_t, c_vp]),
    "gvi_stein_stats_full_workspace": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "gvi_stein_stats_full_f32": (C.c_int, [c_f, C.c_int, C.c_int, c_f, c_f, c_vp, c_f, C.c_int, c_f, c_f, c_vp, C.c_size_t,
                                           c_vp]),
    "gvi_stein_finalize_full_workspace": (C.c_size_t, [C.c_int, C.c_int]),
    "gvi_stein_finalize_full_f32": (C.c_int, [c_f, c_f, C.c_int, C.c_int, C.c_int, c_f, c_vp, C.c_size_t, c_vp]),
    "gvi_stein_diag_workspace": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "gvi_stein_diag_f32": (C.c_int, [c_f, C.c_int, C.c_int, c_f, c_f, c_f, c_f, C.c_int, c_f, c_f, c_vp, C.c_size_t, c_vp]),
    "gvi_more_workspace": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "gvi_more_fit_f32": (C.c_int, [c_f, C.c_int, C.c_int, c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, c_f, c_f, c_i,
                                   c_vp, C.c_size_t, c_vp]),
    "gvi_update_full_workspace": (C.c_size_t, [C.c_int, C.c_int]),
    "gvi_mixture_grad_full_h16_supported": (C.c_int, [C.c_int]),
    "gvi_split_h16_full_f32": (C.c_int, [c_f, C.c_int, C.c_int, c_vp, c_vp, c_f, c_vp]),
    "gvi_mixture_grad_full_h16_f32": (C.c_int, [c_f, c_f, C.c_int, C.c_int, c_f, c_f, c_vp, c_vp, c_f, c_f, c_f, c_f,
                                                C.c_int, c_f, c_vp, C.c_size_t, c_vp]),
    "gvi_gauss_kernel_sum_partials": (C.c_size_t, [C.c_int, C.c_int]),
    "gvi_gauss_kernel_sum_f32": (C.c_int, [c_f, C.c_int, c_f, C.c_int, C.c_int, c_f, c_vp, c_vp]),
    "gvi_cholesky_workspace": (C.c_size_t, [C.c_int, C.c_int]),
    "gvi_cholesky_f32": (C.c_int, [c_f, C.c_int, C.c_int, c_f, c_i, c_vp, C.c_size_t, c_vp]),
    "gvi_planar_robot_f32": (C.c_int, [c_f, C.c_int, C.c_int, c_f, c_f, c_f, C.c_int, C.c_float, c_f, c_f, c_vp]),
    "gvi_tridiag_f32": (C.c_int, [c_f, c_f, C.c_int, C.c_int, c_f, c_f, c_f, c_vp]),
    "gvi_update_full_f32": (C.c_int, [C.c_int, c_f, c_f, c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_float,
                                      c_f, c_f, c_i, c_f, c_f, c_i, c_vp, C.c_size_t, c_vp]),
    "gvi_update_full_general_workspace": (C.c_size_t, [C.c_int, C.c_int]),
    "gvi_update_full_general_f32": (C.c_int, [C.c_int, c_f, c_f, c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, c_f, c_f, c_i,
                                              c_vp, C.c_size_t, c_vp]),
    "gvi_update_diag_f32": (C.c_int, [C.c_int, c_f, c_f, c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_float,
                                      c_f, c_f, c_i, c_f, c_f, c_vp]),
    "gvi_weight_update_f32": (C.c_int, [C.c_int, c_f, c_f, C.c_int, c_f, C.c_float, c_f, c_f, c_vp]),
    "gvi_fill_normal_f32": (C.c_int, [c_f, C.c_longlong, C.c_int, C.c_ulonglong, C.c_ulonglong, C.c_longlong, c_vp]),
    "gvi_fill_normal_dev_f32": (C.c_int, [c_f, C.c_longlong, C.c_int, C.c_ulonglong, c_vp, C.c_ulonglong, C.c_longlong,
                                          c_vp]),
    "gvi_sample_f32": (C.c_int, [C.c_int, c_f, c_i, c_f, c_f, C.c_int, C.c_int, C.c_int, c_f, c_i, c_vp]),
    "gvi_tc_bgemm_supported": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "gvi_tc_bgemm_workspace": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "gvi_tc_bgemm_f32": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_f, C.c_int,
                                   C.c_longlong, c_f, C.c_int, C.c_longlong, c_f, C.c_int, C.c_longlong, c_vp,
                                   C.c_size_t, c_vp]),
    "gvi_tc_bgemm_ex_f32": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_f, C.c_int,
                                      C.c_longlong, c_f, C.c_int, C.c_longlong, c_f, C.c_int, C.c_longlong, C.c_float,
                                      C.c_int, C.c_int, c_vp, C.c_size_t, c_vp]),
    "gvi_more_tensor_cores": (C.c_int, []),
    "gvi_tc_bgemm_h16_workspace": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "gvi_tc_bgemm_h16_f32": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_f, c_f, c_f, C.c_float, C.c_int,
                                       C.c_int, c_vp, C.c_size_t, c_vp]),
    "gvi_bgemm_f32": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_f, C.c_int,
                                C.c_longlong, c_f, C.c_int, C.c_longlong, c_f, C.c_int, C.c_longlong, c_vp]),
}


class GmmviLibraryError(RuntimeError):
    pass


def exported_symbols():
    """Names include/gmmvi_b200.h declares (used by the symbol test)."""
    return sorted(_SIGNATURES)


def register(name, restype, argtypes):
    """Register an additional entry point (used by optional kernel families, e.g. the tcgen05 path)."""
    _SIGNATURES[name] = (restype, argtypes)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype, fn.argtypes = restype, argtypes


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GmmviLibraryError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built.  Run "
                f"`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C gmmvi_b200/csrc`).  "
                "gmmvi_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError as e:  # pragma: no cover
                raise GmmviLibraryError(f"{LIB_PATH} does not export {name}; rebuild it") from e
            fn.restype, fn.argtypes = restype, argtypes
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().gvi_last_error().decode("utf-8", "replace")
        raise GmmviLibraryError(f"{what} failed with code {rc}: {msg}")
