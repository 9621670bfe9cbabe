"""ctypes binding of libdsea.so — the C ABI declared in include/dsea.h.

There is deliberately no fallback: if the CUDA library is missing or fails to load, importing the
compute modules raises.  Build it with `python -m dominantsparseeigenad_b200._build`.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdsea.so")

c_double_p = C.c_void_p      # device pointers travel as integers
c_stream = C.c_void_p

# name -> (restype, argtypes); mirrors include/dsea.h one to one
PROTOTYPES = {
    "dsea_last_error": (C.c_char_p, []),
    "dsea_version": (C.c_int, []),
    "dsea_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "dsea_ctx_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "dsea_ctx_destroy": (C.c_int, [C.c_void_p]),
    "dsea_ctx_rank": (C.c_int, [C.c_void_p]),
    "dsea_ctx_p2p": (C.c_int, [C.c_void_p]),
    "dsea_ctx_world": (C.c_int, [C.c_void_p]),
    "dsea_launch_count": (C.c_int64, [C.c_void_p]),
    "dsea_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "dsea_profile_collect": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_int64)]),
    "dsea_ctx_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "dsea_op_tfim": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "dsea_op_csr": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.POINTER(C.c_void_p)]),
    "dsea_op_dense": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.POINTER(C.c_void_p)]),
    "dsea_op_destroy": (C.c_int, [C.c_void_p]),
    "dsea_op_local_dim": (C.c_int64, [C.c_void_p]),
    "dsea_op_work_doubles": (C.c_int64, [C.c_void_p]),
    "dsea_col_stride": (C.c_int64, [C.c_int64]),
    "dsea_tfim_flip_index": (C.c_int64, [C.c_int, C.c_int64, C.c_int]),
    "dsea_tfim_diag": (C.c_double, [C.c_int, C.c_int64]),
    "dsea_tfim_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "dsea_matvec": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                              c_double_p, c_stream]),
    "dsea_tfim_dHdg": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, c_double_p, c_double_p, c_stream]),
    "dsea_adjoint": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_stream]),
    "dsea_lanczos_work_doubles": (C.c_int64, [C.c_void_p]),
    "dsea_lanczos_basis_doubles": (C.c_int64, [C.c_void_p, C.c_int]),
    "dsea_lanczos": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int, C.c_int, c_double_p, c_double_p,
                               c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                               C.POINTER(C.c_int64), c_stream]),
    "dsea_lanczos_start": (C.c_int, [C.c_void_p, C.c_int64, c_double_p, c_stream]),
    "dsea_lanczos_step": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p,
                                    c_double_p, c_stream]),
    "dsea_lanczos_ritz": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p,
                                    c_double_p, c_double_p, c_double_p, C.POINTER(C.c_int64), c_stream]),
    "dsea_arnoldi_start": (C.c_int, [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_stream]),
    "dsea_arnoldi_step": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p,
                                    c_stream]),
    "dsea_combine": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p,
                               c_stream]),
    "dsea_cg_work_doubles": (C.c_int64, [C.c_void_p]),
    "dsea_cg": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                          C.c_double, C.c_int64, C.POINTER(C.c_int64), c_stream]),
    "dsea_cg_init": (C.c_int, [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p,
                               C.POINTER(C.c_double), c_stream]),
    "dsea_cg_update": (C.c_int, [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p,
                                 C.POINTER(C.c_double), c_stream]),
    "dsea_dot": (C.c_int, [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_double_p, c_stream]),
    "dsea_project": (C.c_int, [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_double_p, c_stream]),
    "dsea_axpby": (C.c_int, [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, c_stream]),
    "dsea_outer": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, c_double_p, c_double_p, c_double_p, c_stream]),
    "dsea_randn": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, c_double_p, c_stream]),
}

DSEA_MIN, DSEA_MAX, DSEA_BOTH = 0, 1, 2
DSEA_ERR_NOCONV = -4

_lib = None


class DseaError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Loads libdsea.so (once) and attaches the prototypes.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DseaError(
            f"{LIB_PATH} not found: the CUDA library has not been built "
            "(run `python -m dominantsparseeigenad_b200._build`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class ConvergenceWarning(RuntimeWarning):
    """CG reached its iteration cap (or a NaN residual) before |r| < eps.  The reference returns the last
    iterate silently in that case (CG.py:32); here the iterate is returned as well, with this warning."""


def check(status: int, allow_noconv: bool = False) -> None:
    if status == 0:
        return
    msg = load().dsea_last_error()
    text = msg.decode() if msg else "?"
    if status == DSEA_ERR_NOCONV and allow_noconv:
        import warnings
        warnings.warn(text, ConvergenceWarning, stacklevel=3)
        return
    raise DseaError(f"libdsea error {status}: {text}")
