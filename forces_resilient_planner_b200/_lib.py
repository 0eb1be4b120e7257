"""ctypes view of libnmpc_b200.so (the C ABI declared in include/nmpc_b200.h)."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnmpc_b200.so")

_vp, _ip, _i = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int


class NmpcOpts(ctypes.Structure):
    """struct nmpc_opts (include/nmpc_b200.h)."""
    _fields_ = [("mu0", ctypes.c_double), ("sigma", ctypes.c_double), ("mu_floor", ctypes.c_double),
                ("tol_stat", ctypes.c_double), ("tol_eq", ctypes.c_double),
                ("tol_ineq", ctypes.c_double), ("tol_comp", ctypes.c_double),
                ("kappa_push", ctypes.c_double), ("s_floor", ctypes.c_double),
                ("maxit", ctypes.c_int), ("max_bt", ctypes.c_int), ("pc", ctypes.c_int), ("mixed", ctypes.c_int)]


class EllipsoidConsts(ctypes.Structure):
    """struct nmpc_ellipsoid_consts (include/nmpc_b200.h)."""
    _fields_ = [(n, ctypes.c_double) for n in ("mass", "drag", "ego_r", "ego_h", "ext_noise_bound", "epsilon", "Ts")]


# every symbol include/*.h declares (tests/test_abi.py checks the library exports each one)
EXPORTS = [
    "nmpc_default_opts", "nmpc_last_error", "nmpc_version", "nmpc_supported_horizon",
    "nmpc_smem_bytes", "nmpc_smem_bytes_pc", "nmpc_solve_batch_f64", "nmpc_solve_batch_f32",
    "nmpc_solve_batch_ex_f64", "nmpc_solve_batch_ordered_f64", "nmpc_solve_batch_mixed_f64", "nmpc_solve_batch_lowlatency_f64",
    "nmpc_solve_batch_host_f64", "nmpc_solve_batch_host_f32", "nmpc_solve_batch_host_mixed_f64", "nmpc_model_eval_host_f64",
    "nmpc_riccati_factor_f64", "nmpc_riccati_factor_f32",
    "nmpc_kkt_backsolve_f64", "nmpc_kkt_backsolve_f32", "nmpc_backsolve_factor_words",
    "nmpc_backsolve_algorithmic_bytes",
    "nmpc_pack_params_f64", "nmpc_shift_warm_start_f64", "nmpc_wrap_yaw_f64", "nmpc_sample_reference_f64", "nmpc_default_ellipsoid_consts", "nmpc_propagate_ellipsoids_f64", "nmpc_select_corridors_f64",
    "nmpc_fma_peak_probe", "nmpc_adopt_plans_f64", "nmpc_rank_longest_first",
    "FORCESNLPsolver_normal_solve", "FORCESNLPsolver_final_solve",
]

_lib = None


def load() -> ctypes.CDLL:
    """Load the CUDA library; there is no fallback -- a missing build is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension is not built "
            "(run `python -m forces_resilient_planner_b200.build` or __graft_entry__.build()). "
            "There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.nmpc_last_error.restype = ctypes.c_char_p
    lib.nmpc_version.restype = ctypes.c_char_p
    lib.nmpc_smem_bytes.restype = ctypes.c_long
    lib.nmpc_default_opts.argtypes = [ctypes.POINTER(NmpcOpts)]
    sig = [_i, _i, _i, _vp, _vp, _vp, _vp, _ip, _i, ctypes.POINTER(NmpcOpts), _vp, _ip, _vp]
    for name in ("nmpc_solve_batch_f64", "nmpc_solve_batch_f32"):
        getattr(lib, name).argtypes = sig + [_vp]
        getattr(lib, name).restype = _i
    for name in ("nmpc_solve_batch_host_f64", "nmpc_solve_batch_host_f32", "nmpc_solve_batch_host_mixed_f64"):
        getattr(lib, name).argtypes = sig
        getattr(lib, name).restype = _i
    lib.nmpc_solve_batch_ex_f64.argtypes = sig + [_vp] * 5
    lib.nmpc_solve_batch_ex_f64.restype = _i
    for name in ("nmpc_solve_batch_mixed_f64", "nmpc_solve_batch_lowlatency_f64"):
        getattr(lib, name).argtypes = sig + [_vp] * 6
        getattr(lib, name).restype = _i
    _lib = lib
    return lib


def last_error() -> str:
    return load().nmpc_last_error().decode()


def default_opts(**kw) -> NmpcOpts:
    """nmpc_default_opts: the reference tolerances (1e-4), for the fp64 and the mixed-precision entry points alike."""
    o = NmpcOpts()
    load().nmpc_default_opts(ctypes.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(f"nmpc_opts has no field {k!r}")
        setattr(o, k, v)
    return o
