"""ctypes binding to oracle/libnmpc_oracle{,_f32}.so.  TEST INFRASTRUCTURE ONLY.

May be imported only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs -- never from forces_resilient_planner_b200/.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class OracleOpts(ctypes.Structure):
    _fields_ = [("mu0", ctypes.c_double), ("sigma", ctypes.c_double), ("mu_floor", ctypes.c_double),
                ("tol_stat", ctypes.c_double), ("tol_eq", ctypes.c_double),
                ("tol_ineq", ctypes.c_double), ("tol_comp", ctypes.c_double),
                ("kappa_push", ctypes.c_double), ("s_floor", ctypes.c_double),
                ("maxit", ctypes.c_int), ("max_bt", ctypes.c_int), ("pc", ctypes.c_int), ("mixed", ctypes.c_int)]


def build(force: bool = False) -> None:
    """Compile the C restatement (and, when /root/reference is present, oracle/_ref)."""
    targets = ["all"]
    if os.path.isdir("/root/reference"):
        targets.append("ref")
    if force:
        subprocess.check_call(["make", "-C", _HERE, "-B"] + targets, stdout=subprocess.DEVNULL)
    else:
        subprocess.check_call(["make", "-C", _HERE] + targets, stdout=subprocess.DEVNULL)


_LIBS = {}


def _lib(dtype):
    dtype = np.dtype(dtype)
    name = "libnmpc_oracle.so" if dtype == np.float64 else "libnmpc_oracle_f32.so"
    if name not in _LIBS:
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        assert lib.nmpc_oracle_sizeof_real() == dtype.itemsize
        lib.nmpc_oracle_default_opts.argtypes = [ctypes.POINTER(OracleOpts)]
        lib.nmpc_oracle_solve_batch.restype = ctypes.c_int
        _LIBS[name] = lib
    return _LIBS[name]


def default_opts(**kw) -> OracleOpts:
    o = OracleOpts()
    _lib(np.float64).nmpc_oracle_default_opts(ctypes.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def solve_batch(batch, dtype=np.float64, opts: OracleOpts | None = None, nthreads: int = 0,
                multipliers: bool = False):
    """Solve a workloads.Batch on the CPU.  Returns dict(z, flag, it, nbt, info_real[, y, zl, zu, lc])."""
    dtype = np.dtype(dtype)
    lib = _lib(dtype)
    B, N, mcap = batch.B, batch.N, batch.mcap
    xinit = np.ascontiguousarray(batch.xinit, dtype)
    z0 = np.ascontiguousarray(batch.z0, dtype)
    hdr = np.ascontiguousarray(batch.hdr, dtype)
    rows = np.ascontiguousarray(batch.rows, dtype)
    nrows = np.ascontiguousarray(batch.nrows, np.int32)
    z = np.zeros((B, N, 17), dtype)
    ii = np.zeros((B, 4), np.int32)
    ir = np.zeros((B, 8), dtype)
    o = opts or default_opts()
    y = np.zeros((B, N, 13), dtype); zl = np.zeros((B, N, 17), dtype); zu = np.zeros((B, N, 17), dtype)
    lc = np.zeros((B, N, max(mcap, 1)), dtype)[:, :, :mcap].copy()
    lib.nmpc_oracle_solve_batch_ex.restype = ctypes.c_int
    rc = lib.nmpc_oracle_solve_batch_ex(B, N, mcap, _p(xinit), _p(z0), _p(hdr), _p(rows), _p(nrows),
                                        int(batch.variant), ctypes.byref(o), _p(z), _p(ii), _p(ir),
                                        _p(y), _p(zl), _p(zu), _p(lc) if mcap else None, int(nthreads))
    if rc != 0:
        raise ValueError(f"nmpc_oracle_solve_batch rejected the arguments (rc={rc})")
    out = dict(z=z, flag=ii[:, 0].copy(), it=ii[:, 1].copy(), nbt=ii[:, 2].copy(), info_real=ir)
    if multipliers:
        out.update(y=y, zl=zl, zu=zu, lc=lc)
    return out


def solve_batch_timed(batch, opts: OracleOpts | None = None):
    """One thread, every solve timed on its own.  Returns dict(z, flag, it, seconds [B])."""
    lib = _lib(np.float64)
    B, N, mcap = batch.B, batch.N, batch.mcap
    a = lambda x, dt=np.float64: np.ascontiguousarray(x, dt)
    xinit, z0, hdr, rows, nrows = a(batch.xinit), a(batch.z0), a(batch.hdr), a(batch.rows), a(batch.nrows, np.int32)
    z = np.zeros((B, N, 17)); ii = np.zeros((B, 4), np.int32); ir = np.zeros((B, 8)); sec = np.zeros(B)
    o = opts or default_opts()
    lib.nmpc_oracle_solve_batch_timed.restype = ctypes.c_int
    rc = lib.nmpc_oracle_solve_batch_timed(B, N, mcap, _p(xinit), _p(z0), _p(hdr), _p(rows), _p(nrows), int(batch.variant),
                                           ctypes.byref(o), _p(z), _p(ii), _p(ir), _p(sec))
    if rc != 0:
        raise ValueError(f"nmpc_oracle_solve_batch_timed rejected the arguments (rc={rc})")
    return dict(z=z, flag=ii[:, 0].copy(), it=ii[:, 1].copy(), seconds=sec)


def model_eval(z, p130, stage, n_stages=20, variant=0):
    """Mirror of the reference casadi2forces callback for one stage (always fp64)."""
    lib = _lib(np.float64)
    z = np.ascontiguousarray(z, np.float64)
    p = np.ascontiguousarray(p130, np.float64)
    f = np.zeros(1); g = np.zeros(17); c = np.zeros(13); jc = np.zeros(13 * 17)
    h = np.zeros(30); jh = np.zeros(30 * 17)
    lib.nmpc_oracle_model_eval(_p(z), _p(p), int(stage), int(n_stages), int(variant),
                               _p(f), _p(g), _p(c), _p(jc), _p(h), _p(jh))
    return dict(f=float(f[0]), grad=g, c=c, jc=jc.reshape(17, 13).T.copy(), h=h,
                jh=jh.reshape(17, 30).T.copy())


def kkt_solve(Phi, g, C, d):
    """Structured KKT solve on explicit blocks (Schur-complement path); fp64."""
    lib = _lib(np.float64)
    N = Phi.shape[0]
    Phi = np.ascontiguousarray(Phi, np.float64); g = np.ascontiguousarray(g, np.float64)
    C = np.ascontiguousarray(C, np.float64); d = np.ascontiguousarray(d, np.float64)
    dz = np.zeros((N, 17)); y = np.zeros((N, 13))
    lib.nmpc_oracle_kkt_solve.restype = ctypes.c_int
    rc = lib.nmpc_oracle_kkt_solve(N, _p(Phi), _p(g), _p(C), _p(d), _p(dz), _p(y))
    return rc, dz, y
