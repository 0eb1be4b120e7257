"""Shared test helpers: KKT residuals of a candidate solution, evaluated with a MODEL ORACLE.

`model` is either oracle.oracle.model_eval (our C restatement; travels to the GPU box) or
oracle.ref_model.RefModel(...).eval (the reference's own CasADi callbacks, oracle/_ref).
Both follow the reference callback's conventions (dense, column-major Jacobians -> returned as
[13,17] / [30,17] arrays here).
"""
from __future__ import annotations

import numpy as np

from forces_resilient_planner_b200 import workloads as W
from oracle import model_np as M

HU = 1e-5


def e_select(z):
    """E z = [x(9); u_prev(4)] (matlab_code/mpc/normal/mpc_generator_normal.m:4-5)."""
    return np.concatenate([z[8:17], z[4:8]])


def kkt_residuals(batch, i, z, y, zl, zu, lc, model):
    """inf-norms (stationarity, equality, inequality, complementarity) of ForcesPro's acceptance
    test (TolStat/TolEq/TolIneq/TolComp) at (z, multipliers) for problem i of `batch`.

    Stage-0 states are fixed by the xinit equality (its multiplier absorbs their stationarity
    row), so stationarity is checked on the free variables, as the solver does.
    """
    N = batch.N
    xinit, _, allp = W.to_forces_params(batch, i) if N == 20 else (batch.xinit[i], None, None)
    rs = req = rin = rcomp = 0.0
    req = max(req, float(np.max(np.abs(z[0, 8:17] - batch.xinit[i]))))
    ev = []
    for k in range(N):
        p = np.zeros(130)
        p[0:10] = batch.hdr[i, k]
        m = int(min(batch.nrows[i, k], 30))
        p[10:10 + 3 * m] = batch.rows[i, k, :m, 0:3].reshape(-1)
        p[100:100 + m] = batch.rows[i, k, :m, 3]
        ev.append((model(z[k], p, k), m))
    for k in range(N):
        e, m = ev[k]
        r = e["grad"].copy() - zl[k] + zu[k]
        if k < N - 1:
            r += e["jc"].T @ y[k + 1]
            req = max(req, float(np.max(np.abs(e["c"] - e_select(z[k + 1])))))
        if k > 0:
            r[8:17] -= y[k, 0:9]
            r[4:8] -= y[k, 9:13]
            r += e["jh"][:m].T @ lc[k, :m]
            rin = max(rin, float(np.max(np.maximum(e["h"][:m] - HU, 0.0), initial=0.0)))
            s = HU - e["h"][:m]
            rcomp = max(rcomp, float(np.max(np.abs(s * lc[k, :m]), initial=0.0)))
        free = np.ones(17, bool)
        if k == 0:
            free[8:] = False
        rs = max(rs, float(np.max(np.abs(r[free]))))
        rin = max(rin, float(np.max(np.maximum(M.LB - z[k], 0)[free])), float(np.max(np.maximum(z[k] - M.UB, 0)[free])))
        rcomp = max(rcomp, float(np.max(((z[k] - M.LB) * zl[k])[free])), float(np.max(((M.UB - z[k]) * zu[k])[free])))
    return rs, req, rin, rcomp


def total_cost(batch, i, z, model):
    f = 0.0
    for k in range(batch.N):
        p = np.zeros(130)
        p[0:10] = batch.hdr[i, k]
        f += model(z[k], p, k)["f"]
    return f


def kkt_residuals_batch(batch, z, y, zl, zu, lc, variant=0, prefer_reference=True):
    """ForcesPro's acceptance test for EVERY problem of a batch (oracle/kkt_check.c), driven with the reference's own
    FORCESNLPsolver_{normal,final}_casadi2forces out of oracle/_ref when that is present (else with the golden-pinned
    restatement behind the same signature).  Returns (res [B,6], used_reference): columns = stationarity, equality,
    inequality, complementarity (max |slack * multiplier|), smallest multiplier, cost."""
    import ctypes
    import os
    import subprocess
    from oracle import oracle as O
    from oracle import ref_model
    here = os.path.dirname(os.path.abspath(O.__file__))
    path = os.path.join(here, "libnmpc_kktcheck.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", here, "libnmpc_kktcheck.so"], stdout=subprocess.DEVNULL)
    chk = ctypes.CDLL(path)
    name = "final" if variant else "normal"
    use_ref = prefer_reference and ref_model.available()
    if use_ref:
        cb_lib = ref_model.RefModel(name).lib
        cb = getattr(cb_lib, f"FORCESNLPsolver_{name}_casadi2forces")
    else:
        cb_lib = O._lib(np.float64)
        cb = getattr(cb_lib, f"nmpc_oracle_casadi2forces_{name}")
    a = lambda x, dt=np.float64: np.ascontiguousarray(x, dt)
    B, N, mcap = batch.B, batch.N, batch.mcap
    xinit, hdr, rows, nrows = a(batch.xinit), a(batch.hdr), a(batch.rows), a(batch.nrows, np.int32)
    z, y, zl, zu = a(z), a(y), a(zl), a(zu)
    lc = a(lc) if mcap else np.zeros((B, N, 1))
    res = np.zeros((B, 6))
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    chk.nmpc_kkt_check.restype = ctypes.c_int
    chk.nmpc_kkt_check.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 10
    rc = chk.nmpc_kkt_check(ctypes.cast(cb, ctypes.c_void_p), B, N, mcap, p(xinit), p(hdr), p(rows), p(nrows), p(z), p(y), p(zl),
                            p(zu), p(lc), p(res))
    assert rc == 0, rc
    return res, use_ref


def compare_with_slsqp(batch, z, flag, index, z_slsqp, f_slsqp, tol_z=1e-3, tol_f=1e-5):
    """Our KKT points against an independent solver's (tests/golden/slsqp_kkt_points.npz), instance by instance.

    Per instance: |z - z_slsqp|_inf and the relative objective difference (ours - SLSQP's) / |SLSQP's|, both costs
    evaluated by the same (golden-pinned) model at points that satisfy the constraints to <= 1e-4 / SLSQP's own tolerance.
    The NLP is non-convex, so a mismatch is classified, not hidden:
      same point     |dz| <= tol_z and |df| <= tol_f                                  (SURVEY.md 8c pin 3)
      slsqp worse    SLSQP stopped at a HIGHER cost (it terminates on a line-search failure near the solution: status 8)
      ours worse     we stopped at a higher cost than a feasible SLSQP point: a different, worse local minimum -- a finding
    Returns a dict with the counts, the worst values and the per-instance table."""
    from oracle import oracle as O
    rows = []
    for i, zs, fs in zip(index, z_slsqp, f_slsqp):
        f_ours = total_cost(batch, int(i), z[i].astype(np.float64), lambda zz, p, k: O.model_eval(zz, p, k, batch.N, batch.variant))
        dz = float(np.max(np.abs(z[i].astype(np.float64) - zs)))
        df = (f_ours - float(fs)) / abs(float(fs))
        kind = "same" if (dz <= tol_z and abs(df) <= tol_f) else ("slsqp_worse" if df < 0 else ("ours_worse" if df > tol_f else "same_cost_other_point"))
        rows.append(dict(i=int(i), flag=int(flag[i]), dz=dz, df=df, kind=kind))
    kinds = [r["kind"] for r in rows]
    return dict(n=len(rows), n_same_point=kinds.count("same"), n_slsqp_worse=kinds.count("slsqp_worse"),
                n_ours_worse=kinds.count("ours_worse"), n_same_cost_other_point=kinds.count("same_cost_other_point"),
                max_dz_same=max([r["dz"] for r in rows if r["kind"] == "same"], default=0.0),
                max_abs_df_same=max([abs(r["df"]) for r in rows if r["kind"] == "same"], default=0.0),
                not_same=[r for r in rows if r["kind"] != "same"])
