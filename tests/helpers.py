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
