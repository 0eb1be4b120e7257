"""Independent NLP solve with scipy (SLSQP) driving a model oracle -- cross-check of the KKT point.

Only used by tests on a handful of instances (SLSQP on 340 variables takes seconds).
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import minimize

from oracle import model_np as M

HU = 1e-5


def solve_slsqp(batch, i, model, z_start=None, maxiter=300):
    N = batch.N
    xinit = batch.xinit[i]
    P = []
    for k in range(N):
        p = np.zeros(130)
        p[0:10] = batch.hdr[i, k]
        m = int(min(batch.nrows[i, k], 30))
        p[10:10 + 3 * m] = batch.rows[i, k, :m, 0:3].reshape(-1)
        p[100:100 + m] = batch.rows[i, k, :m, 3]
        P.append((p, m))

    def unpack(v):
        return v.reshape(N, 17)

    def fun(v):
        z = unpack(v)
        f, g = 0.0, np.zeros((N, 17))
        for k in range(N):
            e = model(z[k], P[k][0], k)
            f += e["f"]; g[k] = e["grad"]
        return f, g.reshape(-1)

    def eq(v):
        z = unpack(v)
        r = [z[0, 8:17] - xinit]
        for k in range(N - 1):
            e = model(z[k], P[k][0], k)
            r.append(e["c"] - np.concatenate([z[k + 1, 8:17], z[k + 1, 4:8]]))
        return np.concatenate(r)

    def eq_jac(v):
        z = unpack(v)
        J = np.zeros((9 + 13 * (N - 1), N * 17))
        J[0:9, 8:17] = np.eye(9)
        for k in range(N - 1):
            e = model(z[k], P[k][0], k)
            r0 = 9 + 13 * k
            J[r0:r0 + 13, k * 17:(k + 1) * 17] = e["jc"]
            J[r0:r0 + 9, (k + 1) * 17 + 8:(k + 1) * 17 + 17] -= np.eye(9)
            J[r0 + 9:r0 + 13, (k + 1) * 17 + 4:(k + 1) * 17 + 8] -= np.eye(4)
        return J

    rows_A, rows_b = [], []
    for k in range(1, N):
        p, m = P[k]
        for j in range(m):
            a = np.zeros(N * 17)
            a[k * 17 + 8:k * 17 + 11] = p[10 + 3 * j:13 + 3 * j]
            rows_A.append(a); rows_b.append(p[100 + j] + HU)
    A = np.array(rows_A).reshape(-1, N * 17); b = np.array(rows_b)
    cons = [dict(type="eq", fun=eq, jac=eq_jac)]
    if len(b):
        cons.append(dict(type="ineq", fun=lambda v: b - A @ v, jac=lambda v: -A))
    lb = np.tile(M.LB, N); ub = np.tile(M.UB, N)
    lb[8:17] = -np.inf; ub[8:17] = np.inf      # stage-0 states are fixed by the xinit equality
    z0 = batch.z0[i].copy() if z_start is None else z_start.copy()
    z0[0, 8:17] = xinit
    v0 = np.clip(z0.reshape(-1), lb, ub)
    res = minimize(fun, v0, jac=True, bounds=list(zip(lb, ub)), constraints=cons, method="SLSQP",
                   options=dict(maxiter=maxiter, ftol=1e-12))
    return res.x.reshape(N, 17), res
