import sys
sys.path.insert(0, ".")
import numpy as np
from oracle import oracle as O
from forces_resilient_planner_b200 import workloads as W, solver as S, _lib
b = W.config3(2048)
ref = O.solve_batch(b)
for kw in [dict(), dict(tol_stat=1e-2, tol_comp=1e-3, mu_floor=1e-4), dict(tol_stat=2e-2, tol_comp=1e-2, mu_floor=1e-3, tol_eq=1e-4, tol_ineq=1e-4),
           dict(tol_stat=5e-2, tol_comp=1e-2, mu_floor=1e-3, tol_eq=1e-3, tol_ineq=1e-3), dict(tol_stat=5e-2, tol_comp=1e-2, mu_floor=1e-3, tol_eq=1e-3, tol_ineq=1e-3, maxit=60)]:
    g = S.solve_host(b, np.float32, opts=_lib.default_opts(**kw))
    ok = g.flag == 1
    dz = np.abs(g.z.astype(np.float64) - ref["z"]).reshape(b.B, -1).max(1)
    print(kw, "flags", dict(zip(*np.unique(g.flag, return_counts=True))), "it mean", g.it.mean(), "| dz(ok) max %.2e p99 %.2e median %.2e" % (dz[ok].max() if ok.any() else -1, np.quantile(dz[ok], 0.99) if ok.any() else -1, np.median(dz[ok]) if ok.any() else -1))
