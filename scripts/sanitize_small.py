"""Small workload for compute-sanitizer (memcheck / racecheck): every kernel once, few problems."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from forces_resilient_planner_b200 import workloads as W, solver as S, kkt, prep, _lib
g = S.solve_host(W.config3(12))
print("solve", g.flag.tolist(), g.it.tolist())
g = S.solve_host(W.config4(3, 40))
print("solve N=40", g.flag.tolist())
phi, jc, gg, d = kkt.random_kkt_problems(6, 20, seed=1)
t = lambda a: torch.from_numpy(a).cuda()
fac, st = kkt.riccati_factor(t(phi), t(jc)); dz, y = kkt.kkt_backsolve(fac, t(gg), t(d))
torch.cuda.synchronize(); print("kkt", st.tolist(), float(dz.abs().max()))
z = t(np.random.default_rng(0).normal(size=(5, 20, 17))); x, z0 = prep.shift_warm_start(z); torch.cuda.synchronize(); print("shift ok")
rng = np.random.default_rng(1)
path = t(np.cumsum(rng.normal(scale=0.05, size=(7, 30, 3)), axis=1)); size = torch.tensor([1, 2, 6, 30, 30, 12, 29], dtype=torch.int32).cuda()
rp, ry, far = prep.sample_reference(path, size, t(rng.uniform(0, 1, 7)), t(rng.uniform(-3, 3, 7)), 20, 0.05, pos1=t(rng.normal(size=(7, 3))))
E = prep.propagate_ellipsoids(t(g.z[:3, :20].astype(np.float64).copy()))
torch.cuda.synchronize(); print("sample + ellipsoids ok", float(E.abs().max()))
cloud = t(rng.uniform([-1, -2, 0], [5, 2, 2], (200, 3))); ref = np.zeros((3, 20, 3)); ref[:, :, 0] = 0.2 * np.arange(20); ref[:, :, 2] = 1.0
out = prep.select_corridors(cloud, torch.tensor([200], dtype=torch.int32).cuda(), t(ref), t(np.zeros((3, 20))), E, max_polys=20, max_rows=30)
torch.cuda.synchronize(); print("corridors ok", out[4].tolist(), out[5].tolist())
zz = t(np.random.default_rng(2).normal(scale=3.0, size=(4, 20, 17))); prep.wrap_yaw(zz); torch.cuda.synchronize(); print("wrap ok")
g = S.solve_host(W.config3(12), opts=_lib.default_opts(pc=1, mu0=10.0)); print("solve pc", g.flag.tolist(), g.it.tolist())
g = S.solve_host(W.config4(3, 40), opts=_lib.default_opts(pc=1, mu0=10.0)); print("solve pc N=40", g.flag.tolist())
