"""CUDA-event timings of the kernels around the solve for a 1024-agent fleet."""
import json, sys
sys.path.insert(0, ".")
import numpy as np, torch
from forces_resilient_planner_b200 import prep, workloads as W, solver as S

dev = torch.device("cuda:0")
t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev)
B, N, P, M = 1024, 20, 60, 512
rng = np.random.default_rng(0)
z = t(S.solve_host(W.config2(B)).z)                      # realistic previous plans
k = np.arange(P)
paths = np.zeros((B, P, 3)); head = rng.uniform(-np.pi, np.pi, B)
ang = head[:, None] + 0.25 * np.sin(0.08 * k[None] + rng.uniform(0, 6, (B, 1)))
paths[:, :, 0] = np.cumsum(0.05 * np.cos(ang), 1); paths[:, :, 1] = np.cumsum(0.05 * np.sin(ang), 1); paths[:, :, 2] = 1.0
cloud = rng.uniform([-4, -4, 0], [4, 4, 2.5], (B, M, 3))
far = np.min(np.linalg.norm(cloud[:, :, None, :2] - paths[:, None, ::6, :2], axis=3), axis=2) > 0.8
cn = far.sum(1).astype(np.int32)
for a in range(B):
    cloud[a, :cn[a]] = cloud[a][far[a]]
d = dict(paths=t(paths), size=t(np.full(B, P), torch.int32), toff=t(rng.uniform(0, 0.5, B)), last=t(rng.uniform(-1, 1, B)),
         cloud=t(cloud), cn=t(cn, torch.int32), ext=t(rng.uniform(-1, 1, (B, 3))))


def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for e0, e1 in ev:
        e0.record(); out = fn(); e1.record()
    torch.cuda.synchronize()
    return float(np.median([e0.elapsed_time(e1) for e0, e1 in ev])) * 1e3, out


res = {"fleet": f"{B} agents, N = {N}, front-end polyline of {P} points, {M}-slot obstacle clouds ({int(cn.mean())} live points on average)"}
res["shift_warm_start_us"], _ = timed(lambda: prep.shift_warm_start(z))
res["propagate_ellipsoids_us"], E = timed(lambda: prep.propagate_ellipsoids(z))
res["sample_reference_us"], (rp, ry, _) = timed(lambda: prep.sample_reference(d["paths"], d["size"], d["toff"], d["last"], N, 0.05))
res["select_corridors_us"], (pA, pb, pm, pidx, npoly, ovf) = timed(lambda: prep.select_corridors(d["cloud"], d["cn"], rp, ry, E, max_polys=20, max_rows=30))
res["pack_params_us"], _ = timed(lambda: prep.pack_params(rp, ry, d["ext"], E, pA, pb, pm, pidx, (7.0, 1.0, 80.0, 12.0, 0.5), 30))
res["polytopes_per_agent_mean"] = float(npoly.float().mean()); res["rows_per_polytope_mean"] = float(pm[pm > 0].float().mean())
res["corridor_overflow_agents"] = int((ovf != 0).sum())
print(json.dumps(res))
