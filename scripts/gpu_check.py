"""Quick GPU-vs-oracle sanity run (development helper)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from oracle import oracle as O
from forces_resilient_planner_b200 import workloads as W, solver as S

def cmp(name, b, dtype=np.float64):
    t = time.time(); g = S.solve_host(b, dtype); tg = time.time() - t
    t = time.time(); c = O.solve_batch(b, dtype); tc = time.time() - t
    dz = np.abs(g.z.astype(np.float64) - c["z"].astype(np.float64)).reshape(b.B, -1).max(1)
    print(f"{name}: B={b.B} gpu {tg*1e3:.1f} ms cpu {tc*1e3:.1f} ms | flags gpu {np.unique(g.flag, return_counts=True)} "
          f"cpu {np.unique(c['flag'], return_counts=True)} | it gpu {g.it.mean():.2f} cpu {c['it'].mean():.2f} "
          f"same-it {(g.it == c['it']).mean():.4f} | max|dz| {dz.max():.3e} median {np.median(dz):.3e}")
    return g, c

cmp("config1", W.config1())
cmp("config2/256", W.config2(256))
cmp("config2/4096", W.config2(4096))
cmp("config3/1024", W.config3(1024))
cmp("config4/256 (N=40)", W.config4(16, 40))
cmp("config2/256 f32", W.config2(256), np.float32)
for _ in range(3):
    t = time.time(); g = S.solve_host(W.config2(4096)); print("host e2e 4096:", time.time() - t)
