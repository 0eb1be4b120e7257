import sys
sys.path.insert(0, ".")
import numpy as np
from oracle import oracle as O
from forces_resilient_planner_b200 import workloads as W, solver as S, _lib
b = W.config2(512)
g = S.solve_host(b); c = O.solve_batch(b)
bad = np.where(g.it != c["it"])[0]
print("differing:", bad[:20], g.it[bad[:20]], c["it"][bad[:20]])
i = int(bad[np.argmax(np.abs(g.it[bad] - c["it"][bad]))])
bb = b.slice(i, i + 1)
np.set_printoptions(linewidth=200, precision=4)
for k in range(0, 14):
    gg = S.solve_host(bb, opts=_lib.default_opts(maxit=k)); cc = O.solve_batch(bb, opts=O.default_opts(maxit=k))
    print(k, "gpu it", gg.it[0], gg.flag[0], gg.nbt[0], gg.info_real[0], "| cpu it", cc["it"][0], cc["flag"][0], cc["nbt"][0], cc["info_real"][0], "| dz", np.abs(gg.z - cc["z"]).max())
