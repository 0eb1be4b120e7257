import sys; sys.path.insert(0, ".")
import numpy as np, torch
from forces_resilient_planner_b200 import solver as S, workloads as W, _lib
db = S.DeviceBatch(W.config2(4096), np.float64, torch.device("cuda:0"))
o = _lib.default_opts(pc=1, mu0=10.0)
for _ in range(3):
    S.solve_device(db, o); torch.cuda.synchronize()
print(db.result().it.mean())
