"""How much does one resident warp slow down when others share its SM?  Fused kernel, config 2, B = 148 * w."""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from forces_resilient_planner_b200 import solver as S, workloads as W, _lib
dev = torch.device("cuda:0")
full = W.config2(148 * 8)
for w in (1, 2, 3, 4, 5, 6, 8):
    b = full.slice(0, 148 * w)
    db = S.DeviceBatch(b, np.float64, dev)
    S.solve_device(db, _lib.default_opts()); torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); S.solve_device(db, _lib.default_opts()); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    r = db.result()
    print(f"B = 148 x {w}: {min(ms):.3f} ms  ({148 * w / min(ms) * 1e3:.0f} solves/s)  mean it {r.it.mean():.2f} max it {r.it.max()}")
