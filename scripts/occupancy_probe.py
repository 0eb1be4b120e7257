"""How much does one resident warp slow down when others share its SM?  Fused kernel, config 2, B = 148 * w,
default algorithm and predictor-corrector."""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from forces_resilient_planner_b200 import solver as S, workloads as W, _lib
dev = torch.device("cuda:0")
full = W.config2(148 * 8)
for label, o in (("default", _lib.default_opts()), ("pc", _lib.default_opts(pc=1, mu0=10.0))):
    for w in (1, 2, 4, 6, 8):
        b = full.slice(0, 148 * w)
        db = S.DeviceBatch(b, np.float64, dev)
        S.solve_device(db, o); torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); S.solve_device(db, o); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        r = db.result()
        print(f"{label:8s} B = 148 x {w}: {min(ms):.3f} ms  ({148 * w / min(ms) * 1e3:.0f} solves/s)  mean it {r.it.mean():.2f} max it {r.it.max()}"
              f"  -> {min(ms) * 1e3 / (r.it.max() + 1):.1f} us per iteration of the slowest problem", flush=True)
