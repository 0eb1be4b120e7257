"""Fused kernel with and without the predictor-corrector on config 2 / config 3 sizes (CUDA events)."""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from forces_resilient_planner_b200 import solver as S, workloads as W, _lib
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, b in (("config2 B=4096", W.config2(4096)), ("config3 B=16384", W.config3(16384)), ("config4 N=40 B=4096", W.config4(64, 40))):
    db = S.DeviceBatch(b, np.float64, dev)
    for label, o in (("default", _lib.default_opts()), ("pc mu0=10", _lib.default_opts(pc=1, mu0=10.0)), ("pc mu0=3", _lib.default_opts(pc=1, mu0=3.0)),
                     ("pc mu0=30", _lib.default_opts(pc=1, mu0=30.0))):
        S.solve_device(db, o); torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); S.solve_device(db, o); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        r = db.result()
        print(f"{name:22s} {label:10s}: {np.mean(ms):7.3f} ms  {b.B / np.mean(ms) * 1e3:9.0f} solves/s  it mean {r.it.mean():.2f} max {r.it.max()}  converged {np.mean(r.flag == 1):.4f}", flush=True)
