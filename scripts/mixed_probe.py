"""Timing probe of the mixed-precision kernel (and the fp64 kernel beside it) on BASELINE configs 2 / 3.
  python scripts/mixed_probe.py [--config 2|3] [--batch B] [--reps R]
Prints one JSON line; run it under ncu (-k regex:nmpc_ipm_mixed) for the kernel's profile."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from forces_resilient_planner_b200 import _lib, solver as S, workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=3)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--horizon", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda", 0)
B = a.batch or (4096 if a.config == 2 else 65536)
b = W.config2(B, a.horizon) if a.config == 2 else W.config3(B, a.horizon)
out = {"config": a.config, "batch": B, "horizon": a.horizon}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, dt, kw, opts in (("fp64", np.float64, {}, _lib.default_opts()),
                           ("mixed_f64io", np.float64, dict(mixed=True), _lib.default_opts()),
                           ("mixed_f32io", np.float32, {}, _lib.default_opts()),
                           ("mixed_f32io_no_resolve", np.float32, {}, _lib.default_opts(mixed=-1))):
    db = S.DeviceBatch(b, dt, dev)
    S.solve_device(db, opts, **kw); torch.cuda.synchronize(dev)
    st = torch.cuda.current_stream(dev)
    ms = []
    for _ in range(a.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); S.solve_device(db, opts, **kw); e1.record(st); torch.cuda.synchronize(dev)
        ms.append(e0.elapsed_time(e1))
    r = db.result()
    out[name] = dict(ms=min(ms), solves_per_sec=B / (min(ms) * 1e-3), converged=float(np.mean(r.flag == 1)),
                     flags={int(k): int(v) for k, v in zip(*np.unique(r.flag, return_counts=True))},
                     resolved=float(np.mean(r.resolved == 1)), mean_it=float(r.it.mean()), max_it=int(r.it.max()),
                     smem=int(_lib.load().nmpc_smem_bytes(a.horizon, db.mcap, 8 if name == "fp64" else 4)))
print(json.dumps(out), flush=True)
