"""Latency of ONE launch over a small fleet: one-warp fp64 / one-warp mixed / warp-group kernels (cold and warm starts)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from forces_resilient_planner_b200 import _lib, solver as S, workloads as W
dev = torch.device("cuda", 0)
out = {}
for B in (1, 128, 296, 1024):
    b = W.config2(B)
    ref = None
    for name, kw in (("fp64_warp", {}), ("mixed_warp", dict(mixed=True)), ("mixed_group", dict(lowlatency=True))):
        db = S.DeviceBatch(b, np.float64, dev)
        o = _lib.default_opts()
        S.solve_device(db, o, **kw); torch.cuda.synchronize()
        st = torch.cuda.current_stream(dev); ms = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); S.solve_device(db, o, **kw); e1.record(st); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        r = db.result()
        if ref is None: ref = r
        dz = np.abs(r.z - ref.z).reshape(B, -1).max(1)
        out[f"B{B}_{name}"] = dict(ms=min(ms), us_per_iteration_of_slowest=1e3 * min(ms) / r.it.max(), flags={int(k): int(v) for k, v in zip(*np.unique(r.flag, return_counts=True))},
                                   mean_it=float(r.it.mean()), max_it=int(r.it.max()), same_it_as_fp64=float(np.mean(r.it == ref.it)), max_dz_vs_fp64=float(dz.max()),
                                   resolved=float(r.resolved.mean()))
print(json.dumps(out, indent=1))
