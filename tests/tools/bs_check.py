"""Test tool (uses the CPU oracle as the checker, hence under tests/).  kkt_backsolve_kernel: correctness against the Schur oracle (fp64 / fp32, N = 20 / 40) and CUDA-event
timings at the bench size.  GPU only."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from forces_resilient_planner_b200 import kkt  # noqa: E402
from oracle import oracle as O  # noqa: E402

dev = torch.device("cuda:0")
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6458.4}
tt = lambda a, dt=None: (torch.from_numpy(np.ascontiguousarray(a)).to(dt) if dt else torch.from_numpy(np.ascontiguousarray(a))).to(dev)
out = {}
variants = [0]
for N in (20, 40):
    B = 64
    phi, jc, g, d = kkt.random_kkt_problems(B, N, seed=N)
    for dt, tol in ((torch.float64, 1e-9), (torch.float32, 2e-2)):
        fac, status = kkt.riccati_factor(tt(phi, dt), tt(jc, dt))
        assert torch.all(status == 0)
        for v in variants:
            dz, y = kkt.kkt_backsolve(fac, tt(g, dt), tt(d, dt))
            torch.cuda.synchronize()
            dz, y = dz.double().cpu().numpy(), y.double().cpu().numpy()
            err = 0.0
            for b in range(0, B, 7):
                Phi, C = kkt.dense_from_compact(phi[b], jc[b])
                rc, dz_ref, y_ref = O.kkt_solve(Phi, g[b], C, d[b, :N - 1])
                err = max(err, np.max(np.abs(dz[b] - dz_ref)) / max(1.0, np.max(np.abs(dz_ref))),
                          np.max(np.abs(y[b, 1:] - y_ref[1:])) / max(1.0, np.max(np.abs(y_ref))) * 0.1)
            ok = err < tol and np.all(y[:, 0] == 0)
            print(f"N={N} {dt}: rel err {err:.2e} {'ok' if ok else 'FAIL'}", flush=True)
            out[f"check_N{N}_{str(dt)[-7:]}"] = {"err": float(err), "ok": bool(ok)}

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for N, Bk, dt in ((20, 16384, torch.float64), (40, 8192, torch.float64), (20, 32768, torch.float32)):
    phi, jc, g, d = kkt.random_kkt_problems(Bk, N, seed=1)
    fac, status = kkt.riccati_factor(tt(phi, dt), tt(jc, dt))
    gz, dd = tt(g, dt), tt(d, dt)
    dz = torch.empty_like(gz); yy = torch.empty_like(dd)
    item = 8 if dt == torch.float64 else 4
    for v in variants:
        for _ in range(3):
            kkt.kkt_backsolve(fac, gz, dd, dz, yy)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); kkt.kkt_backsolve(fac, gz, dd, dz, yy); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.mean(ts))
        by = kkt.algorithmic_bytes(N, item) * Bk
        gbs = by / (ms * 1e-3) / 1e9
        print(f"N={N} B={Bk} {dt}: {ms * 1e3:.1f} us (min {min(ts) * 1e3:.1f}) -> {gbs:.0f} GB/s = "
              f"{100 * gbs / peaks['hbm_gbs']:.1f} % of {peaks['hbm_gbs']:.0f}", flush=True)
        out[f"time_N{N}_B{Bk}_{str(dt)[-7:]}"] = {"ms": ms, "min_ms": float(min(ts)), "gbs": gbs, "frac": gbs / peaks["hbm_gbs"]}
    del fac, gz, dd, dz, yy
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bs_check.json", "w"), indent=1)
