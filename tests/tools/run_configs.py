"""Test tool (config 3 is checked against the CPU oracle on a sample, hence under tests/).
Full-size runs of BASELINE configs 3 and 4 (parity-test configs, not the bench line).

  python tests/tools/run_configs.py --config 3                       # 1 GPU: B = 65536, N = 20, 4-10 rows
  torchrun --nproc-per-node 8 ... tests/tools/run_configs.py --config 4   # 8 GPUs: B = 262144, N = 40, wind sweep,
                                                                       # NCCL all-gather of the results

Each prints one JSON line: solves/s (CUDA events around the fused launch), converged fraction,
iteration statistics, for the fp64 kernel and for the mixed-precision kernel (float arrays), both at the reference tolerances,
plus (config 3) parity against the CPU oracle on a 2048-problem sample and (config 4) the time of
the end-of-batch all-gather with a checksum-of-checksums check.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from forces_resilient_planner_b200 import _lib, distributed as D, solver as S, workloads as W


def timed_solve(batch, dtype, opts, dev, reps=3):
    db = S.DeviceBatch(batch, dtype, dev)
    S.solve_device(db, opts); torch.cuda.synchronize(dev)          # warm-up
    st = torch.cuda.current_stream(dev)
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); S.solve_device(db, opts); e1.record(st); torch.cuda.synchronize(dev)
        ms.append(e0.elapsed_time(e1))
    return db, db.result(), min(ms)


def stats(res, B, ms):
    return dict(solves_per_sec=B / (ms * 1e-3), ms=ms, converged_frac=float(np.mean(res.flag == 1)),
                resolved_in_fp64_frac=float(np.mean(res.resolved == 1)),
                mean_it=float(res.it.mean()), max_it=int(res.it.max()),
                flags={int(k): int(v) for k, v in zip(*np.unique(res.flag, return_counts=True))})


def config3():
    dev = torch.device("cuda", 0)
    b = W.config3(65536)
    out = {"config": "config3: B=65536, N=20, corridor rows 4..10 per problem (ragged), 1 GPU"}
    _, r64, ms64 = timed_solve(b, np.float64, _lib.default_opts(), dev)
    out["fp64_reference_tolerances"] = stats(r64, b.B, ms64)
    _, r32, ms32 = timed_solve(b, np.float32, _lib.default_opts(), dev)
    out["mixed_f32_reference_tolerances"] = stats(r32, b.B, ms32)
    _, r32n, ms32n = timed_solve(b, np.float32, _lib.default_opts(mixed=-1), dev)
    out["mixed_f32_without_fp64_resolve"] = stats(r32n, b.B, ms32n)
    dz = np.abs(r32.z.astype(np.float64) - r64.z).reshape(b.B, -1).max(1)
    out["mixed_vs_fp64_dz"] = dict(max=float(dz.max()), p99=float(np.quantile(dz, 0.99)), median=float(np.median(dz)))
    from oracle import oracle as O                                   # checker on a sample
    sm = b.slice(0, 2048)
    c = O.solve_batch(sm)
    d = np.abs(r64.z[:2048] - c["z"]).reshape(2048, -1).max(1)
    out["fp64_vs_cpu_oracle_sample2048"] = dict(max_dz=float(d.max()), same_iterations=float(np.mean(r64.it[:2048] == c["it"])),
                                                same_flags=bool(np.array_equal(r64.flag[:2048], c["flag"])))
    print(json.dumps(out), flush=True)


def config4():
    import torch.distributed as dist
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    side, N = 512, 40
    B = side * side
    lo, hi = D.shard_range(B, rank, world)
    # wind sweep of the global index range [lo, hi): |f| = linspace(0,4,side) x azimuth linspace(0,2pi,side)
    idx = np.arange(lo, hi)
    mag = np.linspace(0, 4, side)[idx // side]; az = np.linspace(0, 2 * np.pi, side, endpoint=False)[idx % side]
    fext = np.stack([mag * np.cos(az), mag * np.sin(az), np.zeros(hi - lo)], -1)
    b = W.config2(hi - lo, N, seed=W.SEED + 4 + 1000 * rank, fext=fext)
    out = {"config": f"config4: B={B} (512x512 wind sweep), N=40, sharded over {world} GPU(s), NCCL all-gather of z"}
    for name, dt, opts in (("fp64_reference_tolerances", np.float64, _lib.default_opts()),
                           ("fp64_predictor_corrector", np.float64, _lib.default_opts(pc=1, mu0=10.0)),
                           ("mixed_f32_reference_tolerances", np.float32, _lib.default_opts())):
        db, res, ms = timed_solve(b, dt, opts, dev)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        cnt = torch.tensor([float(np.sum(res.flag == 1)), float(res.it.sum()), float(res.it.max())], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            mx = cnt[2:].clone(); dist.all_reduce(cnt, op=dist.ReduceOp.SUM); dist.all_reduce(mx, op=dist.ReduceOp.MAX); cnt[2] = mx[0]
        # end-of-batch collation: all ranks get all results (z + flags + iterations)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        if world > 1:
            gz, gf, gi = D.all_gather_results(db.z, db.info_int[:, 0].contiguous(), db.info_int[:, 1].contiguous(), B)
        else:
            gz, gf, gi = db.z, db.info_int[:, 0], db.info_int[:, 1]
        torch.cuda.synchronize(dev)
        gather_ms = (time.perf_counter() - t0) * 1e3
        # checksum of checksums: sum over the gathered tensor == all-reduced sum of the local shards
        local_sum = db.z.double().sum().reshape(1)
        if world > 1:
            dist.all_reduce(local_sum, op=dist.ReduceOp.SUM)
        full_sum = gz.double().sum()
        ok = bool(torch.isclose(full_sum, local_sum[0], rtol=1e-9, atol=1e-3)) and gz.shape[0] == B
        out[name] = dict(solves_per_sec=B / (float(t.item()) * 1e-3), ms=float(t.item()), converged_frac=float(cnt[0].item()) / B,
                         mean_it=float(cnt[1].item()) / B, max_it=int(cnt[2].item()), allgather_ms=gather_ms,
                         allgather_bytes=int(gz.numel() * gz.element_size()), checksum_of_checksums_ok=ok)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[3, 4])
    a = ap.parse_args()
    config3() if a.config == 3 else config4()
