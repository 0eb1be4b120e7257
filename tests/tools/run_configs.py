"""Test tool (config 3 is checked against the CPU oracle on a sample, hence under tests/).
Full-size runs of BASELINE configs 3 and 4 (parity-test configs, not the bench line).

  python tests/tools/run_configs.py --config 3                       # 1 GPU: B = 65536, N = 20, 4-10 rows
  torchrun --nproc-per-node 8 ... tests/tools/run_configs.py --config 4   # 8 GPUs: B = 262144, N = 40, wind sweep,
                                                                       # solve + in-place NCCL all-gather (C ABI)

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
    """BASELINE config 4: 262144 problems (512 x 512 wind sweep), N = 40, sharded over the GPUs of the box; solve +
    end-of-batch collation through the C ABI's own NCCL path (nmpc_solve_batch_sharded_*: the kernel writes into its
    slice of one NCCL-registered buffer, in-place ncclAllGather on the same stream)."""
    import torch.distributed as dist
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("gloo")                  # plumbing only: the NCCL id travels over it
    side, N = 512, 40
    B = side * side
    per = D.per_rank(B, world)
    lo, hi = D.shard_range(B, rank, world)
    idx = np.arange(lo, hi)
    mag = np.linspace(0, 4, side)[idx // side]; az = np.linspace(0, 2 * np.pi, side, endpoint=False)[idx % side]
    fext = np.stack([mag * np.cos(az), mag * np.sin(az), np.zeros(hi - lo)], -1)
    b = W.config2(hi - lo, N, seed=W.SEED + 4 + 1000 * rank, fext=fext)
    assert b.B == per, "512*512 divides evenly over 1/2/4/8 ranks"
    col = D.NcclCollator(rank, world, dev) if world > 1 else None
    out = {"config": f"config4: B={B} (512x512 wind sweep), N=40, sharded over {world} GPU(s), solve + in-place NCCL all-gather of z and info",
           "nccl_version": int(_lib.load().nmpc_comm_nccl_version()) if world > 1 else None}
    st = torch.cuda.current_stream(dev)
    for name, dt, opts, mixed in (("fp64_reference_tolerances", np.float64, _lib.default_opts(), False),
                                  ("fp64_predictor_corrector", np.float64, _lib.default_opts(pc=1, mu0=10.0), False),
                                  ("mixed_f32_reference_tolerances", np.float32, _lib.default_opts(), True)):
        db = S.DeviceBatch(b, dt, dev)
        t_dt = torch.float64 if dt == np.float64 else torch.float32
        if col:
            z_all = col.alloc((world * per, N, 17), t_dt); ii_all = col.alloc((world * per, 4), torch.int32)
            run = lambda: col.solve_sharded(db, z_all, ii_all, opts, mixed=mixed)
        else:
            z_all, ii_all = db.z, db.info_int
            run = lambda: S.solve_device(db, opts, mixed=mixed)
        run(); torch.cuda.synchronize(dev)                      # warm-up (also NCCL's first-collective set-up)
        ms_all, ms_solve = [], []
        for rep in range(3):
            if world > 1:
                dist.barrier()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(st); run(); e1.record(st); torch.cuda.synchronize(dev)
            ms_all.append(e0.elapsed_time(e1))
            e0.record(st); S.solve_device(db, opts, mixed=mixed); e2.record(st); torch.cuda.synchronize(dev)   # the solve alone, for the split
            ms_solve.append(e0.elapsed_time(e2))
        t = torch.tensor([min(ms_all), min(ms_solve)], dtype=torch.float64)
        mine = ii_all[rank * per:(rank + 1) * per].cpu().numpy() if col else ii_all.cpu().numpy()
        cnt = torch.tensor([float(np.sum(mine[:, 0] == 1)), float(mine[:, 1].sum()), float(mine[:, 3].sum())], dtype=torch.float64)
        mx = torch.tensor([float(mine[:, 1].max())], dtype=torch.float64)
        # checksum of checksums: the gathered buffer on every rank sums to the sum of the shards' own sums
        local_sum = (z_all[rank * per:(rank + 1) * per] if col else z_all).double().sum().cpu().reshape(1)
        full_sum = z_all.double().sum().cpu().reshape(1)
        flags_all = ii_all[:, 0].cpu().numpy()
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(local_sum, op=dist.ReduceOp.SUM)
        ok = bool(torch.isclose(full_sum, local_sum, rtol=1e-9, atol=1e-3)) and z_all.shape[0] == B
        out[name] = dict(solves_per_sec=B / (float(t[0]) * 1e-3), ms_solve_plus_collation=float(t[0]), ms_solve_only=float(t[1]),
                         collation_bytes=int(z_all.numel() * z_all.element_size() + ii_all.numel() * 4),
                         converged_frac=float(cnt[0]) / B, converged_frac_seen_in_gathered_flags=float(np.mean(flags_all == 1)),
                         mean_it=float(cnt[1]) / B, max_it=int(mx[0]), resolved_in_fp64_frac=float(cnt[2]) / B,
                         checksum_of_checksums_ok=ok)
        del z_all, ii_all
    if rank == 0:
        print(json.dumps(out), flush=True)
    if col:
        col.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[3, 4])
    a = ap.parse_args()
    config3() if a.config == 3 else config4()
