#!/usr/bin/env python
"""bench.py -- NMPC solves/sec of the receding-horizon hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   CPU baseline arm (see below)

A "step" is one pass of the hot path over one batch of synthetic problems: BASELINE config 2
(batch = 4096 problems per GPU, N = 20 stages, 6-half-space corridors, randomised x0 / goal /
f_ext, cold start, fp64).  Weak scaling: every rank solves its own 4096-problem shard (different
seed); the problems are independent, so the solve has no collective.

  value   whole-job solves/s with the inputs already resident in HBM, device-timed with CUDA
          events around each fused-IPM launch (L2 flushed between steps, outside the events),
          max over ranks.
  e2e     the same through the host-pointer C ABI (nmpc_solve_batch_host_f64): pinned host
          buffers, H2D + solve + D2H inside every timed call.
  collate (N > 1) the same step with the results collated on every rank, two ways: nmpc_solve_batch_sharded_p2p_f64 (the
          solve kernel's epilogue stores every result into all ranks' buffers over NVLink, barrier kernels around it)
          and, under "nccl", nmpc_solve_batch_sharded_f64: every rank's kernel writes into its
          slice of one NCCL-registered buffer and an in-place ncclAllGather on the same stream collates the
          results of all ranks on all ranks -- inside the CUDA events.  `value_with_collation`, bytes, GB/s.
  roofline           the dominant kernel (fused IPM) against its compulsory HBM traffic
  roofline_fma       the same kernel against the measured fp64 FMA peak (what actually bounds it)
  roofline_backsolve the stand-alone KKT backsolve kernel (HBM-bound), the kernel the north star
                     puts the 40 % target on
  mixed              the mixed-precision kernel (fp32 Newton system, fp64 iterate / residuals, REFERENCE
                     tolerances) on config 2 and on BASELINE config 3 (65536 problems, 4-10 rows, float arrays)
  config4            BASELINE config 4: 262144 problems x N = 40, constant-wind sweep, float arrays, sharded over the
                     N GPUs, solve + NCCL collation (mixed-precision kernel)
  config5            BASELINE config 5: 1024 agents x 500 warm-started replans (agents sharded over the N GPUs),
                     host-visible per-replan latency p50 / p99
  cpu_baseline       the CPU port on this box's host cores, same workload: the product's own algorithm
                     (Riccati, fp64) -- the fastest CPU implementation this repo has -- with the ForcesPro-style
                     Schur-complement restatement beside it, single-thread per-solve latency, and one logged
                     attempt to run the reference's own (licence-locked) solver archive

--impl reference runs ONLY the CPU baseline (oracle/, all host threads) on the same config, honouring
--steps / --warmup, and prints the same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "nmpc_solves_per_sec"
UNIT = "solves/s"
BATCH = 4096
HORIZON = 20
FLOPS_GPU_RICCATI = 0.24e6       # per IPM iteration, structured Riccati
FLOPS_CPU_SCHUR = 0.76e6         # per IPM iteration, ForcesPro-style Schur complement, dense A


def config_block(batch):
    """Identical in both arms (the driver compares it)."""
    from forces_resilient_planner_b200 import workloads as W
    return {"workload": (f"config2: batch={batch} per GPU, N={HORIZON}, 9-state/4-input (17-wide stage vector), "
                         f"6-halfspace corridors, randomised x0/goal/f_ext, cold start, fp64"),
            "batch_per_gpu": batch, "horizon": HORIZON, "seed": W.SEED,
            "l2": "flushed between steps (256 MiB device write, outside the timed events)"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-extras", action="store_true", help="only value / e2e / collate (no rooflines, configs 3-5, CPU legs)")
    return ap.parse_args()


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=float(d["hbm_gbs"]), source="MEASURED_PEAKS.json (of measured)")
    return dict(hbm_gbs=6650.0, source="fallback 6.65 TB/s (of fallback)")


def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


# ------------------------------------------------------------------------- CPU baseline arm --
CPU_ALGOS = {"riccati_fp64": 2, "schur_fp64": 0}


def cpu_throughput(batch, steps, warmup, algo="riccati_fp64", nthreads=0):
    """Times oracle/ (test infrastructure; allowed here only as the measured baseline).
    explicit thread count: torchrun exports OMP_NUM_THREADS=1, which would silently serialise the baseline"""
    from oracle import oracle as O
    O.build()
    nthreads = nthreads if nthreads > 0 else host_threads()
    o = O.default_opts(mixed=CPU_ALGOS[algo])
    for _ in range(warmup):
        O.solve_batch(batch, opts=o, nthreads=nthreads)
    times, res = [], None
    for _ in range(steps):
        t0 = time.perf_counter()
        res = O.solve_batch(batch, opts=o, nthreads=nthreads)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    return dict(value=batch.B * steps / total, cores=nthreads, ms_per_step=1e3 * total / steps,
                converged=float(np.mean(res["flag"] == 1)), mean_it=float(res["it"].mean()))


def pct(ms):
    ms = np.asarray(ms)
    return {"mean": float(ms.mean()), "p50": float(np.median(ms)), "p99": float(np.quantile(ms, 0.99)), "n": int(ms.size)}


def cpu_single_thread_latency():
    """one thread, one solve at a time -- how the planner runs its solver
    (num_of_threads = 1, plan_manage/src/forces_normal.cpp:31)."""
    from forces_resilient_planner_b200 import workloads as W
    from oracle import oracle as O
    out = {}
    b1 = W.config1()
    rep = W.Batch(*(np.repeat(a, 200, axis=0) for a in (b1.xinit, b1.z0, b1.hdr, b1.rows, b1.nrows)), b1.variant)
    b2 = W.config2(1000)
    for algo, m in CPU_ALGOS.items():
        o = O.default_opts(mixed=m)
        O.solve_batch_timed(b2.slice(0, 20), o)
        out[algo] = {"config1_x200": pct(O.solve_batch_timed(rep, o)["seconds"] * 1e3),
                     "config2_1000_instances": pct(O.solve_batch_timed(b2, o)["seconds"] * 1e3)}
    out["unit"] = "ms per solve, 1 thread, CLOCK_MONOTONIC around each solve inside the C loop"
    return out


def gpu_single_solve_latency():
    """The reference's own use: ONE vehicle, one FORCESNLPsolver_normal_solve call per replan (host structs in, host structs
    out: packing, H2D, solve, D2H inside the timer).  Same instances as cpu_single_thread_latency(); per kernel choice of
    the shim (NMPC_B200_SHIM)."""
    from forces_resilient_planner_b200 import forces, workloads as W
    out = {}
    b1, b2 = W.config1(), W.config2(1000)
    saved = os.environ.get("NMPC_B200_SHIM")
    try:
        for name, env in (("warp_group_default", None), ("fp64_one_warp", "fp64")):
            if env is None:
                os.environ.pop("NMPC_B200_SHIM", None)
            else:
                os.environ["NMPC_B200_SHIM"] = env
            w = forces.FORCESNormal()
            rec = {}
            for label, batch, idx in (("config1_x200", b1, [0] * 203), ("config2_200_instances", b2, list(range(203)))):
                ts, ok = [], 0
                for n, i in enumerate(idx):
                    xinit, x0, allp = W.to_forces_params(batch, i)
                    w.params_.xinit[:] = xinit.tolist(); w.params_.x0[:] = x0.tolist(); w.params_.all_parameters[:] = allp.tolist()
                    t0 = time.perf_counter()
                    flag = w.solve_params()
                    dt = time.perf_counter() - t0
                    if n >= 3:                                   # first calls: context, arena, streams
                        ts.append(dt * 1e3); ok += int(flag == 1)
                rec[label] = {**pct(np.array(ts)), "converged_frac": ok / len(ts)}
            out[name] = rec
    finally:
        if saved is None:
            os.environ.pop("NMPC_B200_SHIM", None)
        else:
            os.environ["NMPC_B200_SHIM"] = saved
    out["unit"] = "ms per FORCESNLPsolver_normal_solve call (host structs in / out), cold starts"
    return out


def reference_binary_attempt():
    """one real call of the reference's solver archive (oracle/_ref/forces_attempt, linked
    in the build container against SOLN/FORCESNLPsolver_normal/lib/libFORCESNLPsolver_normal.a where it lies)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "forces_attempt")
    if not os.path.exists(exe):
        return {"linked": False, "exitflag": None, "note": "oracle/_ref/forces_attempt not built (needs /root/reference)"}
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        rec = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
        said = " ".join(ln.strip() for ln in out.stdout.splitlines() if ln.strip() and not ln.startswith("{"))
        rec["solver_message"] = said[:200]
        rec["note"] = "exit -100 = LICENSE_ERROR (header :139): the archive runs only on its authors' machine, hence no ForcesPro timing"
        return rec
    except Exception as e:      # noqa: BLE001
        return {"linked": True, "exitflag": None, "note": f"could not run: {e}"}


def run_reference(args):
    rank, _, _ = env_rank()
    if rank != 0:
        return 0
    from forces_resilient_planner_b200 import workloads as W
    batch = W.config2(args.batch, HORIZON)
    r = cpu_throughput(batch, args.steps, args.warmup, "riccati_fp64")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_block(args.batch),
        "stats": {"converged_frac": r["converged"], "mean_iterations": r["mean_it"]},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "algorithm": "primal-dual IPM of this repo with a dense fp64 Riccati KKT solve (oracle opts.mixed = 2): the "
                                      "product's own algorithm on the CPU, 6x faster here than the ForcesPro-style Schur restatement",
                         "sample": f"the whole {args.batch}-problem batch per step ({args.steps} steps after {args.warmup} warm-up), "
                                   "OpenMP over problems, all host threads; the ForcesPro core itself is a licence-locked "
                                   "binary (exit -100), so the CPU arm is this repo's C port (oracle/nmpc_oracle.c)"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------- clocks -----
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc, self.path = None, f"/tmp/nmpc_clocks_{os.getpid()}.csv"
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            busy = [s for s in sm if s > 0.5 * max(sm)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------------- main -----
def run_b200(args):
    import torch
    import torch.distributed as dist
    from forces_resilient_planner_b200 import _lib, distributed as D, kkt, solver as S, workloads as W

    rank, local_rank, world = env_rank()
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    B = args.batch
    batch = W.config2(B, HORIZON, seed=W.SEED + rank)
    db = S.DeviceBatch(batch, np.float64, dev, pinned=True)
    opts = _lib.default_opts()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    lib = _lib.load()

    def timed(fn, steps, warmup):
        """`steps` launches of fn, each bracketed by CUDA events on the launching stream, L2 flushed in between
        (outside the events); returns per-step ms (this rank)."""
        for _ in range(warmup):
            fn()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for e0, e1 in evs:
            flush.zero_()
            e0.record(stream)
            fn()
            e1.record(stream)
        barrier()
        return [e0.elapsed_time(e1) for e0, e1 in evs]

    # ---- value: device-resident inputs, CUDA events on the launching stream ------------------
    sampler = None
    for _ in range(args.warmup):
        S.solve_device(db, opts)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    step_ms = timed(lambda: S.solve_device(db, opts), args.steps, 0)
    total_s = max_over_ranks(sum(step_ms) * 1e-3)
    res = db.result()
    value = B * world * args.steps / total_s
    kernel_ms = sum(step_ms) / len(step_ms)

    # ---- e2e: host-pointer C ABI, pinned host buffers, H2D + solve + D2H every step ----------
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    hz, hii, hir = pin((B, HORIZON, 17), torch.float64), pin((B, 4), torch.int32), pin((B, 8), torch.float64)
    h = db.h
    call = lambda: lib.nmpc_solve_batch_host_f64(
        B, HORIZON, db.mcap, h["xinit"].data_ptr(), h["z0"].data_ptr(), h["hdr"].data_ptr(), h["rows"].data_ptr(),
        h["nrows"].data_ptr(), db.variant, ctypes.byref(opts), hz.data_ptr(), hii.data_ptr(), hir.data_ptr())
    for _ in range(args.warmup):
        assert call() == 0, _lib.last_error()
    barrier()
    e2e_times = []
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        rc = call()
        e2e_times.append(time.perf_counter() - t0)
        assert rc == 0, _lib.last_error()
    barrier()
    e2e_total = max_over_ranks(sum(e2e_times))
    e2e_value = B * world * args.steps / e2e_total
    assert np.array_equal(hz.numpy(), res.z), "host-pointer and device-pointer paths disagree"
    clocks = sampler.stop() if sampler else None

    # ---- collate (N > 1): solve + in-place NCCL all-gather through the C ABI, inside the events ----
    collate = None
    col = None
    if world > 1:
        col = D.NcclCollator(rank, world, dev)
        z_all = col.alloc((world * B, HORIZON, 17), torch.float64)
        ii_all = col.alloc((world * B, 4), torch.int32)
        c_ms = timed(lambda: col.solve_sharded(db, z_all, ii_all, opts), args.steps, max(args.warmup, 3))
        c_total = max_over_ranks(sum(c_ms) * 1e-3)
        bare_ms = timed(lambda: (col.collate(z_all), col.collate(ii_all)), 10, 3)       # the two all-gathers alone
        bare = max_over_ranks(sum(bare_ms) / len(bare_ms))
        # every rank now holds every rank's results: its own slice equals what the plain solve produced
        mine_ok = bool(torch.equal(z_all[rank * B:(rank + 1) * B], db.z))
        flags_all = ii_all[:, 0].cpu().numpy()
        cs_local = sum_over_ranks(float(db.z.double().sum().item()))
        cs_full = float(z_all.double().sum().item())
        nbytes = int(z_all.numel() * 8 + ii_all.numel() * 4)
        collate = {"value_with_collation": B * world * args.steps / c_total, "unit": UNIT,
                   "ms_per_step": 1e3 * c_total / args.steps, "allgather_bytes_per_step": nbytes,
                   "allgather_ms_alone": bare, "allgather_gbs_alone": nbytes / (bare * 1e-3) / 1e9,
                   "own_slice_identical": mine_ok, "converged_frac_all_ranks": float(np.mean(flags_all == 1)),
                   "checksum_of_checksums_ok": bool(abs(cs_full - cs_local) <= 1e-9 * abs(cs_local) + 1e-6),
                   "nccl_version": int(lib.nmpc_comm_nccl_version()),
                   "api": "nmpc_solve_batch_sharded_f64: kernel writes into its slice of one ncclMemAlloc'd + registered "
                          "buffer; in-place ncclAllGather (z + info) on the same stream"}

        # ---- the same step with the exchange fused into the solve kernel: peer stores over NVLink + barrier kernels ----
        pc, why = make_peer_collator(D, dist, torch, rank, world, dev, B, HORIZON, np.float64)
        if pc is None:
            collate = {"p2p_unavailable": why, **collate, "method": "nccl_allgather"}
        else:
            p_ms = timed(lambda: pc.solve_sharded(db, opts), args.steps, max(args.warmup, 3))
            p_total = max_over_ranks(sum(p_ms) * 1e-3)
            barrier_ms = max_over_ranks(float(np.mean(timed(lambda: pc.barrier(), 10, 3))))
            pc.check()
            p_mine_ok = bool(torch.equal(pc.z_all[rank * B:(rank + 1) * B], db.z))
            p_same = bool(torch.equal(pc.z_all, z_all)) and bool(torch.equal(pc.info_all[:, :3], ii_all[:, :3]))
            nccl = collate
            collate = {"method": "peer_stores", "value_with_collation": B * world * args.steps / p_total, "unit": UNIT,
                       "ms_per_step": 1e3 * p_total / args.steps, "peer_store_bytes_per_step_per_gpu": nbytes // world * (world - 1),
                       "barrier_kernel_ms_alone": barrier_ms, "own_slice_identical": p_mine_ok,
                       "identical_to_nccl_allgather_result": p_same,
                       "converged_frac_all_ranks": float(np.mean(pc.info_all[:, 0].cpu().numpy() == 1)),
                       "api": "nmpc_solve_batch_sharded_p2p_f64: the solve kernel's epilogue writes every problem's z and info into its "
                              "slice of every rank's buffer (TMA bulk stores to CUDA-IPC-mapped peer memory); barrier kernel before and after",
                       "nccl": nccl}
            dist.barrier()
            pc.close()

    extras = {}
    if not args.no_extras:
        extras.update(run_config4(args, torch, dist, D, S, W, _lib, dev, rank, world, col, max_over_ranks, sum_over_ranks))
        extras.update(run_config5(torch, dist, W, dev, rank, local_rank, world, max_over_ranks, sum_over_ranks))
    if col:
        col.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- rooflines ------------------------------------------------------------------------------
    peaks = measured_peaks()
    algo_bytes = batch.algorithmic_bytes(8) * B
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "fused_kernel_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roofline = {"kernel": "nmpc_ipm_kernel<double,20>", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peaks["source"],
                "algorithmic_bytes_per_launch": algo_bytes,
                "note": "compulsory I/O only (the whole solve runs out of shared memory); the kernel is FMA-issue/"
                        "latency bound, see roofline_fma; the HBM-bound kernel is roofline_backsolve"}
    if not args.no_extras and world == 1:
        tf = ctypes.c_double(0)
        lib.nmpc_fma_peak_probe(8, ctypes.byref(tf))
        ach_tf = FLOPS_GPU_RICCATI * float(res.it.sum()) / (kernel_ms * 1e-3) / 1e12
        extras["roofline_fma"] = {"kernel": "nmpc_ipm_kernel<double,20>", "bound": "fp64_fma", "achieved": ach_tf,
                                  "peak": tf.value, "unit": "TFLOP/s", "frac": ach_tf / tf.value if tf.value else None,
                                  "peak_source": "nmpc_fma_peak_probe (measured live, fp64 FFMA-chain kernel)",
                                  "algorithmic_flops_per_iteration": FLOPS_GPU_RICCATI}
        extras["roofline_backsolve"] = run_backsolve(torch, kkt, dev, stream, flush, peaks)
        extras["mixed"] = run_mixed(torch, S, W, _lib, dev, stream, flush, batch)
        cb = cpu_throughput(batch, 3, 1, "riccati_fp64")
        cs = cpu_throughput(batch, 1, 0, "schur_fp64")
        extras["cpu_baseline"] = {
            "value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": "port",
            "algorithm": "this repo's IPM with a dense fp64 Riccati KKT solve (oracle opts.mixed = 2): the product's own algorithm "
                         "on the CPU",
            "sample": f"the same {B}-problem batch, 3 timed passes after 1 warm-up, OpenMP over problems, all host threads "
                      f"(oracle/nmpc_oracle.c; ForcesPro binary unrunnable: licence exit -100)",
            "mean_iterations": cb["mean_it"], "converged_frac": cb["converged"],
            "schur_restatement": {"value": cs["value"], "unit": UNIT,
                                  "what": "the same IPM with the KKT solve done the way the ForcesPro binary's symbol table says it "
                                          "does (17x17 Cholesky, 13x13 Schur blocks, block-tridiagonal Cholesky) + 2 refinement rounds",
                                  "flops_per_iteration": {"cpu_schur_dense": FLOPS_CPU_SCHUR, "gpu_riccati_structured": FLOPS_GPU_RICCATI,
                                                          "ratio": FLOPS_CPU_SCHUR / FLOPS_GPU_RICCATI}},
            "note": "a GPU/CPU ratio is a reported baseline, not a quality measure of the kernel (see roofline_fma); against the "
                    "round-1 Schur restatement the ratio was ~6x larger -- that CPU arm did ~3x the flops per iteration plus refinement",
            "single_thread_ms": cpu_single_thread_latency(),
            "reference_binary_attempt": reference_binary_attempt()}
        extras["single_solve_ms"] = gpu_single_solve_latency()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_block(B),
        "stats": {"converged_frac": float(np.mean(res.flag == 1)), "mean_iterations": float(res.it.mean()),
                  "max_iterations": int(res.it.max()), "smem_bytes_per_problem": int(lib.nmpc_smem_bytes(HORIZON, db.mcap, 8))},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": db.h2d_bytes, "d2h_bytes_per_step": db.d2h_bytes,
                "ms_per_step": 1e3 * e2e_total / args.steps, "api": "nmpc_solve_batch_host_f64 (pinned host buffers)"},
        "gpu_launches": args.steps,
        "roofline": roofline,
    }
    if collate:
        line["collate"] = collate
    line.update(extras)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def make_peer_collator(D, dist, torch, rank, world, dev, per, N, dtype):
    """PeerCollator on every rank or on none: CUDA IPC needs the GPUs of one node with P2P access, and a rank that cannot
    map its peers must not leave the others waiting at a barrier kernel."""
    pc, why = None, ""
    try:
        pc = D.PeerCollator(rank, world, dev, per, N, dtype=dtype)
    except RuntimeError as e:
        why = str(e)
    ok = torch.tensor([1.0 if pc is not None else 0.0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if ok.item() < 1.0:
        if pc is not None:
            pc.close()
        return None, why or "a peer rank could not map this rank's buffers"
    return pc, ""


def run_backsolve(torch, kkt, dev, stream, flush, peaks):
    """stand-alone KKT backsolve: factor once, then time backsolves (inputs >> L2)"""
    Bk = 16384
    phi, jc, g, d = kkt.random_kkt_problems(Bk, HORIZON, seed=1)
    tt = lambda a: torch.from_numpy(a).to(dev)
    fac, status = kkt.riccati_factor(tt(phi), tt(jc))
    gz, dd = tt(g), tt(d)
    dz = torch.empty_like(gz); yy = torch.empty_like(dd)
    for _ in range(3):
        kkt.kkt_backsolve(fac, gz, dd, dz, yy)
    torch.cuda.synchronize(dev)
    reps = 10
    bevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for e0, e1 in bevs:
        flush.zero_()
        e0.record(stream)
        kkt.kkt_backsolve(fac, gz, dd, dz, yy)
        e1.record(stream)
    torch.cuda.synchronize(dev)
    bms = sum(e0.elapsed_time(e1) for e0, e1 in bevs) / reps
    bbytes = kkt.algorithmic_bytes(HORIZON, 8) * Bk
    bach = bbytes / (bms * 1e-3) / 1e9
    btraffic = None
    bpath = os.path.join(ROOT, "profiles", "backsolve_kernel_traffic.json")
    if os.path.exists(bpath):
        btraffic = json.load(open(bpath)).get("dram_bytes_per_launch")
    return {"kernel": "kkt_backsolve_kernel<double,20>", "bound": "hbm", "achieved": bach, "peak": peaks["hbm_gbs"],
            "unit": "GB/s", "frac": bach / peaks["hbm_gbs"], "traffic": btraffic, "peak_source": peaks["source"],
            "algorithmic_bytes_per_launch": bbytes, "ms_per_launch": bms, "batch": Bk,
            "backsolves_per_sec": Bk / (bms * 1e-3), "all_factor_ok": bool((status == 0).all().item()),
            "role": "benchmark kernel for the stored-factor split the reference makes (multi-rhs re-solves); the fused "
                    "solver keeps its factor in shared memory and does not call it"}


def run_mixed(torch, S, W, _lib, dev, stream, flush, batch2):
    """The mixed-precision kernel at the REFERENCE tolerances: config 2 (double arrays) and BASELINE config 3 at full size
    (65536 problems, ragged 4-10 rows, float arrays).  Parity of the very results that were timed, against the fp64 kernel."""
    out = {"what": "single-precision Newton system in delta form, double-precision iterate / residuals / line search; "
                   "nmpc_default_opts unchanged (1e-4); problems the fp32 factorisation cannot carry are re-solved by the fp64 kernel "
                   "inside the timed region"}
    o = _lib.default_opts()

    def one(b, dt, mixed, reps=3):
        db = S.DeviceBatch(b, dt, dev)
        S.solve_device(db, o, mixed=mixed); torch.cuda.synchronize(dev)
        ms = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); S.solve_device(db, o, mixed=mixed); e1.record(stream); torch.cuda.synchronize(dev)
            ms.append(e0.elapsed_time(e1))
        return db.result(), sum(ms) / len(ms)

    def entry(b, r, ms, r64):
        dz = np.abs(r.z.astype(np.float64) - r64.z).reshape(b.B, -1).max(1)
        return {"value": b.B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch": b.B,
                "converged_frac": float(np.mean(r.flag == 1)), "resolved_in_fp64_frac": float(np.mean(r.resolved == 1)),
                "mean_iterations": float(r.it.mean()), "max_iterations": int(r.it.max()),
                "max_abs_dz_vs_fp64_kernel": float(dz.max()), "median_abs_dz_vs_fp64_kernel": float(np.median(dz)),
                "max_reported_residual": float(np.max(r.info_real[:, 0:4]))}

    r64, ms64 = one(batch2, np.float64, False)
    rm, msm = one(batch2, np.float64, True)
    out["config2"] = entry(batch2, rm, msm, r64)
    out["config2"]["fp64_kernel_value"] = batch2.B / (ms64 * 1e-3)
    b3 = W.config3(65536)
    r64, ms64 = one(b3, np.float64, False, reps=2)
    rm, msm = one(b3, np.float32, True, reps=2)
    out["config3"] = entry(b3, rm, msm, r64)
    out["config3"]["fp64_kernel_value"] = b3.B / (ms64 * 1e-3)
    out["config3"]["workload"] = "config3: batch=65536, N=20, corridor rows 4..10 per problem (ragged), float arrays in HBM, 1 GPU"
    return out


def run_config4(args, torch, dist, D, S, W, _lib, dev, rank, world, col, max_over_ranks, sum_over_ranks):
    """BASELINE config 4, sharded over the ranks of this run: 512 x 512 constant-wind sweep, N = 40, float arrays,
    mixed-precision kernel, solve + in-place NCCL collation of z and info through the C ABI."""
    side, N = 512, 40
    Btot = side * side
    per = D.per_rank(Btot, world)
    lo, hi = D.shard_range(Btot, rank, world)
    idx = np.arange(lo, hi)
    mag = np.linspace(0, 4, side)[idx // side]; az = np.linspace(0, 2 * np.pi, side, endpoint=False)[idx % side]
    fext = np.stack([mag * np.cos(az), mag * np.sin(az), np.zeros(hi - lo)], -1)
    b = W.config2(hi - lo, N, seed=W.SEED + 4 + 1000 * rank, fext=fext)
    if b.B != per:
        return {"config4": {"skipped": f"262144 does not split evenly over {world} ranks"}}
    db = S.DeviceBatch(b, np.float32, dev)
    o = _lib.default_opts()
    st = torch.cuda.current_stream(dev)
    pc, method = None, "none (one GPU)"
    if world > 1:
        pc, _ = make_peer_collator(D, dist, torch, rank, world, dev, per, N, np.float32)
    if pc is not None:
        z_all, ii_all = pc.z_all, pc.info_all
        run = lambda: pc.solve_sharded(db, o)
        method = "peer stores from the solve kernel's epilogue + barrier kernels (nmpc_solve_batch_sharded_p2p_f32)"
    elif col:
        z_all = col.alloc((world * per, N, 17), torch.float32); ii_all = col.alloc((world * per, 4), torch.int32)
        run = lambda: col.solve_sharded(db, z_all, ii_all, o)
        method = "in-place NCCL all-gather of z and info (nmpc_solve_batch_sharded_f32)"
    else:
        z_all, ii_all = db.z, db.info_int
        run = lambda: S.solve_device(db, o)
    run(); torch.cuda.synchronize(dev)
    ms = []
    for _ in range(2):
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); run(); e1.record(st); torch.cuda.synchronize(dev)
        ms.append(e0.elapsed_time(e1))
    t = max_over_ranks(min(ms))
    flags = ii_all.cpu().numpy()
    mine = flags[rank * per:(rank + 1) * per]
    cs_local = sum_over_ranks(float(db.z.double().sum().item()) if world == 1 else float(z_all[rank * per:(rank + 1) * per].double().sum().item()))
    cs_full = float(z_all.double().sum().item())
    it_sum = sum_over_ranks(float(mine[:, 1].sum())); res_sum = sum_over_ranks(float(mine[:, 3].sum()))
    it_max = max_over_ranks(float(mine[:, 1].max()))
    conv = float(np.mean(flags[:, 0] == 1)); cs_full_ok = bool(abs(cs_full - cs_local) <= 1e-6 * abs(cs_local) + 1e-3)
    nbytes = int(z_all.numel() * 4 + ii_all.numel() * 4) if world > 1 else 0
    if pc is not None:
        pc.check()
        del z_all, ii_all
        dist.barrier()
        pc.close()
    return {"config4": {
        "workload": f"config4: batch=262144 (512x512 constant-wind sweep, |f| 0..4 m/s^2), N=40, float arrays, sharded {per} per GPU over "
                    f"{world} GPU(s), mixed-precision kernel, solve + collation of z and info on every rank",
        "value": Btot / (t * 1e-3), "unit": UNIT, "ms_solve_plus_collation": t, "n_gpus": world, "collation": method,
        "collation_bytes": nbytes, "converged_frac_all_ranks": conv, "mean_iterations": it_sum / Btot, "max_iterations": int(it_max),
        "resolved_in_fp64_frac": res_sum / Btot, "checksum_of_checksums_ok": cs_full_ok}}


def run_config5(torch, dist, W, dev, rank, local_rank, world, max_over_ranks, sum_over_ranks, agents=1024, replans=500):
    """BASELINE config 5: 1024 agents x 500 warm-started replans, agents sharded over the ranks; per replan: references on the
    host -> shift / pack / solve on the device (CUDA graph, one pinned H2D + D2H copy inside it) -> first commands on the
    host.  Latency = max over ranks.  Two streams: the mixed-precision kernels (the warp-group kernel when a GPU holds at
    most one agent per SM, the one-warp kernel otherwise) and the fp64 kernel."""
    from forces_resilient_planner_b200 import distributed as D, stream as ST
    lo, hi = D.shard_range(agents, rank, world)
    batch = W.config2(agents).slice(lo, hi)
    out = {"workload": f"config5: {agents} agents x {replans} replans, warm-started receding horizon (adopt + shift + pack + solve on the "
                       f"device, CUDA graph), {hi - lo} agents per GPU on {world} GPU(s); latency = refs on host -> first commands on host, "
                       f"max over ranks", "n_gpus": world}
    WARM = 3
    for name, mixed in (("mixed", True), ("fp64", False)):
        rng = np.random.Generator(np.random.PCG64(W.SEED + 5 + 1000 * rank))
        s = ST.RecedingHorizonStream(batch, device=f"cuda:{local_rank}", use_graph=True, mixed=mixed)
        ext = batch.hdr[:, 0, 3:6].copy()
        lat, its, fails = [], [], 0
        for step in range(replans):
            ref, yaw, ext = ST.synthetic_refs(batch, step, rng, ext)
            torch.cuda.synchronize(dev)
            if world > 1 and step >= WARM:
                dist.barrier()
            t0 = time.perf_counter()
            cmd, flag, it = s.replan(ref, yaw, ext)
            lat.append(time.perf_counter() - t0)
            its.append(float(it.mean())); fails += int((flag != 1).sum())
        lat = np.array(lat[WARM:]) * 1e3
        if world > 1:
            t = torch.from_numpy(lat).to(dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            lat = t.cpu().numpy()
        out[name] = {"latency_ms": {"p50": float(np.median(lat)), "p99": float(np.quantile(lat, 0.99)), "mean": float(lat.mean())},
                     "kernel": ("nmpc_ipm_group_kernel (256 threads per agent)" if s.lowlatency else
                                ("nmpc_ipm_mixed_kernel" if mixed else "nmpc_ipm_kernel<double>")),
                     "agent_replans_per_sec": agents / (float(lat.mean()) * 1e-3), "mean_warm_iterations": float(np.mean(its[WARM:])),
                     "failed_solves": int(sum_over_ranks(float(fails)))}
        del s
    out["latency_ms"] = out["mixed"]["latency_ms"]
    return {"config5": out}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
