"""BASELINE config 5: warm-started receding-horizon stream, 1024 agents x 500 replans.
Prints per-replan latency p50/p99 (host-visible: refs on host -> first commands on host) as JSON.

  python scripts/stream_bench.py [--agents 1024] [--replans 500] [--no-graph] [--ellipsoids]           # 1 GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/stream_bench.py     # 8 GPUs:
      the agents are sharded contiguously (128 per GPU), every replan ends with an NCCL all-gather of the
      first commands and exit flags (what a fleet coordinator would read); latency = max over ranks.
"""
import argparse, json, os, sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
from forces_resilient_planner_b200 import distributed as D, stream as ST, workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--agents", type=int, default=1024)
ap.add_argument("--replans", type=int, default=500)
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--ellipsoids", action="store_true", help="propagate the disturbance ellipsoids on the device every replan")
ap.add_argument("--mixed", action="store_true", help="replans through the mixed-precision kernel (nmpc_solve_batch_mixed_f64)")
a = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
lo, hi = D.shard_range(a.agents, rank, world)
batch = W.config2(a.agents).slice(lo, hi)            # the same fleet whatever the number of GPUs
rng = np.random.Generator(np.random.PCG64(W.SEED + 5 + 1000 * rank))
s = ST.RecedingHorizonStream(batch, device=f"cuda:{local}", use_graph=not a.no_graph, dynamic_ellipsoids=a.ellipsoids, mixed=a.mixed)
ext = batch.hdr[:, 0, 3:6].copy()
lat, its, fails, resets = [], [], 0, 0
gcmd = torch.empty((a.agents, 5), dtype=torch.float64, device=dev) if world > 1 else None
WARM = 3     # cycle 0 cold start, cycle 1 direct warm launch, cycle 2 graph capture + first replay: not timed
for step in range(a.replans):
    if step == WARM and world > 1:      # the process group comes up after the graph capture (its watchdog thread
        import torch.distributed as dist    # must not touch the device while a stream is capturing)
        dist.init_process_group("nccl", device_id=dev)
    multi = world > 1 and step >= WARM
    ref, yaw, ext = ST.synthetic_refs(batch, step, rng, ext)
    torch.cuda.synchronize()
    if multi:
        dist.barrier()
    t0 = time.perf_counter()
    cmd, flag, it = s.replan(ref, yaw, ext)
    if multi:   # fleet-wide collation: first command + exit flag of every agent on every rank
        mine = torch.cat([s.z[:, 0, 0:4], s.info_int[:, 0:1].double()], dim=1).contiguous()
        if a.agents % world == 0:
            dist.all_gather_into_tensor(gcmd, mine)
        else:
            parts = [torch.empty((D.shard_range(a.agents, r, world)[1] - D.shard_range(a.agents, r, world)[0], 5),
                                 dtype=torch.float64, device=dev) for r in range(world)]
            dist.all_gather(parts, mine); gcmd = torch.cat(parts)
        _ = gcmd[:, 4].cpu()
    lat.append(time.perf_counter() - t0)
    its.append(it.mean()); fails += int((flag != 1).sum())
    resets += int((flag != 1).sum())      # failed agents restart cold next cycle (nmpc_adopt_plans_f64, inside the replan)
lat = np.array(lat[WARM:]) * 1e3
if world > 1:
    t = torch.from_numpy(lat).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)           # a replan is done when the slowest rank is
    lat = t.cpu().numpy()
    c = torch.tensor([float(fails), float(resets), float(np.sum(its[WARM:])) * (hi - lo), float(its[0]) * (hi - lo)], dtype=torch.float64, device=dev)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    fails, resets = int(c[0].item()), int(c[1].item())
    its = [c[3].item() / a.agents] + [0.0] * (WARM - 1) + [c[2].item() / a.agents / max(len(its) - WARM, 1)] * (len(its) - WARM)
if rank == 0:
  print(json.dumps({"config": f"config5: {a.agents} agents x {a.replans} replans, shift warm start, mu0_warm 0.1, {world} GPU(s)" + (f", {hi - lo} agents per GPU, NCCL all-gather of commands + flags every replan" if world > 1 else ""),
                  "latency_ms_p50": float(np.percentile(lat, 50)), "latency_ms_p99": float(np.percentile(lat, 99)),
                  "latency_ms_mean": float(lat.mean()), "replans_per_sec_per_agent": 1e3 / float(lat.mean()),
                  "agent_solves_per_sec": a.agents * 1e3 / float(lat.mean()), "mean_iterations_warm": float(np.mean(its[WARM:])),
                  "mean_iterations_cold": float(its[0]), "failed_solves": fails, "cold_restarts": resets,
                  "cuda_graph": not a.no_graph, "propagated_ellipsoids": a.ellipsoids}))
if world > 1:
    dist.destroy_process_group()
