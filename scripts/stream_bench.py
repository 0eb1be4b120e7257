"""BASELINE config 5: warm-started receding-horizon stream, 1024 agents x 500 replans on one GPU.
Prints per-replan latency p50/p99 (host-visible: refs on host -> first commands on host) as JSON.

  python scripts/stream_bench.py [--agents 1024] [--replans 500] [--no-graph]
"""
import argparse, json, sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
from forces_resilient_planner_b200 import stream as ST, workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--agents", type=int, default=1024)
ap.add_argument("--replans", type=int, default=500)
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--ellipsoids", action="store_true", help="propagate the disturbance ellipsoids on the device every replan")
a = ap.parse_args()
batch = W.config2(a.agents)
rng = np.random.Generator(np.random.PCG64(W.SEED + 5))
s = ST.RecedingHorizonStream(batch, use_graph=not a.no_graph, dynamic_ellipsoids=a.ellipsoids)
ext = batch.hdr[:, 0, 3:6].copy()
lat, its, fails, resets = [], [], 0, 0
for step in range(a.replans):
    ref, yaw, ext = ST.synthetic_refs(batch, step, rng, ext)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cmd, flag, it = s.replan(ref, yaw, ext)
    lat.append(time.perf_counter() - t0)
    its.append(it.mean()); fails += int((flag != 1).sum())
    resets += s.reset_failed(flag)
lat = np.array(lat[2:]) * 1e3      # cycles 0/1 are the cold start and the graph warm-up
print(json.dumps({"config": f"config5: {a.agents} agents x {a.replans} replans, shift warm start, mu0_warm 0.1, 1 GPU",
                  "latency_ms_p50": float(np.percentile(lat, 50)), "latency_ms_p99": float(np.percentile(lat, 99)),
                  "latency_ms_mean": float(lat.mean()), "replans_per_sec_per_agent": 1e3 / float(lat.mean()),
                  "agent_solves_per_sec": a.agents * 1e3 / float(lat.mean()), "mean_iterations_warm": float(np.mean(its[2:])),
                  "mean_iterations_cold": float(its[0]), "failed_solves": fails, "cold_restarts": resets,
                  "cuda_graph": not a.no_graph, "propagated_ellipsoids": a.ellipsoids}))
