"""Generates tests/golden/slsqp_kkt_points.npz: an INDEPENDENT solver's answer to 100 of the synthetic NLPs.

scipy SLSQP (a sequential-QP method: no barrier, no Riccati, nothing shared with this repo's solver), started from the
planner's cold guess, driving the REFERENCE's own model callbacks (oracle/_ref: FORCESNLPsolver_normal_casadi2forces,
compiled from /root/reference where it lies) through ctypes.  SURVEY.md section 7-1d / 8c pin 3: ~100 instances, compare
the KKT point (|dz| <= 1e-3) AND the objective (relative 1e-5), report mismatches instead of hiding them (the NLP is
non-convex: two solvers may, in principle, stop in different local minima).

Run here (needs /root/reference for oracle/_ref; ~2 min on 8 cores):   python tests/golden/make_slsqp_golden.py
The fixture stores, per instance: which workload / index, SLSQP's z [20,17], objective, iteration count, status and its
worst constraint violation -- and nothing from this repo's solvers.
"""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [("config2", i) for i in range(60)] + [("config3", i) for i in range(40)]


def batch_of(name):
    from forces_resilient_planner_b200 import workloads as W
    return W.config2(60) if name == "config2" else W.config3(40)


def work(case):
    import scipy_nlp
    from oracle import ref_model
    name, i = case
    b = batch_of(name)
    model = ref_model.RefModel("normal").eval
    z, res = scipy_nlp.solve_slsqp(b, i, model, maxiter=400)
    # worst violation of the equalities / corridor rows at SLSQP's point (its own feasibility, for the record)
    viol = float(np.max(np.abs(z[0, 8:17] - b.xinit[i])))
    for k in range(b.N - 1):
        p = np.zeros(130); p[0:10] = b.hdr[i, k]
        e = model(z[k], p, k)
        viol = max(viol, float(np.max(np.abs(e["c"] - np.concatenate([z[k + 1, 8:17], z[k + 1, 4:8]])))))
    return name, i, z, float(res.fun), int(res.nit), int(res.status), viol


if __name__ == "__main__":
    with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
        out = pool.map(work, CASES, chunksize=1)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "slsqp_kkt_points.npz"),
                        workload=np.array([o[0] for o in out]), index=np.array([o[1] for o in out], np.int32),
                        z=np.array([o[2] for o in out]), fun=np.array([o[3] for o in out]),
                        nit=np.array([o[4] for o in out], np.int32), status=np.array([o[5] for o in out], np.int32),
                        eq_violation=np.array([o[6] for o in out]))
    print("wrote", len(out), "instances; statuses", np.unique([o[5] for o in out], return_counts=True))
