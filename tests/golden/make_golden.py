"""Generate tests/golden/*.json.  Run in the BUILD container only (needs /root/reference).

  model_vectors.json   outputs of the REFERENCE's CasADi callbacks (oracle/_ref, compiled from
                       /root/reference/.../solver/{normal,final}/FORCESNLPsolver_*_casadi*.c) on
                       seeded random (z, p, stage) triples + the two known-answer vectors of
                       SURVEY.md §8c.  These pin the model layer of oracle/ and of the CUDA kernel.
  config1_solution.json  the anchor problem solved by the C oracle (NOT reference-derived: the
                       ForcesPro core cannot run here, exit -100); a regression pin only.

Usage:  make -C oracle all ref && python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import model_np as M  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle.ref_model import RefModel  # noqa: E402
from forces_resilient_planner_b200 import workloads as W  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def kat_inputs():
    mg = 0.745319 * 9.81
    z1 = np.array([0, 0, 0, mg, 0, 0, 0, mg, 1, 2, 1.5, 0, 0, 0, 0, 0, 0.0])
    p1 = np.zeros(130); p1[0:3] = [1, 2, 1.5]; p1[6:9] = [7, 1, 80]
    z2 = np.array([0.1, -0.2, 0.05, 7.5, 0.02, -0.01, 0.03, 7.3, 1, 2, 1.5, 0.5, -0.3, 0.2, 0.1, -0.05, 0.3])
    p2 = np.zeros(130); p2[0:10] = [1.2, 2.1, 1.4, 0.5, -0.2, 0.1, 7, 1, 80, 0.25]
    p2[10:19] = [1, 0, 0, -1, 0, 0, 0.6, 0.8, 0]; p2[100:103] = [3, 3, 2.5]
    return [("kat_hover", z1, p1, 1), ("kat_generic", z2, p2, 1), ("kat_generic_stage0", z2, p2, 0),
            ("kat_generic_stage19", z2, p2, 19)]


def main():
    rng = np.random.Generator(np.random.PCG64(20260101))
    cases = []
    for variant in ("normal", "final"):
        ref = RefModel(variant)
        todo = list(kat_inputs())
        for t in range(24):
            z = M.LB + (M.UB - M.LB) * rng.random(17)
            p = np.zeros(130)
            p[0:3] = rng.uniform(-5, 5, 3); p[3:6] = rng.uniform(-2, 2, 3)
            p[6:9] = [rng.uniform(1, 15), rng.uniform(0.2, 2), rng.uniform(10, 100)]
            p[9] = rng.uniform(-3, 3)
            m = int(rng.integers(0, 31))
            A = rng.normal(size=(m, 3)); A /= np.linalg.norm(A, axis=1, keepdims=True)
            p[10:10 + 3 * m] = A.reshape(-1); p[100:100 + m] = rng.uniform(0.5, 3, m)
            todo.append((f"rand{t}", z, p, int(rng.choice([0, 1, 7, 18, 19]))))
        for name, z, p, stage in todo:
            r = ref.eval(z, p, stage)
            cases.append(dict(name=name, variant=variant, stage=stage, z=z.tolist(), p=p.tolist(),
                              f=r["f"], grad=r["grad"].tolist(), c=r["c"].tolist(),
                              jc=r["jc"].tolist(), h=r["h"].tolist(),
                              jh_nnz=[[int(i), int(j), float(r["jh"][i, j])]
                                      for i, j in zip(*np.nonzero(r["jh"]))]))
    with open(os.path.join(HERE, "model_vectors.json"), "w") as fh:
        json.dump(dict(source="reference CasADi C via oracle/_ref (see make_golden.py)", cases=cases), fh)
    b = W.config1()
    r = O.solve_batch(b)
    with open(os.path.join(HERE, "config1_solution.json"), "w") as fh:
        json.dump(dict(source="oracle/nmpc_oracle.c fp64 (NOT reference-derived; regression pin)",
                       flag=int(r["flag"][0]), it=int(r["it"][0]), pobj=float(r["info_real"][0, 4]),
                       z=r["z"][0].tolist()), fh)
    print("wrote", len(cases), "model cases; config1 it =", int(r["it"][0]))


if __name__ == "__main__":
    main()
