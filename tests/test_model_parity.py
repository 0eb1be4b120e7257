"""Model layer: numpy restatement, C oracle (and, when present, the reference's own CasADi C)
against the committed golden vectors generated from the reference (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import model_np as M
from oracle import oracle as O
from oracle import ref_model

GOLD = os.path.join(os.path.dirname(__file__), "golden", "model_vectors.json")
CASES = json.load(open(GOLD))["cases"]
TOL = 1e-12   # relative to max(1, |value|); fp64 straight-line arithmetic


def _close(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)), initial=0.0)


def _jh(case):
    jh = np.zeros((30, 17))
    for i, j, v in case["jh_nnz"]:
        jh[i, j] = v
    return jh


@pytest.mark.parametrize("case", CASES, ids=[f'{c["variant"]}-{c["name"]}' for c in CASES])
def test_numpy_model_matches_reference_vectors(case):
    z, p, st = np.array(case["z"]), np.array(case["p"]), case["stage"]
    f, g, _ = M.objective(z, p, st, case["variant"])
    assert _close(f, case["f"]) < TOL and _close(g, case["grad"]) < TOL
    if st < 19:
        c, J = M.dynamics(z, p[3:6])
        assert _close(c, case["c"]) < TOL and _close(J, case["jc"]) < TOL
    h, Jh = M.corridor(z, p)
    assert _close(h, case["h"]) < TOL and _close(Jh, _jh(case)) < TOL


@pytest.mark.parametrize("case", CASES, ids=[f'{c["variant"]}-{c["name"]}' for c in CASES])
def test_c_oracle_model_matches_reference_vectors(case):
    z, p, st = np.array(case["z"]), np.array(case["p"]), case["stage"]
    r = O.model_eval(z, p, st, 20, 1 if case["variant"] == "final" else 0)
    assert _close(r["f"], case["f"]) < TOL and _close(r["grad"], case["grad"]) < TOL
    if st < 19:
        assert _close(r["c"], case["c"]) < TOL and _close(r["jc"], case["jc"]) < TOL
    assert _close(r["h"], case["h"]) < TOL and _close(r["jh"], _jh(case)) < TOL


def test_known_answer_vectors_of_the_survey():
    """SURVEY.md §8c KAT-generic numbers, typed in independently of make_golden.py."""
    c = next(c for c in CASES if c["name"] == "kat_generic" and c["variant"] == "normal")
    assert abs(c["f"] - 7.2832774485648955) < 1e-14
    assert np.allclose(c["c"][:3], [1.0251906119001011, 1.9834791118754478, 1.5103713446729301], atol=1e-15)
    jc = np.array(c["jc"])
    assert abs(jc[3, 3] - (-0.0014703703124166188)) < 1e-16
    assert abs(jc[3, 14] - 0.15095530665380721) < 1e-15
    assert abs((jc ** 2).sum() - 13.461604204128514) < 1e-12
    c0 = next(c for c in CASES if c["name"] == "kat_generic_stage0" and c["variant"] == "normal")
    assert abs(c0["f"] - 7.2972774485648957) < 1e-14
    hov = next(c for c in CASES if c["name"] == "kat_hover" and c["variant"] == "normal")
    assert hov["f"] == 0.0 and np.allclose(hov["grad"], 0) and abs(hov["c"][12] - 7.31157939) < 1e-12


def test_jacobian_against_finite_differences():
    rng = np.random.default_rng(5)
    for _ in range(10):
        z = M.LB + (M.UB - M.LB) * rng.random(17)
        fe = rng.uniform(-2, 2, 3)
        _, J = M.dynamics(z, fe)
        Jfd = np.zeros((13, 17))
        for i in range(17):
            dz = np.zeros(17); dz[i] = 1e-6
            Jfd[:, i] = (M.dynamics(z + dz, fe, jac=False) - M.dynamics(z - dz, fe, jac=False)) / 2e-6
        assert np.max(np.abs(J - Jfd)) < 1e-7


@pytest.mark.skipif(not ref_model.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_c_oracle_model_matches_live_reference_callbacks():
    rng = np.random.default_rng(11)
    for variant, vi in (("normal", 0), ("final", 1)):
        ref = ref_model.RefModel(variant)
        for _ in range(300):
            z = M.LB + (M.UB - M.LB) * rng.random(17)
            p = rng.normal(size=130); p[6:9] = rng.random(3) * 50
            st = int(rng.choice([0, 3, 18, 19]))
            a, b = ref.eval(z, p, st), O.model_eval(z, p, st, 20, vi)
            keys = ("f", "grad", "h", "jh") + (("c", "jc") if st < 19 else ())
            for k in keys:
                assert _close(b[k], a[k]) < TOL, (variant, st, k)
