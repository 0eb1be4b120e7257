"""CPU oracle (oracle/nmpc_oracle.c): converges on the synthetic configs, its KKT points pass
ForcesPro's acceptance test when re-evaluated with the reference's own callbacks, agree with an
independent scipy solve, and its Schur-complement KKT solve agrees with a dense numpy solve."""
import json
import os

import numpy as np
import pytest

import helpers as H
import scipy_nlp
from forces_resilient_planner_b200 import workloads as W
from oracle import model_np as M
from oracle import oracle as O
from oracle import ref_model

TOL = 1e-4   # codeoptions.nlp.Tol{Stat,Eq,Ineq,Comp} (mpc_generator_normal.m:76-79)


def _model(variant=0):
    """Prefer the reference's own callbacks; fall back to the (golden-pinned) C restatement."""
    if ref_model.available():
        return ref_model.RefModel("final" if variant else "normal").eval
    return lambda z, p, k: O.model_eval(z, p, k, 20, variant)


@pytest.mark.parametrize("variant", [0, 1])
def test_config1_anchor(variant):
    b = W.config1(variant=variant)
    r = O.solve_batch(b, multipliers=True)
    assert r["flag"][0] == 1 and r["it"][0] <= 12
    rs, req, rin, rc = H.kkt_residuals(b, 0, r["z"][0], r["y"][0], r["zl"][0], r["zu"][0], r["lc"][0], _model(variant))
    assert max(rs, req, rin, rc) <= TOL
    z = r["z"][0]
    assert np.allclose(z[0, 8:17], b.xinit[0])
    assert z[-1, 8] > (0.1 if variant else 0.5)   # it actually flies along +x towards the moving reference
    assert np.all(z >= M.LB - 1e-9) and np.all(z <= M.UB + 1e-9)


def test_config1_matches_committed_solution():
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "config1_solution.json")))
    r = O.solve_batch(W.config1())
    assert r["it"][0] == gold["it"] and r["flag"][0] == gold["flag"]
    assert np.max(np.abs(r["z"][0] - np.array(gold["z"]))) < 1e-9


@pytest.mark.parametrize("maker,kw", [(W.config2, dict(B=96)), (W.config3, dict(B=96)),
                                      (W.config2, dict(B=48, variant=1))])
def test_random_configs_converge_and_pass_reference_kkt(maker, kw):
    b = maker(**kw)
    r = O.solve_batch(b, multipliers=True)
    assert np.all(r["flag"] == 1)
    assert r["it"].max() <= 40
    model = _model(b.variant)
    for i in range(0, b.B, 8):
        res = H.kkt_residuals(b, i, r["z"][i], r["y"][i], r["zl"][i], r["zu"][i], r["lc"][i], model)
        assert max(res) <= TOL, (i, res)
        # the solver's self-reported residuals are the same quantities
        assert abs(res[0] - r["info_real"][i, 2]) < 1e-6 and abs(res[1] - r["info_real"][i, 0]) < 1e-7


@pytest.mark.parametrize("maker,kw", [(W.config2, dict(B=96)), (W.config3, dict(B=96)),
                                      (W.config2, dict(B=48, variant=1)), (W.config4, dict(side=8, n_stages=40))])
def test_predictor_corrector_option_reaches_the_same_kkt_points(maker, kw):
    """opts.pc = 1 (Mehrotra predictor-corrector): every point still passes ForcesPro's acceptance test when
    re-evaluated with the reference callbacks, it is the point the default algorithm finds, and it takes
    markedly fewer iterations."""
    b = maker(**kw)
    r = O.solve_batch(b, opts=O.default_opts(pc=1, mu0=10.0), multipliers=True)
    base = O.solve_batch(b)
    assert np.all(r["flag"] == 1) and np.all(base["flag"] == 1)
    assert r["it"].mean() < 0.8 * base["it"].mean() and r["it"].max() <= base["it"].max() + 4
    assert np.max(np.abs(r["z"] - base["z"])) < 5e-3
    if b.N == 20:
        model = _model(b.variant)
        for i in range(0, b.B, 8):
            res = H.kkt_residuals(b, i, r["z"][i], r["y"][i], r["zl"][i], r["zu"][i], r["lc"][i], model)
            assert max(res) <= TOL, (i, res)


def test_long_horizon_config4_converges():
    b = W.config4(8, 40)
    r = O.solve_batch(b)
    assert np.all(r["flag"] == 1)


@pytest.mark.parametrize("which", ["config1", "config2"])
def test_same_kkt_point_as_scipy_slsqp(which):
    b, i = (W.config1(), 0) if which == "config1" else (W.config2(4), 1)
    r = O.solve_batch(b)
    model = _model(0)
    z, res = scipy_nlp.solve_slsqp(b, i, model)
    assert np.max(np.abs(z - r["z"][i])) < 1e-3                      # SURVEY.md §8c pin 3
    f_ipm = H.total_cost(b, i, r["z"][i], model)
    assert abs(res.fun - f_ipm) / abs(res.fun) < 1e-5


def test_schur_kkt_solve_matches_dense_numpy():
    rng = np.random.default_rng(3)
    N = 20
    Phi = np.zeros((N, 17, 17)); g = rng.normal(size=(N, 17))
    C = np.zeros((N - 1, 13, 17)); d = rng.normal(size=(N - 1, 13)) * 0.1
    for k in range(N):
        A = rng.normal(size=(17, 17))
        Phi[k] = A @ A.T + 17 * np.eye(17)
        if k < N - 1:
            _, C[k] = M.dynamics(M.LB + (M.UB - M.LB) * rng.random(17), rng.normal(size=3))
    rc, dz, y = O.kkt_solve(Phi, g, C, d)
    assert rc == 0
    # dense reference: free variables = all but stage-0 states
    nz = N * 17
    free = np.ones(nz, bool); free[8:17] = False
    H_ = np.zeros((nz, nz)); A_ = np.zeros((13 * (N - 1), nz))
    for k in range(N):
        H_[k * 17:(k + 1) * 17, k * 17:(k + 1) * 17] = Phi[k]
    for k in range(N - 1):
        r0 = 13 * k
        A_[r0:r0 + 13, k * 17:(k + 1) * 17] = C[k]
        A_[r0:r0 + 9, (k + 1) * 17 + 8:(k + 1) * 17 + 17] -= np.eye(9)
        A_[r0 + 9:r0 + 13, (k + 1) * 17 + 4:(k + 1) * 17 + 8] -= np.eye(4)
    Hf = H_[np.ix_(free, free)]; Af = A_[:, free]
    K = np.block([[Hf, Af.T], [Af, np.zeros((Af.shape[0],) * 2)]])
    sol = np.linalg.solve(K, np.concatenate([-g.reshape(-1)[free], -d.reshape(-1)]))
    dz_ref = np.zeros(nz); dz_ref[free] = sol[:free.sum()]
    assert np.max(np.abs(dz.reshape(-1) - dz_ref)) < 1e-9
    assert np.max(np.abs(y[1:].reshape(-1) - sol[free.sum():])) < 1e-8


def test_infeasible_problem_reports_failure_not_success():
    b = W.config2(4)
    # wall 5 cm in front of a vehicle moving at 1.5 m/s towards it: no feasible trajectory
    b.xinit[:, 3:6] = [1.5, 0, 0]; b.z0[:, :, 11:14] = [1.5, 0, 0]
    b.rows[:, :, 0, 0:3] = [1.0, 0, 0]
    b.rows[:, :, 0, 3] = b.xinit[:, None, 0] + 0.05
    r = O.solve_batch(b, opts=O.default_opts(maxit=60))
    assert np.all(r["flag"] != 1)


def test_nan_input_is_flagged():
    b = W.config2(2)
    b.hdr[0, 3, 0] = np.nan
    r = O.solve_batch(b)
    assert r["flag"][0] in (-6, -7) and r["flag"][1] == 1
