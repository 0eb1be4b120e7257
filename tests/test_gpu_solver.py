"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle, the committed
golden fixtures and size-independent properties.  All need a GPU."""
import ctypes
import json
import os

import numpy as np
import pytest

import helpers as H
from forces_resilient_planner_b200 import _lib, forces, solver as S, workloads as W
from oracle import oracle as O
from oracle import prep_np as PN
from oracle import ref_model

pytestmark = pytest.mark.gpu
TOL = 1e-4
GOLD_DIR = os.path.join(os.path.dirname(__file__), "golden")


def _model(variant=0):
    if ref_model.available():
        return ref_model.RefModel("final" if variant else "normal").eval
    return lambda z, p, k: O.model_eval(z, p, k, 20, variant)


def _compare(g, c, frac_same_it=0.99, tol_z=1e-4, tol_z_typical=1e-9):
    """GPU (Riccati) vs oracle (Schur complement + refinement): same exit flags, same iterates.

    fp64 tolerance, stated: the two factorizations round differently, yet the iterate paths stay
    together to ~1e-13, so >= 99 % of the problems must agree to 1e-9 with identical iteration
    counts.  The remainder may take one iteration more or less when a residual lands next to its
    1e-4 threshold; those points then differ by at most the stopping tolerance (1e-4)."""
    assert np.array_equal(g.flag, c["flag"])
    dz = np.abs(g.z - c["z"]).reshape(g.z.shape[0], -1).max(1)
    assert dz.max() < tol_z and np.quantile(dz, 0.99) < tol_z_typical
    assert np.mean(g.it == c["it"]) >= frac_same_it and np.max(np.abs(g.it - c["it"])) <= 1


def test_device_model_matches_reference_vectors():
    cases = json.load(open(os.path.join(GOLD_DIR, "model_vectors.json")))["cases"]
    for variant in ("normal", "final"):
        cs = [c for c in cases if c["variant"] == variant]
        z = np.array([c["z"] for c in cs]); p = np.array([c["p"] for c in cs])
        st = np.array([c["stage"] for c in cs], np.int32)
        r = S.model_eval_device(z, p, st, 20, 1 if variant == "final" else 0)
        for i, c in enumerate(cs):
            rel = lambda a, b: np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(1, np.abs(np.asarray(b))), initial=0)
            assert rel(r["f"][i], c["f"]) < 1e-12 and rel(r["grad"][i], c["grad"]) < 1e-12
            if c["stage"] < 19:
                assert rel(r["c"][i], c["c"]) < 1e-12 and rel(r["jc"][i], c["jc"]) < 1e-12
            jh = np.zeros((30, 17))
            for a, b, v in c["jh_nnz"]:
                jh[a, b] = v
            assert rel(r["h"][i], c["h"]) < 1e-12 and rel(r["jh"][i], jh) < 1e-12


def test_config1_anchor_and_golden_solution():
    gold = json.load(open(os.path.join(GOLD_DIR, "config1_solution.json")))
    g = S.solve_host(W.config1())
    assert g.flag[0] == 1 and g.it[0] == gold["it"]
    assert np.max(np.abs(g.z[0] - np.array(gold["z"]))) < 1e-6
    assert abs(g.info_real[0, 4] - gold["pobj"]) < 1e-6


@pytest.mark.parametrize("maker,kw", [(W.config2, dict(B=512)), (W.config3, dict(B=512)),
                                      (W.config2, dict(B=128, variant=1)), (W.config4, dict(side=12, n_stages=40))])
def test_gpu_matches_oracle(maker, kw):
    b = maker(**kw)
    _compare(S.solve_host(b), O.solve_batch(b))


@pytest.mark.parametrize("maker,kw", [(W.config2, dict(B=512)), (W.config3, dict(B=512)),
                                      (W.config2, dict(B=128, variant=1)), (W.config4, dict(side=12, n_stages=40))])
def test_predictor_corrector_matches_oracle(maker, kw):
    """opts.pc = 1 (Mehrotra predictor-corrector through one factorisation), cold start with mu0 = 10: the same
    parity bar as the default algorithm, and fewer iterations than it."""
    b = maker(**kw)
    g = S.solve_host(b, opts=_lib.default_opts(pc=1, mu0=10.0))
    c = O.solve_batch(b, opts=O.default_opts(pc=1, mu0=10.0))
    _compare(g, c)
    base = O.solve_batch(b)
    assert np.all(base["flag"] == 1) and g.it.mean() < 0.8 * base["it"].mean()
    # same KKT point as the default algorithm, within the solver tolerance
    assert np.max(np.abs(g.z - base["z"])) < 5e-3


def test_device_pointer_api_equals_host_pointer_api():
    b = W.config2(300)
    a, c = S.solve(b), S.solve_host(b)
    assert np.array_equal(a.z, c.z) and np.array_equal(a.flag, c.flag) and np.array_equal(a.it, c.it)


def test_kkt_point_passes_forcespro_acceptance_with_reference_callbacks():
    for b in (W.config2(64), W.config3(64), W.config2(32, variant=1)):
        res, mult = S.solve_with_multipliers(b)
        assert np.all(res.flag == 1)
        model = _model(b.variant)
        for i in range(0, b.B, 4):
            r = H.kkt_residuals(b, i, res.z[i], mult["y"][i], mult["zl"][i], mult["zu"][i], mult["lc"][i], model)
            assert max(r) <= TOL, (i, r)
            assert abs(r[0] - res.info_real[i, 2]) < 1e-6      # self-reported rsnorm is honest


def test_full_size_config2_properties():
    """BASELINE config 2 at full size (B=4096): everything converges inside the iteration cap and
    the size-independent properties hold (feasibility, bounds, fixed initial state, idempotence)."""
    b = W.config2(4096)
    g = S.solve_host(b)
    assert np.all(g.flag == 1) and g.it.max() <= 40
    assert np.all(g.info_real[:, 0:4] <= TOL)
    assert np.array_equal(g.z[:, 0, 8:17], b.xinit)
    from oracle import model_np as M
    assert np.all(g.z >= M.LB - 1e-9) and np.all(g.z <= M.UB + 1e-9)
    viol = np.einsum("bkmj,bkj->bkm", b.rows[:, 1:, :, 0:3], g.z[:, 1:, 8:11]) - b.rows[:, 1:, :, 3] - 1e-5
    assert viol.max() <= TOL
    # u_prev chain: z_{k+1}[4:8] = z_k[0:4]
    assert np.max(np.abs(g.z[:, 1:, 4:8] - g.z[:, :-1, 0:4])) <= TOL
    # idempotence: restarting from the solution stays at the solution
    b2 = W.Batch(b.xinit, g.z.copy(), b.hdr, b.rows, b.nrows, b.variant)
    g2 = S.solve_host(b2)
    assert np.all(g2.flag == 1)
    assert np.max(np.abs(g2.z - g.z)) < 5e-3 and np.median(np.abs(g2.z - g.z).reshape(b.B, -1).max(1)) < 1e-4


# The reference symbols run one solve at a time: by default on the warp-group kernel (mixed precision at the reference
# tolerances: within 1e-3 of the fp64 KKT point, SURVEY 8c pin 4), with NMPC_B200_SHIM=fp64 on the one-warp fp64 kernel
# (then bit-identical to the batched fp64 entry points).
SHIM_MODES = [("fp64", 1e-12, 1e-9, 0), ("mixed", 1e-3, 1e-3, 1), (None, 1e-3, 1e-3, 1)]


def _set_shim(monkeypatch, mode):
    if mode is None:
        monkeypatch.delenv("NMPC_B200_SHIM", raising=False)
    else:
        monkeypatch.setenv("NMPC_B200_SHIM", mode)


@pytest.mark.parametrize("mode,tol,tol9,dit", SHIM_MODES)
def test_forces_abi_shim_matches_batch_api(monkeypatch, mode, tol, tol9, dit):
    """FORCESNLPsolver_normal_solve / _final_solve with the reference's padded 130-slot layout."""
    _set_shim(monkeypatch, mode)
    for variant, cls in ((0, forces.FORCESNormal), (1, forces.FORCESFinal)):
        b = W.config2(3, variant=variant)
        ref = S.solve_host(b)
        w = cls()
        for i in range(b.B):
            xinit, x0, allp = W.to_forces_params(b, i)
            w.params_.xinit[:] = xinit.tolist(); w.params_.x0[:] = x0.tolist()
            w.params_.all_parameters[:] = allp.tolist()
            flag = w.solve_params()
            assert flag == 1 == ref.flag[i]
            assert np.max(np.abs(w.output_array() - ref.z[i])) < tol
            assert abs(w.info_.it - ref.it[i]) <= dit and w.info_.solvetime > 0
            assert abs(w.info_.pobj - ref.info_real[i, 4]) < max(tol, 1e-12) * max(1.0, abs(ref.info_real[i, 4]))
            assert max(w.info_.rsnorm, w.info_.res_eq, w.info_.res_ineq, w.info_.rcompnorm) <= TOL


@pytest.mark.parametrize("mode,tol,tol9,dit", SHIM_MODES)
def test_reference_wrapper_flow_solveNormal_updateNormal(monkeypatch, mode, tol, tol9, dit):
    """The planner's own call sequence (nmpc_solver.cpp:384,421): setParas -> solve -> update."""
    _set_shim(monkeypatch, mode)
    b = W.config1()
    w = forces.FORCESNormal()
    w.setParasNormal(7.0, 1.0, 80.0, 12.0, 0.5)
    mpc_output = [np.array(b.z0[0, 0]) for _ in range(21)]            # initMPCOutput
    A = b.rows[0, 1, :6, 0:3]; braw = b.rows[0, 1, :6, 3] + np.linalg.norm(A * W.EGO_E, axis=1)
    E = np.diag(W.EGO_E)
    flag = w.solveNormal(mpc_output, [0.0, 0.0, 0.0], [b.hdr[0, k, 0:3] for k in range(20)],
                         [b.hdr[0, k, 9] for k in range(20)], [E] * 20, [(A, braw)], np.zeros(20))
    assert flag == 1
    w.updateNormal(mpc_output)
    direct = S.solve_host(b)
    assert np.max(np.abs(np.array(mpc_output[:20]) - direct.z[0])) < tol9


def test_edge_cases_rows_and_batch_sizes():
    assert S.solve_host(W.config2(0)).z.shape == (0, 20, 17)             # empty batch
    b = W.config2(5)
    b0 = W.Batch(b.xinit, b.z0, b.hdr, np.zeros((5, 20, 0, 4)), np.zeros((5, 20), np.int32))   # no corridor at all
    r0 = S.solve_host(b0); c0 = O.solve_batch(b0)
    assert np.all(r0.flag == 1) and np.max(np.abs(r0.z - c0["z"])) < 1e-4
    # ragged: a different live-row count at every stage, capacity 30 (the ABI maximum)
    rng = np.random.default_rng(0)
    b = W.config3(64, mcap=30)
    b.nrows[:] = np.minimum(b.nrows, rng.integers(0, 11, b.nrows.shape)).astype(np.int32)
    _compare(S.solve_host(b), O.solve_batch(b))
    # nrows larger than mcap is clamped, not read out of bounds
    b = W.config2(4); b.nrows[:] = 99
    c = W.config2(4); c.nrows[:] = c.mcap
    assert np.array_equal(S.solve_host(b).z, S.solve_host(c).z)


def test_failure_flags():
    b = W.config2(6)
    b.hdr[0, 3, 0] = np.nan                                   # NaN in a function evaluation -> -6
    b.xinit[1, 3:6] = [1.5, 0, 0]; b.z0[1, :, 11:14] = [1.5, 0, 0]
    b.rows[1, :, 0, 0:3] = [1.0, 0, 0]; b.rows[1, :, 0, 3] = b.xinit[1, 0] + 0.05   # wall: infeasible
    g = S.solve_host(b, opts=_lib.default_opts(maxit=60))
    c = O.solve_batch(b, opts=O.default_opts(maxit=60))
    assert g.flag[0] in (-6, -7) and g.flag[1] != 1
    assert np.all(g.flag[2:] == 1)                            # a failing instance never stalls the others
    assert np.array_equal(g.flag[2:], c["flag"][2:])


def test_bad_arguments_are_rejected_loudly():
    lib = _lib.load()
    b = W.config2(2)
    x = np.zeros(4)
    rc = lib.nmpc_solve_batch_host_f64(2, 21, 8, x.ctypes.data, x.ctypes.data, x.ctypes.data, x.ctypes.data,
                                       x.ctypes.data, 0, None, x.ctypes.data, x.ctypes.data, x.ctypes.data)
    assert rc == -11 and b"N=21" in lib.nmpc_last_error()
    with pytest.raises(RuntimeError):
        S.solve_host(W.Batch(b.xinit, b.z0, b.hdr, np.zeros((2, 20, 40, 4)), b.nrows))   # mcap > 30


def test_warm_start_shift_uses_fewer_iterations():
    b = W.config2(256)
    g = S.solve_host(b)
    xinit, z0 = W.shift_warm_start(g.z)
    hdr = b.hdr.copy()
    hdr[:, :-1, 0:3] = b.hdr[:, 1:, 0:3]; hdr[:, -1, 0:3] = 2 * b.hdr[:, -1, 0:3] - b.hdr[:, -2, 0:3]
    hdr[:, :-1, 9] = b.hdr[:, 1:, 9]
    b2 = W.Batch(xinit, z0, hdr, b.rows, b.nrows, b.variant)
    g2 = S.solve_host(b2, opts=_lib.default_opts(mu0=0.1))
    c2 = O.solve_batch(b2, opts=O.default_opts(mu0=0.1))
    assert np.all(g2.flag == 1) and g2.it.mean() < g.it.mean()
    _compare(g2, c2)


def test_cpp_dropin_demo_runs_the_planner_call_sequence():
    """host/dropin_demo.cpp: FORCESNormal/FORCESFinal (C++ mirror of forces_normal.cpp) against the .so."""
    import subprocess
    host = os.path.join(os.path.dirname(GOLD_DIR), "..", "forces_resilient_planner_b200", "host")
    subprocess.check_call(["make", "-C", host, "-s"])
    out = subprocess.run([os.path.join(host, "dropin_demo")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("exitflag 1") == 4
    # host/pipeline_demo.cpp: three closed-loop fleet replans (shift, ellipsoids, references, corridors, pack, solve)
    # through the device-pointer C ABI only, from C++
    out = subprocess.run([os.path.join(host, "pipeline_demo")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("64/64 accepted") == 3 and out.stdout.count("corridor overflow 0") == 3, out.stdout
    # host/sharded_demo.cpp: the multi-GPU host side from C++ -- one process per rank, CUDA IPC handle exchange, the fused
    # peer-store collation; every rank re-solves every shard and compares bit for bit (ranks share the GPU if there is one)
    out = subprocess.run([os.path.join(host, "sharded_demo"), "2", "32"], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0 and "0 of 2 slices differ" in out.stdout and "64 / 64 with exit flag 1" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("maker,kw,mixed", [(W.config2, dict(B=4096), False), (W.config3, dict(B=4096), False),
                                            (W.config2, dict(B=1024, variant=1), False),
                                            (W.config2, dict(B=4096), True), (W.config3, dict(B=4096), True),
                                            (W.config2, dict(B=1024, variant=1), True),
                                            (W.config2, dict(B=4096), "group"), (W.config3, dict(B=4096), "group"),
                                            (W.config2, dict(B=1024, variant=1), "group")])
def test_whole_batch_passes_forcespro_acceptance_with_reference_callbacks(maker, kw, mixed):
    """Tier bar #1 at BASELINE size: EVERY problem of config 2 (4096), config 3 (4096 of its 65536) and the final
    variant (1024) -- through the fp64 kernel, the mixed-precision kernel and the warp-group kernel (the one the
    reference symbols run on) -- is re-evaluated with the
    reference's own casadi2forces callbacks (oracle/_ref; oracle/kkt_check.c) at the returned point and multipliers:
    the four inf-norms ForcesPro stops on are <= 1e-4, every multiplier is non-negative (a sign-blind complementarity
    check would let a negative one through), and the residuals the kernel reports are the ones recomputed here."""
    b = maker(**kw)
    res, mult = S.solve_with_multipliers(b, mixed=bool(mixed), lowlatency=(mixed == "group"))
    assert np.all(res.flag == 1)
    chk, used_ref = H.kkt_residuals_batch(b, res.z, mult["y"], mult["zl"], mult["zu"], mult["lc"], variant=b.variant)
    assert used_ref == ref_model.available()
    assert chk[:, 0:4].max() <= TOL * (1 + 1e-9), chk[:, 0:4].max(0)
    assert chk[:, 4].min() >= 0.0, chk[:, 4].min()
    assert np.max(np.abs(chk[:, 0] - res.info_real[:, 2])) < 1e-6 and np.max(np.abs(chk[:, 1] - res.info_real[:, 0])) < 1e-7
    assert np.max(np.abs(chk[:, 5] - res.info_real[:, 4])) <= 1e-9 * np.max(np.abs(chk[:, 5]))     # pobj is the reference's cost


def test_long_horizon_config4_matches_oracle_at_4096_problems():
    """BASELINE config 4 (N = 40; no reference counterpart exists: N = 20 is baked into its ABI) on a 64 x 64 wind sweep =
    4096 problems: the fp64 kernel against the CPU port (same algorithm, dense Riccati in double precision -- oracle
    opts.mixed = 2, which itself walks the Schur restatement's path iteration for iteration), and the mixed-precision
    kernel against both."""
    b = W.config4(64, 40)
    c = O.solve_batch(b, opts=O.default_opts(mixed=2))
    assert np.all(c["flag"] == 1)
    _compare(S.solve_host(b), c, frac_same_it=0.98)
    m = S.solve_host(b, np.float32)
    assert np.all(m.flag == 1) and m.resolved.mean() < 0.02
    dz = np.abs(m.z.astype(np.float64) - c["z"]).reshape(b.B, -1).max(1)
    assert dz.max() < 2e-3 and np.median(dz) < 2e-5, (dz.max(), np.median(dz))
    assert abs(m.it.mean() - c["it"].mean()) < 0.03 * c["it"].mean()


def test_infeasible_initial_state_is_reported_not_solved():
    """xinit outside a stage-0 corridor row, or outside a bound: the stage-0 states are fixed by the xinit equality, so the
    reference's NLP has no feasible point.  Stated behaviour (include/nmpc_b200.h): NOPROGRESS (-7) after zero iterations
    with the violation in res_ineq -- from the fp64 kernel, the mixed-precision kernel, the reference-named symbol and the
    CPU port alike; neighbours in the same batch are unaffected, and a violation within TolIneq is not an error."""
    b = W.config2(8)
    base = S.solve_host(b)
    bad = W.Batch(b.xinit.copy(), b.z0.copy(), b.hdr.copy(), b.rows.copy(), b.nrows.copy(), b.variant)
    # problem 2: 0.3 m beyond row 0 of its stage-0 polytope; problem 5: vertical speed beyond the 2 m/s bound
    a0 = bad.rows[2, 0, 0, 0:3]
    bad.xinit[2, 0:3] += a0 * (bad.rows[2, 0, 0, 3] - a0 @ bad.xinit[2, 0:3] + 0.3)
    bad.z0[2, :, 8:11] = bad.xinit[2, 0:3]
    bad.xinit[5, 5] = 2.5; bad.z0[5, :, 13] = 2.5
    # problem 6: at rest, 5e-5 beyond the BACK plane of its box (row 3, two metres behind the start of the reference):
    # within TolIneq, so still a valid problem -- and a solvable one, the reference pulls it inside
    a6 = bad.rows[6, 0, 3, 0:3]
    bad.xinit[6, 0:3] += a6 * (bad.rows[6, 0, 3, 3] + 1e-5 - a6 @ bad.xinit[6, 0:3] + 5e-5)
    bad.xinit[6, 3:6] = 0.0
    bad.z0[6, :, 8:17] = bad.xinit[6]
    for r in (S.solve_host(bad), S.solve_host(bad, mixed=True), S.solve_host(bad, np.float32)):
        assert list(r.flag[[2, 5]]) == [-7, -7] and list(r.it[[2, 5]]) == [0, 0]
        assert abs(r.info_real[2, 1] - (0.3 - 1e-5)) < 1e-6 and abs(r.info_real[5, 1] - 0.5) < 1e-6
        keep = [0, 1, 3, 4, 7]
        assert np.all(r.flag[keep] == 1) and r.flag[6] == 1
        if r.z.dtype == np.float64:
            assert np.max(np.abs(r.z[keep] - base.z[keep])) < 1e-3
    c = O.solve_batch(bad)
    assert list(c["flag"][[2, 5]]) == [-7, -7] and c["flag"][6] == 1
    w = forces.FORCESNormal()
    xinit, x0, allp = W.to_forces_params(bad, 2)
    w.params_.xinit[:] = xinit.tolist(); w.params_.x0[:] = x0.tolist(); w.params_.all_parameters[:] = allp.tolist()
    assert w.solve_params() == -7 and w.info_.it == 0 and abs(w.info_.res_ineq - (0.3 - 1e-5)) < 1e-6


def test_kkt_points_match_independent_slsqp_solutions():
    """SURVEY.md 7-1d / 8c pin 3 on the device: 100 instances solved by scipy SLSQP driving the reference's callbacks
    (tests/golden/slsqp_kkt_points.npz, made by tests/golden/make_slsqp_golden.py) against the fp64 kernel, the
    mixed-precision kernel and the warp-group kernel.  Every instance is reported; see test_oracle_solver.py for the same check of the CPU port."""
    g = np.load(os.path.join(GOLD_DIR, "slsqp_kkt_points.npz"))
    for name, maker in (("config2", lambda: W.config2(60)), ("config3", lambda: W.config3(40))):
        sel = g["workload"] == name
        b = maker()
        for r in (S.solve_host(b), S.solve_host(b, mixed=True), S.solve(b, lowlatency=True)):
            rep = H.compare_with_slsqp(b, r.z, r.flag, g["index"][sel], g["z"][sel], g["fun"][sel])
            assert rep["n"] == int(sel.sum()) and rep["n_same_point"] >= rep["n"] - rep["n_slsqp_worse"], rep
            assert rep["n_ours_worse"] == 0, rep


MIXED_CASES = [(W.config2, dict(B=512)), (W.config3, dict(B=1024)), (W.config2, dict(B=128, variant=1)),
               (W.config4, dict(side=16, n_stages=40))]


@pytest.mark.parametrize("maker,kw", MIXED_CASES)
def test_mixed_precision_meets_the_reference_tolerances(maker, kw):
    """BASELINE configs 3/4 name fp32.  The *_f32 entry points run the mixed-precision kernel (single-precision
    Newton system in delta form, double-precision iterate / residuals / line search) with nmpc_default_opts
    UNCHANGED: the reference's 1e-4 stopping test (mpc_generator_normal.m:76-79).  Stated tolerance: every problem
    ends with exit flag 1, |z - z_fp64_oracle|_inf <= 1e-3 (SURVEY.md 8c pin 4; measured ~5e-4 worst, 1e-6 median --
    both solvers stop anywhere inside the 1e-4 residual ball), iteration counts as the fp64 solver's."""
    b = maker(**kw)
    c = O.solve_batch(b)
    assert np.all(c["flag"] == 1)
    for res in (S.solve_host(b, np.float32), S.solve_host(b, np.float64, mixed=True)):
        assert np.all(res.flag == 1), np.unique(res.flag, return_counts=True)
        assert np.all(res.info_real[:, 0:4] <= TOL * (1 + 1e-6))
        dz = np.abs(res.z.astype(np.float64) - c["z"]).reshape(b.B, -1).max(1)
        assert dz.max() < 1e-3 and np.median(dz) < 2e-5, (dz.max(), np.median(dz))
        assert abs(res.it.mean() - c["it"].mean()) < 0.05 * c["it"].mean()
        assert res.resolved.mean() <= 0.05          # the fp64 safety net is the exception, not the path
    # the CPU restatement of the same algorithm (oracle opts.mixed = 1: dense single-precision Riccati) walks the same path
    m = O.solve_batch(b, opts=O.default_opts(mixed=1))
    same = (m["flag"] == 1) & (res.resolved == 0)
    assert np.mean(np.abs(res.it[same] - m["it"][same]) <= 1) >= 0.97


@pytest.mark.parametrize("maker,kw", [(W.config2, dict(B=300)), (W.config3, dict(B=300)), (W.config2, dict(B=64, variant=1)),
                                      (W.config4, dict(side=12, n_stages=40)), (W.config1, dict())])
def test_warp_group_kernel_is_the_mixed_kernel_spread_over_256_threads(maker, kw):
    """nmpc_solve_batch_lowlatency_f64 (one warp-group per problem) against the one-warp mixed-precision kernel and the fp64
    CPU port: same exit flags, same KKT points (both stop inside the 1e-4 residual ball: |dz| <= 1e-3 vs fp64), iteration
    counts equal to the one-warp kernel's on >= 97 % of the problems (the two sweeps round differently), and every point
    passes ForcesPro's acceptance test with the reference callbacks."""
    b = maker(**kw)
    c = O.solve_batch(b)
    w = S.solve(b, mixed=True)
    g, mult = S.solve_with_multipliers(b, lowlatency=True)
    assert np.array_equal(g.flag, c["flag"]) and np.all(g.flag == 1)
    dz = np.abs(g.z - c["z"]).reshape(b.B, -1).max(1)
    assert dz.max() < 1e-3 and np.median(dz) < 2e-5
    assert np.mean(np.abs(g.it - w.it) <= 1) >= 0.97 and abs(g.it.mean() - c["it"].mean()) < 0.05 * c["it"].mean()
    assert np.all(g.info_real[:, 0:4] <= TOL)
    if b.N == 20:
        chk, _ = H.kkt_residuals_batch(b, g.z, mult["y"], mult["zl"], mult["zu"], mult["lc"], variant=b.variant)
        assert chk[:, 0:4].max() <= TOL * (1 + 1e-9) and chk[:, 4].min() >= 0.0


def test_mixed_precision_kkt_points_pass_forcespro_acceptance_with_reference_callbacks():
    for b in (W.config3(96), W.config2(48, variant=1)):
        res, mult = S.solve_with_multipliers(b, mixed=True)
        assert np.all(res.flag == 1)
        model = _model(b.variant)
        for i in range(0, b.B, 4):
            r = H.kkt_residuals(b, i, res.z[i], mult["y"][i], mult["zl"][i], mult["zu"][i], mult["lc"][i], model)
            assert max(r) <= TOL, (i, r)
            assert abs(r[0] - res.info_real[i, 2]) < 1e-6      # the fp64 residual the kernel stops on is honest
            assert min(mult["zl"][i].min(), mult["zu"][i].min(), mult["lc"][i].min()) >= 0.0


def test_mixed_precision_hands_unsolved_problems_to_the_fp64_kernel():
    """The safety net, made deterministic: with an iteration cap of 3 the mixed kernel leaves every problem at exit
    flag 0, so ALL of them are re-solved by the fp64 kernel (reading / writing the caller's float or double arrays, on the
    same stream).  A problem given up at -5 / 0 restarts from the mixed kernel's last iterate with a small barrier
    (mu0 = 0.1): the results must be exactly what the fp64 kernel produces from that start, and carry the re-solve mark.
    opts.mixed = -1 switches the net off."""
    b = W.config3(300)
    o = _lib.default_opts(maxit=3)
    off = S.solve_host(b, np.float64, opts=_lib.default_opts(maxit=3, mixed=-1), mixed=True)      # the mixed kernel alone
    assert np.all(off.resolved == 0) and np.all(off.flag == 0) and np.all(off.it == 3)
    ref0 = S.solve_host(b, np.float64, opts=o)
    assert np.max(np.abs(off.z - ref0.z)) < 1e-3         # three iterations of the same algorithm in mixed precision
    m = S.solve_host(b, np.float64, opts=o, mixed=True)
    restart = W.Batch(b.xinit, off.z.copy(), b.hdr, b.rows, b.nrows, b.variant)
    ref = S.solve_host(restart, np.float64, opts=_lib.default_opts(maxit=3, mu0=0.1))
    assert np.all(m.resolved == 1) and np.array_equal(m.z, ref.z) and np.array_equal(m.flag, ref.flag) and np.array_equal(m.it, ref.it)
    # float arrays: the fp64 kernel reads and writes them directly (io32)
    off32 = S.solve_host(b, np.float32, opts=_lib.default_opts(maxit=3, mixed=-1))
    b32 = b.astype(np.float32).astype(np.float64)
    restart32 = W.Batch(b32.xinit, off32.z.astype(np.float64), b32.hdr, b32.rows, b32.nrows, b32.variant)
    ref32 = S.solve_host(restart32, np.float64, opts=_lib.default_opts(maxit=3, mu0=0.1))
    m32 = S.solve_host(b, np.float32, opts=o)
    assert np.all(m32.resolved == 1) and np.array_equal(m32.z, ref32.z.astype(np.float32))
    # NaN input: nothing to restart from -- the re-solve starts from the caller's guess and reports the bad input itself
    bad = W.config3(8)
    bad.hdr[2, 5, 1] = np.nan
    r = S.solve_host(bad, np.float64, mixed=True)
    assert r.flag[2] in (-6, -7) and np.all(np.delete(r.flag, 2) == 1)


def test_receding_horizon_stream_matches_cpu_closed_loop():
    """Config 5 in miniature: 48 agents x 12 replans, everything device-resident (shift + pack + solve,
    CUDA-graph replay from cycle 2), against the same closed loop run with the CPU oracle and the
    numpy restatement of the packing loop."""
    from forces_resilient_planner_b200 import stream as ST, prep
    b = W.config2(48)
    rng_a, rng_b = (np.random.Generator(np.random.PCG64(9)) for _ in range(2))
    s = ST.RecedingHorizonStream(b, use_graph=True)
    ext_g = b.hdr[:, 0, 3:6].copy(); ext_c = ext_g.copy()
    A = b.rows[:, 1, :, 0:3][:, None]; braw = (b.rows[:, 1, :, 3] + np.linalg.norm(b.rows[:, 1, :, 0:3] * W.EGO_E, axis=-1))[:, None]
    pm = b.nrows[:, 1:2].astype(np.int32); pidx = np.zeros((b.B, b.N), np.int32)
    E = np.tile(np.diag(W.EGO_E).reshape(1, 1, 9), (b.B, b.N, 1))
    xinit, z0 = b.xinit.copy(), b.z0.copy()
    for step in range(12):
        ref, yaw, ext_g = ST.synthetic_refs(b, step, rng_a, ext_g)
        cmd, flag, it = s.replan(ref, yaw, ext_g)
        # CPU twin of the same cycle
        ref_c, yaw_c, ext_c = ST.synthetic_refs(b, step, rng_b, ext_c)
        hdr, rows, nrows = PN.pack_params_reference(ref_c, yaw_c, ext_c, E, A, braw, pm, pidx, s.weights, b.mcap)
        cb = W.Batch(xinit, z0, hdr, rows, nrows, 0)
        c = O.solve_batch(cb, opts=O.default_opts(mu0=1.0 if step == 0 else 0.1))
        assert np.all(flag == 1) and np.all(c["flag"] == 1), step
        assert np.array_equal(it, c["it"]), step
        assert np.max(np.abs(cmd - c["z"][:, 0, 0:4])) < 1e-7, step
        xinit, z0 = W.shift_warm_start(c["z"])
    assert s.graph is not None


def test_failed_agent_leaves_nothing_behind_and_restarts_cold():
    """Result handling of solveNMPC (nmpc_solver.cpp:398-427, 363-364), per agent: a failed solve (here: NaN
    references for one agent during one cycle -> NaN output, exit -6) is not adopted; that agent restarts from the
    cold guess at its state on the next cycle, the others are untouched -- same commands as a fleet without the fault."""
    from forces_resilient_planner_b200 import stream as ST
    b = W.config2(24)
    runs = []
    for inject in (False, True):
        rng = np.random.Generator(np.random.PCG64(5))
        s = ST.RecedingHorizonStream(b, use_graph=True)
        ext = b.hdr[:, 0, 3:6].copy()
        hist = []
        for step in range(8):
            ref, yaw, ext = ST.synthetic_refs(b, step, rng, ext)
            if inject and step == 4:
                ref = ref.copy(); ref[3] = np.nan
            cmd, flag, it = s.replan(ref, yaw, ext)
            hist.append((cmd.copy(), flag.copy(), it.copy()))
            if inject and step == 4:
                assert flag[3] in (-6, -7) and np.all(np.delete(flag, 3) == 1)
                bad_out = s.z[3].cpu().numpy().copy()                          # whatever the failed solve left behind ...
            if inject and step == 5:
                zp = s.zprev.cpu().numpy()
                assert np.all(np.isfinite(zp))                                 # ... never entered the plan in force:
                cold = np.zeros(17); cold[3] = cold[7] = 7.3                   # agent 3 holds the cold guess at its state
                assert np.array_equal(zp[3, :, 0:8], np.tile(cold[0:8], (b.N, 1))) and not np.array_equal(zp[3], bad_out)
                assert flag[3] == 1 and it[3] > hist[3][2][3]                  # cold restart: more iterations than a warm one
        runs.append(hist)
    clean, faulty = runs
    for step in range(8):
        keep = np.arange(b.B) != 3
        assert np.array_equal(clean[step][1][keep], faulty[step][1][keep])
        assert np.max(np.abs(clean[step][0][keep] - faulty[step][0][keep])) == 0.0, step
    assert np.all(faulty[7][1] == 1) and np.all(np.isfinite(faulty[7][0]))
    # the restarted agent is back on the fleet's track two cycles later (same problem, different start of the iteration)
    assert np.max(np.abs(clean[7][0][3] - faulty[7][0][3])) < 0.5


def test_rank_longest_first_equals_stable_argsort():
    import torch
    from forces_resilient_planner_b200 import prep
    rng = np.random.default_rng(2)
    for B in (1, 37, 1024, 5000):
        ii = np.zeros((B, 4), np.int32)
        ii[:, 0] = rng.choice([1, 1, 1, 1, 0, -5, -7], B); ii[:, 1] = rng.integers(3, 30, B)
        key = ii[:, 1] + 1000 * (ii[:, 0] != 1)
        want = np.argsort(-key, kind="stable").astype(np.int32)
        got = prep.rank_longest_first(torch.from_numpy(ii).cuda(), torch.empty(B, dtype=torch.int32, device="cuda")).cpu().numpy()
        assert np.array_equal(got, want), B


@pytest.mark.parametrize("B,expect_group", [(40, True), (200, True), (400, False)])
def test_mixed_precision_stream_follows_the_fp64_closed_loop(B, expect_group):
    """The receding-horizon stream through the mixed-precision kernels (warp-group kernel for a small fleet -- its
    one-CTA-per-SM build at 40 agents, its two-CTAs-per-SM build at 200 --, one-warp kernel for a large one) against the
    same stream through the fp64 kernel: every replan converges, the commands agree
    to the solver tolerance, the iteration counts agree on >= 95 % of the (agent, replan) pairs."""
    from forces_resilient_planner_b200 import stream as ST
    b = W.config2(B)
    runs = []
    for mixed in (False, True):
        rng = np.random.Generator(np.random.PCG64(17))
        s = ST.RecedingHorizonStream(b, use_graph=True, mixed=mixed)
        assert s.lowlatency == (mixed and expect_group)
        ext = b.hdr[:, 0, 3:6].copy()
        hist = []
        for step in range(10):
            ref, yaw, ext = ST.synthetic_refs(b, step, rng, ext)
            hist.append(s.replan(ref, yaw, ext))
        runs.append(hist)
    same = []
    for (c0, f0, i0), (c1, f1, i1) in zip(*runs):
        assert np.all(f0 == 1) and np.all(f1 == 1)
        assert np.max(np.abs(c0 - c1)) < 5e-3
        same.append(np.mean(i0 == i1))
    assert np.mean(same) >= 0.95


def test_stream_with_propagated_ellipsoids_matches_cpu_closed_loop():
    """The same loop with the corridor tightened by the disturbance ellipsoids propagated along the
    previous plan on the device, against the literal CPU restatement."""
    from forces_resilient_planner_b200 import stream as ST
    from oracle import ellipsoid_np as EN
    b = W.config2(6)
    rng_a, rng_b = (np.random.Generator(np.random.PCG64(11)) for _ in range(2))
    s = ST.RecedingHorizonStream(b, use_graph=True, dynamic_ellipsoids=True)
    ext_g = b.hdr[:, 0, 3:6].copy(); ext_c = ext_g.copy()
    A = b.rows[:, 1, :, 0:3][:, None]; braw = (b.rows[:, 1, :, 3] + np.linalg.norm(b.rows[:, 1, :, 0:3] * W.EGO_E, axis=-1))[:, None]
    pm = b.nrows[:, 1:2].astype(np.int32); pidx = np.zeros((b.B, b.N), np.int32)
    xinit, z0 = b.xinit.copy(), b.z0.copy()
    prev = z0
    for step in range(5):
        ref, yaw, ext_g = ST.synthetic_refs(b, step, rng_a, ext_g)
        cmd, flag, it = s.replan(ref, yaw, ext_g)
        ref_c, yaw_c, ext_c = ST.synthetic_refs(b, step, rng_b, ext_c)
        E = EN.propagate_batch(prev).reshape(b.B, b.N, 9)
        hdr, rows, nrows = PN.pack_params_reference(ref_c, yaw_c, ext_c, E, A, braw, pm, pidx, s.weights, b.mcap)
        assert np.max(np.abs(s.rows.cpu().numpy() - rows)) < 1e-10, step
        c = O.solve_batch(W.Batch(xinit, z0, hdr, rows, nrows, 0), opts=O.default_opts(mu0=1.0 if step == 0 else 0.1))
        assert np.all(flag == 1) and np.all(c["flag"] == 1), step
        assert np.array_equal(it, c["it"]), step
        assert np.max(np.abs(cmd - c["z"][:, 0, 0:4])) < 1e-7, step
        prev = c["z"]
        xinit, z0 = W.shift_warm_start(c["z"])
    # the later stages are tightened more than the static ego ellipsoid would
    static = b.rows[:, 1:, :6, 3]
    assert np.all(rows[:, 5:, :6, 3] < static[:, 4:] - 0.05)


def test_planner_pipeline_matches_cpu_twin():
    """setFORCESParams + solveNormal for a few agents, every step on the device (shift, ellipsoids, references,
    corridors from an obstacle cloud, packing, solve), against the same replans assembled from the CPU
    restatements (oracle/prep_np.py, ellipsoid_np.py, corridor_np.py) and solved by the CPU oracle."""
    from forces_resilient_planner_b200 import stream as ST
    from oracle import corridor_np as CN, ellipsoid_np as EN
    rng = np.random.default_rng(21)
    B, N, P, Ts, mcap = 5, 20, 60, 0.05, 30
    # front end: a gently curving 1 m/s polyline per agent through clutter with a clear tube, sampled at Ts
    k = np.arange(P)
    paths = np.zeros((B, P, 3)); clouds = np.zeros((B, 400, 3)); cn = np.zeros(B, np.int32)
    for a in range(B):
        head = rng.uniform(-np.pi, np.pi)
        ang = head + 0.25 * np.sin(0.08 * k + rng.uniform(0, 6))
        paths[a, :, 0] = np.cumsum(0.05 * np.cos(ang)); paths[a, :, 1] = np.cumsum(0.05 * np.sin(ang)); paths[a, :, 2] = 1.0 + 0.1 * a
        pts = rng.uniform(paths[a].min(0) - [2.5, 2.5, 1.0], paths[a].max(0) + [2.5, 2.5, 1.0], (400, 3))
        keep = pts[np.min(np.linalg.norm(pts[:, None] - paths[a][None], axis=2), axis=1) > 0.9]
        cn[a] = len(keep); clouds[a, :len(keep)] = keep
    size = np.full(B, P, np.int32)
    xinit = np.zeros((B, 9)); xinit[:, 0:3] = paths[:, 0]; xinit[:, 8] = np.arctan2(paths[:, 5, 1] - paths[:, 0, 1], paths[:, 5, 0] - paths[:, 0, 0])
    pipe = ST.PlannerPipeline(xinit, paths, size, clouds, cn, mcap=mcap, max_polys=20)
    w5 = pipe.weights
    prev = W.cold_start(xinit, N); x_c, z0_c = xinit.copy(), prev.copy()
    ext = rng.uniform(-0.5, 0.5, (B, 3))
    for cyc in range(4):
        t_off = np.full(B, cyc * Ts) + rng.uniform(0, 0.01, B)
        cmd, flag, it = pipe.replan(ext, t_off)
        # ---- CPU twin ----
        if cyc > 0:
            x_c, z0_c = W.shift_warm_start(prev)
        E = EN.propagate_batch(prev)
        rp, ry, far = PN.sample_reference(paths, size, t_off, prev[:, 1, 16], N, Ts, pos1=prev[:, 1, 8:11])
        pA = np.zeros((B, 20, mcap, 3)); pb = np.zeros((B, 20, mcap)); pm = np.zeros((B, 20), np.int32); pidx = np.zeros((B, N), np.int32)
        for a in range(B):
            polys, idx = CN.select_corridors(rp[a], ry[a], E[a], clouds[a, :cn[a]])
            pidx[a] = idx
            for q, (A, b) in enumerate(polys):
                m = min(len(b), mcap); pA[a, q, :m] = A[:m]; pb[a, q, :m] = b[:m]; pm[a, q] = m
        hdr, rows, nrows = PN.pack_params_reference(rp, ry, ext, E.reshape(B, N, 9), pA, pb, pm, pidx, w5, mcap)
        L = pipe.last
        assert np.max(np.abs(L["ref_pos"].cpu().numpy() - rp)) < 1e-12 and np.max(np.abs(L["ref_yaw"].cpu().numpy() - ry)) < 1e-11, cyc
        assert np.array_equal(L["poly_idx"].cpu().numpy(), pidx) and np.all(L["overflow"].cpu().numpy() == 0), cyc
        assert np.array_equal(L["nrows"].cpu().numpy(), nrows), cyc
        assert np.max(np.abs(L["rows"].cpu().numpy() - rows)) < 1e-8 and np.max(np.abs(L["hdr"].cpu().numpy() - hdr)) < 1e-11, cyc
        c = O.solve_batch(W.Batch(x_c, z0_c, hdr, rows, nrows, 0), opts=O.default_opts(mu0=1.0 if cyc == 0 else 0.1))
        assert np.all(flag == 1) and np.all(c["flag"] == 1), (cyc, flag, c["flag"])
        assert np.array_equal(it, c["it"]), cyc
        assert np.max(np.abs(cmd - c["z"][:, 0, 0:4])) < 1e-6, cyc
        prev = PN.wrap_yaw(c["z"])                       # updateFORCESResults: the adopted plan is kept wrapped
    # the vehicles actually follow their paths
    assert np.max(np.linalg.norm(prev[:, 1, 8:11] - rp[:, 0], axis=1)) < 0.3


def test_scheduling_order_does_not_change_results():
    """nmpc_solve_batch_ordered_f64: CTA i solves problem order[i]; outputs stay at the problems' own indices."""
    import torch
    b = W.config3(300)
    db = S.DeviceBatch(b, np.float64, "cuda:0")
    S.solve_device(db)
    ref = db.result()
    lib = _lib.load()
    fn = lib.nmpc_solve_batch_ordered_f64
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.POINTER(_lib.NmpcOpts)] + [ctypes.c_void_p] * 5
    order = torch.from_numpy(np.random.default_rng(3).permutation(b.B).astype(np.int32)).cuda()
    d = db.d
    z = torch.zeros_like(db.z); ii = torch.zeros_like(db.info_int); ir = torch.zeros_like(db.info_real)
    o = _lib.default_opts()
    rc = fn(b.B, b.N, b.mcap, d["xinit"].data_ptr(), d["z0"].data_ptr(), d["hdr"].data_ptr(), d["rows"].data_ptr(),
            d["nrows"].data_ptr(), 0, ctypes.byref(o), z.data_ptr(), ii.data_ptr(), ir.data_ptr(), order.data_ptr(), None)
    torch.cuda.synchronize()
    assert rc == 0
    assert np.array_equal(z.cpu().numpy(), ref.z) and np.array_equal(ii.cpu().numpy()[:, 1], ref.it)


@pytest.mark.parametrize("mode,tol,tol9,dit", SHIM_MODES)
def test_forces_shim_with_full_30_row_corridors_and_interleaved_zero_rows(monkeypatch, mode, tol, tol9, dit):
    """The reference layout allows 30 rows per stage; DecompROS polytopes can also leave zero rows in
    the middle once tightened rows are dropped upstream.  The shim compacts them; the result must
    equal the native batched call on the compacted problem."""
    _set_shim(monkeypatch, mode)
    rng = np.random.default_rng(4)
    b = W.config3(2, mcap=30)
    w = forces.FORCESNormal()
    for i in range(b.B):
        xinit, x0, allp = W.to_forces_params(b, i)
        allp = allp.reshape(20, 130).copy()
        # fill up to 30 rows with far-away (inactive) planes, then blank a few at random positions
        for k in range(20):
            m = int(b.nrows[i, k])
            for j in range(m, 30):
                a = rng.normal(size=3); a /= np.linalg.norm(a)
                allp[k, 10 + 3 * j:13 + 3 * j] = a
                allp[k, 100 + j] = a @ b.xinit[i, 0:3] + 50.0
            for j in rng.choice(np.arange(m, 30), size=5, replace=False):
                allp[k, 10 + 3 * j:13 + 3 * j] = 0.0; allp[k, 100 + j] = 0.0
        w.params_.xinit[:] = xinit.tolist(); w.params_.x0[:] = x0.tolist()
        w.params_.all_parameters[:] = allp.reshape(-1).tolist()
        assert w.solve_params() == 1
        # native twin: compact the same rows
        rows = np.zeros((1, 20, 30, 4)); nrows = np.zeros((1, 20), np.int32)
        for k in range(20):
            live = [j for j in range(30) if np.any(allp[k, 10 + 3 * j:13 + 3 * j] != 0) or allp[k, 100 + j] != 0]
            nrows[0, k] = len(live)
            for q, j in enumerate(live):
                rows[0, k, q, 0:3] = allp[k, 10 + 3 * j:13 + 3 * j]; rows[0, k, q, 3] = allp[k, 100 + j]
        nb = W.Batch(b.xinit[i:i + 1], b.z0[i:i + 1], b.hdr[i:i + 1], rows, nrows, 0)
        ref = S.solve_host(nb)
        assert ref.flag[0] == 1 and np.max(np.abs(w.output_array() - ref.z[0])) < tol9
        # inactive far planes must not move the solution away from the original problem's
        base = S.solve_host(b.slice(i, i + 1))
        assert np.max(np.abs(base.z[0] - ref.z[0])) < 1e-4


def test_large_batch_chunked_host_path_equals_device_path():
    """B >= 2048 takes the chunked, multi-stream host path; it must return exactly what one launch returns."""
    b = W.config2(5000)
    a, c = S.solve(b), S.solve_host(b)
    assert np.array_equal(a.z, c.z) and np.array_equal(a.flag, c.flag) and np.array_equal(a.it, c.it)
    assert np.all(a.flag == 1)
