"""Host-side pieces around the solve: the numpy restatements in oracle/prep_np.py
against hand-computed known answers, and the exit-code acceptance policy (Python and C++ twins).  CPU only."""
import math
import os
import subprocess

import numpy as np

from forces_resilient_planner_b200 import forces
from oracle import prep_np as PN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_sampling_on_a_straight_line():
    # 1 m/s along +x sampled at Ts = 0.05 (kino_path_), plan starts 0.02 s after the path (t_off)
    Ts, N, P = 0.05, 20, 40
    path = np.zeros((1, P, 3)); path[0, :, 0] = 0.05 * np.arange(P); path[0, :, 2] = 1.0
    rp, ry, far = PN.sample_reference(path, np.array([P]), np.array([0.02]), np.array([0.5]), N, Ts,
                                      pos1=np.array([[0.02, 0.0, 1.0]]))
    assert np.allclose(rp[0, :, 0], 0.05 * np.arange(N) + 0.02, atol=1e-12) and np.all(rp[0, :, 2] == 1.0)
    # heading 0 towards the look-ahead point; low-pass from last_yaw = 0.5: yaw_i = 0.5 * 0.2^(i+1)
    assert np.allclose(ry[0], 0.5 * 0.2 ** (np.arange(N) + 1), atol=1e-15)
    assert far[0] == 0


def test_reference_sampling_clamps_at_the_path_end_and_flags_a_far_start():
    Ts, N = 0.05, 20
    path = np.zeros((1, 8, 3)); path[0, :, 1] = 0.1 * np.arange(8)
    rp, ry, far = PN.sample_reference(path, np.array([6]), np.array([0.0]), np.array([0.0]), N, Ts,
                                      pos1=np.array([[2.0, 0.0, 0.0]]))
    assert np.allclose(rp[0, :5, 1], 0.1 * np.arange(5)) and np.all(rp[0, 5:, 1] == 0.5)   # only 6 live points
    assert far[0] == 1                                                                       # 2 m away at index 0
    # once the reference sits on the last point the direction is shorter than 0.1: yaw keeps decaying towards itself
    assert abs(ry[0, -1] - ry[0, -2]) < 1e-12


def test_yaw_unwrap_uses_the_reference_constant():
    # heading -3.1 rad while the running yaw is +3.0: |diff| > PI  ->  yaw_temp + 2 PI, with PI = 3.1415926
    Ts = 0.05
    path = np.zeros((1, 12, 3))
    ang = -3.1
    path[0, :, 0] = math.cos(ang) * 0.1 * np.arange(12); path[0, :, 1] = math.sin(ang) * 0.1 * np.arange(12)
    _, ry, _ = PN.sample_reference(path, np.array([12]), np.array([0.0]), np.array([3.0]), 1, Ts)
    assert abs(ry[0, 0] - (0.2 * 3.0 + 0.8 * (ang + 2 * 3.1415926))) < 1e-12
    z = np.zeros((1, 2, 17)); z[0, 0, 16] = 3.2; z[0, 1, 16] = -3.3
    w = PN.wrap_yaw(z)
    assert w[0, 0, 16] == 3.2 - 2 * 3.1415926 and w[0, 1, 16] == -3.3 + 2 * 3.1415926


def _scenario(policy, codes):
    return [(policy.consume(c), policy.fail_count, policy.replan_count, policy.kino_replan) for c in codes]


def test_acceptance_policy_follows_solveNMPC():
    p = forces.SolveAcceptance()
    assert p.consume(1) is True and not p.next_solve_is_cold()
    # three failures in a row force a front-end replan; the output is not adopted
    out = _scenario(p, [-7, 0, -6])
    assert [o[0] for o in out] == [False, False, False] and out[-1][1:] == (0, 1, True)
    assert p.next_solve_is_cold()
    # a maxit exit is only tolerated after more than three forced replans
    p = forces.SolveAcceptance(); p.replan_count = 3
    assert p.consume(0) is False
    p.replan_count = 4
    assert p.consume(0) is True and (p.fail_count, p.replan_count) == (0, 0)
    p.replan_count = 4
    assert p.consume(-7) is False                     # any other code never is
    assert forces.SolveAcceptance().next_solve_is_cold(initialized_output=False)


def test_cpp_acceptance_policy_is_the_same(tmp_path):
    codes = [1, 0, -7, -6, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, -100, 0, 2, 1]
    src = tmp_path / "acc.cpp"
    src.write_text('#include "forces_wrappers.hpp"\n#include <cstdio>\nint main(){ resilient_planner::SolveAcceptance p; int codes[] = {'
                   + ",".join(map(str, codes)) + '};\nfor (int c : codes) { bool u = p.consume(c); '
                   'std::printf("%d %d %d %d %d\\n", (int)u, p.fail_count, p.replan_count, (int)p.kino_replan, (int)p.next_solve_is_cold(true)); }\nreturn 0; }\n')
    exe = tmp_path / "acc"
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "forces_resilient_planner_b200", "host"), str(src), "-o", str(exe)])
    got = [tuple(map(int, ln.split())) for ln in subprocess.check_output([str(exe)]).decode().splitlines()]
    p = forces.SolveAcceptance()
    want = []
    for c in codes:
        u = p.consume(c)
        want.append((int(u), p.fail_count, p.replan_count, int(p.kino_replan), int(p.next_solve_is_cold(True))))
    assert got == want


def test_ellipsoid_restatement_properties():
    """oracle/ellipsoid_np.py against facts that do not depend on its own code path."""
    from oracle import ellipsoid_np as EN
    c = EN.EllipsoidConsts()
    z = np.zeros((20, 17)); z[:, 3] = c.mass * 9.81; z[:, 16] = 0.3        # hover, yawed
    E = EN.propagate(z, c)
    # stage 0: the ego ellipsoid itself (hover: R = Rz, the disc is rotation invariant)
    assert np.allclose(E[0], np.diag([c.ego_r, c.ego_r, c.ego_h]), atol=1e-15)
    # the shape matrices grow monotonically along the horizon (disturbance accumulates)
    tr = np.trace(E, axis1=1, axis2=2)
    assert np.all(np.diff(tr) > 0)
    # the Sylvester solution is the finite-horizon Gramian: check one against numerical quadrature
    Phi, _ = EN.update_matrix(z[0, 14:17], z[0, 11:14], z[0, 3], c)
    import scipy.linalg as sla
    t = c.Ts
    N0 = np.zeros((9, 9)); N0[3, 3] = t * c.ext_noise_bound ** 2
    X = sla.solve_sylvester(Phi, Phi.T, N0 - sla.expm(-Phi * t) @ N0 @ sla.expm(-Phi.T * t))
    xs, ws = np.polynomial.legendre.leggauss(12)
    Xq = sum(w * 0.5 * t * (sla.expm(-Phi * (0.5 * t * (x + 1))) @ N0 @ sla.expm(-Phi.T * (0.5 * t * (x + 1)))) for x, w in zip(xs, ws))
    assert np.max(np.abs(X - Xq)) < 1e-12 * np.max(np.abs(X))
    # closed loop used by the reference is stable at hover (otherwise the Sylvester equation could be singular)
    assert np.max(np.linalg.eigvals(Phi).real) < 0


def _scene(rng, n_pts=300, clear=0.35):
    """A straight reference through random clutter, with a clear tube around it."""
    N = 20
    ref = np.zeros((N, 3)); ref[:, 0] = 0.25 * np.arange(N); ref[:, 1] = 0.3 * np.sin(0.3 * np.arange(N)); ref[:, 2] = 1.0
    yaw = np.arctan2(np.gradient(ref[:, 1]), np.gradient(ref[:, 0]))
    pts = rng.uniform([-1.5, -3.0, 0.0], [6.5, 3.0, 2.2], (n_pts, 3))
    dist = np.min(np.linalg.norm(pts[:, None, :] - ref[None], axis=2), axis=1)
    return ref, yaw, pts[dist > clear]


def test_corridor_restatement_properties():
    """oracle/corridor_np.py against facts that do not depend on its own code path."""
    from oracle import corridor_np as CN
    p1 = np.array([1.0, 2.0, 1.0]); p2 = p1 + np.array([0.1, 0.0, 0.0])
    # no obstacles: only the local box, 2 m beyond each end, 2 m to the sides, 1 m up and down
    A, b = CN.dilate_segment(p1, p2, np.zeros((0, 3)))
    assert A.shape == (6, 3) and np.allclose(np.linalg.norm(A, axis=1), 1.0)
    ext = {tuple(np.round(a, 12)): bb for a, bb in zip(A, b)}
    assert np.isclose(ext[(1.0, 0.0, 0.0)], 1.1 + 2.0) and np.isclose(ext[(-1.0, 0.0, 0.0)], -(1.0 - 2.0))
    assert np.isclose(ext[(0.0, 0.0, 1.0)], 2.0) and np.isclose(ext[(0.0, 0.0, -1.0)], 0.0)
    # one obstacle: the seed ellipsoid is a sphere, so the plane is tangent to the sphere's metric at the point
    pt = np.array([2.0, 2.5, 1.2])
    A, b = CN.dilate_segment(p1, p2, pt[None])
    n = (pt - (p1 + p2) / 2); n /= np.linalg.norm(n)
    assert A.shape == (7, 3) and np.allclose(A[0], n) and np.isclose(b[0], n @ pt)
    # a second point hidden behind the first plane produces no plane of its own; one in front does
    A2, _ = CN.dilate_segment(p1, p2, np.stack([pt, pt + 0.5 * n, np.array([1.05, 1.2, 1.0])]))
    assert A2.shape == (8, 3)
    # the safe-corridor property on clutter: midpoint strictly inside, every obstacle of the box on or outside a face
    rng = np.random.default_rng(0)
    ref, yaw, cloud = _scene(rng)
    polys, idx = CN.select_corridors(ref, yaw, np.tile(np.diag([0.27, 0.27, 0.0425]), (20, 1, 1)), cloud)
    assert len(polys) >= 2 and np.all(np.diff(idx) >= 0) and idx[0] == 0 and idx[-1] == len(polys) - 1
    for k, (A, b) in enumerate(polys):
        first = int(np.argmax(idx == k))
        mid = ref[first] + 0.05 * np.array([np.cos(yaw[first]), np.sin(yaw[first]), 0.0])
        assert np.all(A @ mid - b < 0)
        assert np.all(np.max(cloud @ A.T - b, axis=1) > -1e-9)
    # an obstacle within the 5 cm seed sphere shrinks the ellipsoid instead of being swallowed
    C, d = CN.find_ellipsoid(p1, p2, [np.array([1.05, 2.03, 1.0])])
    assert np.linalg.norm(np.linalg.inv(C) @ (np.array([1.05, 2.03, 1.0]) - d)) >= 1 - 1e-9


def test_corridor_restatement_degenerate_seeds():
    """Edge cases of add_local_bbox / vec3_to_rotation: a vertical seed segment (no horizontal direction: the
    reference falls back to dir_h = (-1, 0, 0)) and obstacle points exactly on a box face (kept, non-exclusive)."""
    from oracle import corridor_np as CN
    p1 = np.array([0.0, 0.0, 1.0]); p2 = np.array([0.0, 0.0, 1.1])
    planes = CN.local_bbox_planes(p1, p2, (2.0, 2.0, 1.0))
    normals = np.array([n for _, n in planes])
    assert np.allclose(normals[0], [-1, 0, 0]) and np.allclose(normals[2], [0, 0, 1])
    assert np.allclose(np.abs(np.linalg.det(np.stack([normals[0], normals[2], normals[4]]))), 1.0)   # orthonormal frame
    A, b = CN.dilate_segment(p1, p2, np.array([[0.5, 0.0, 1.05], [2.0, 0.0, 1.05], [2.0 + 1e-6, 0.0, 1.05]]))
    assert A.shape[0] == 7                                   # the first point carves; the one on the face is behind that plane; the one outside is ignored
    assert np.all(A @ ((p1 + p2) / 2) - b < 0)
    R = CN.vec3_to_rotation(np.array([0.0, 0.0, 0.1]))
    assert np.allclose(R @ np.array([1.0, 0.0, 0.0]), [0.0, 0.0, 1.0])      # the ellipsoid's long axis follows the segment
