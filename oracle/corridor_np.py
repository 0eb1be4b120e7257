"""numpy restatement of the reference's corridor generation and per-stage polytope selection.
TEST INFRASTRUCTURE ONLY (checker for csrc/nmpc_corridor.cuh); never imported by the
product package.

Follows, statement by statement,
  * NMPCSolver::getSikangConst and the loop of NMPCSolver::setFORCESParams that calls it
    (/root/reference/src/resilient_planner/plan_manage/src/nmpc_solver.cpp:288-332, 484-521),
  * DecompROS as vendored by the reference (/root/reference/src/ThirdParty/DecompROS/decomp_ros_utils/include/):
    EllipsoidDecomp::dilate (decomp_util/ellipsoid_decomp.h:76-100), LineSegment::dilate / add_local_bbox /
    find_ellipsoid<3> (decomp_util/line_segment.h:31-35, 46-88, 137-208), DecompBase::set_obs / find_polyhedron
    (decomp_util/decomp_base.h:33-38, 66-85), Ellipsoid::dist / closest_point / closest_hyperplane
    (decomp_geometry/ellipsoid.h:22-24, 42-60), Polyhedron::inside (decomp_geometry/polyhedron.h:47-55),
    LinearConstraint(p0, hyperplanes) (polyhedron.h:100-120), vec3_to_rotation (geometric_utils.h:27-35).
Plain loops, one agent at a time.
"""
from __future__ import annotations

import math

import numpy as np

EPS = 1e-10                      # decomp_basis/data_type.h:129


def _rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def _rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def _rot_z(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def vec3_to_rotation(v):
    pitch = math.atan2(-v[2], math.hypot(v[0], v[1]))
    yaw = math.atan2(v[1], v[0])
    return _rot_z(yaw) @ _rot_y(pitch)             # qz * qy * qx with zero roll


def local_bbox_planes(p1, p2, bbox):
    """LineSegment::add_local_bbox: six (point, outward normal) pairs."""
    dir_ = (p2 - p1) / np.linalg.norm(p2 - p1)
    dir_h = np.array([dir_[1], -dir_[0], 0.0])
    if np.linalg.norm(dir_h) == 0:
        dir_h = np.array([-1.0, 0.0, 0.0])
    dir_h = dir_h / np.linalg.norm(dir_h)
    dir_v = np.cross(dir_, dir_h)
    return [(p1 + dir_h * bbox[1], dir_h), (p1 - dir_h * bbox[1], -dir_h),
            (p2 + dir_ * bbox[0], dir_), (p1 - dir_ * bbox[0], -dir_),
            (p1 + dir_v * bbox[2], dir_v), (p1 - dir_v * bbox[2], -dir_v)]


def _edist(Cinv, d, pt):
    return float(np.linalg.norm(Cinv @ (pt - d)))


def _closest(Cinv, d, pts):
    best, bi = float("inf"), -1
    for i, p in enumerate(pts):
        dd = _edist(Cinv, d, p)
        if dd < best:
            best, bi = dd, i
    return bi


def find_ellipsoid(p1, p2, obs, offset_x=0.0):
    """LineSegment<3>::find_ellipsoid.  Returns (C, d)."""
    f = float(np.linalg.norm(p1 - p2)) / 2
    axes = np.array([f + offset_x, f, f])
    C = np.diag([f + offset_x, f, f])
    if axes[0] > 0:
        ratio = axes[1] / axes[0]
        axes = axes * ratio
        C = C * ratio
    Ri = vec3_to_rotation(p2 - p1)
    C = Ri @ C @ Ri.T
    d = (p1 + p2) / 2
    Rf = Ri
    obs_in = [p for p in obs if _edist(np.linalg.inv(C), d, p) <= 1]
    obs0 = list(obs_in)
    while obs_in:
        pw = obs_in[_closest(np.linalg.inv(C), d, obs_in)]
        p = Ri.T @ (pw - d)
        roll = math.atan2(p[2], p[1])
        Rf = Ri @ _rot_x(roll)
        p = Rf.T @ (pw - d)
        if p[0] < axes[0]:
            axes[1] = abs(p[1]) / math.sqrt(1 - (p[0] / axes[0]) ** 2)
        C = Rf @ np.diag([axes[0], axes[1], axes[1]]) @ Rf.T
        Ci = np.linalg.inv(C)
        obs_in = [q for q in obs_in if 1 - _edist(Ci, d, q) > EPS]
    C = Rf @ np.diag([axes[0], axes[1], axes[2]]) @ Rf.T
    Ci = np.linalg.inv(C)
    obs_in = [q for q in obs0 if _edist(Ci, d, q) <= 1]
    while obs_in:
        pw = obs_in[_closest(np.linalg.inv(C), d, obs_in)]
        p = Rf.T @ (pw - d)
        dd = 1 - (p[0] / axes[0]) ** 2 - (p[1] / axes[1]) ** 2
        if dd > EPS:
            axes[2] = abs(p[2]) / math.sqrt(dd)
        C = Rf @ np.diag([axes[0], axes[1], axes[2]]) @ Rf.T
        Ci = np.linalg.inv(C)
        obs_in = [q for q in obs_in if 1 - _edist(Ci, d, q) > EPS]
    return C, d


def find_polyhedron(C, d, obs):
    """DecompBase::find_polyhedron: list of (point, outward normal)."""
    Ci = np.linalg.inv(C)
    planes = []
    remain = list(obs)
    while remain:
        pt = remain[_closest(Ci, d, remain)]
        n = Ci @ Ci.T @ (pt - d)
        n = n / np.linalg.norm(n)
        planes.append((pt, n))
        remain = [q for q in remain if float(n @ (q - pt)) < 0]
    return planes


def dilate_segment(p1, p2, cloud, bbox=(2.0, 2.0, 1.0)):
    """EllipsoidDecomp::dilate for one segment + get_constraints: returns (A [m,3], b [m])."""
    box = local_bbox_planes(p1, p2, bbox)
    obs = [q for q in cloud if all(float(n @ (q - p)) <= EPS for p, n in box)]       # set_obs
    C, d = find_ellipsoid(p1, p2, obs)
    planes = find_polyhedron(C, d, obs) + box
    p0 = (p1 + p2) / 2
    A = np.zeros((len(planes), 3)); b = np.zeros(len(planes))
    for i, (p, n) in enumerate(planes):
        c = float(p @ n)
        if float(n @ p0) - c > 0:
            n, c = -n, -c
        A[i] = n; b[i] = c
    return A, b


def select_corridors(ref_pos, ref_yaw, E, cloud, bbox=(2.0, 2.0, 1.0)):
    """The poly_indices / poly_constraints_ part of setFORCESParams (:493-516) with getSikangConst (:288-332):
    walk the stages; keep the last polytope while the reference point, inflated by 1.1 ||E_i a_j||, stays inside,
    otherwise dilate a new one around the 0.1 m seed segment along the yaw reference.
    Returns (polys = [(A, b), ...], poly_idx [N])."""
    N = ref_pos.shape[0]
    polys, idx = [], np.zeros(N, np.int32)
    for i in range(N):
        if polys:
            A, b = polys[-1]
            ok = True
            for j in range(len(b)):
                if float(A[j] @ ref_pos[i]) - (b[j] - 1.1 * float(np.linalg.norm(E[i] @ A[j]))) > 0:
                    ok = False
                    break
            if ok:
                idx[i] = len(polys) - 1
                continue
        p1 = ref_pos[i].copy()
        p2 = p1 + np.array([0.1 * math.cos(ref_yaw[i]), 0.1 * math.sin(ref_yaw[i]), 0.0])
        polys.append(dilate_segment(p1, p2, cloud, bbox))
        idx[i] = len(polys) - 1
    return polys, idx
