"""numpy/scipy restatement of the reference's disturbance-ellipsoid propagation.  TEST INFRASTRUCTURE ONLY
(checker for csrc/nmpc_ellipsoid.cuh); never imported by the product package.

Follows, statement by statement, NMPCSolver::setFORCESParams (the ellipsoid part),
NMPCSolver::updateMatrix, NMPCSolver::getDistrEllipsoid and NMPCSolver::eulerToRot of
/root/reference/src/resilient_planner/plan_manage/src/nmpc_solver.cpp (:484-521, :615-699, :567-611,
:552-564) and the constructor's constant matrices (:8-32).  Like the reference it solves the
Sylvester equation by Bartels-Stewart (scipy.linalg.solve_sylvester) and uses Pade matrix
exponentials (scipy.linalg.expm) -- the device kernel uses a different (series) formulation of the same
quantities, so the two check each other.

One deliberate deviation: the reference accumulates `temp += sqrt(X.trace())` into an UNINITIALISED
`double temp;` (nmpc_solver.cpp:573, :597 -- undefined behaviour; whatever the stack held).  Here, and in
the kernel, temp starts at 0, which is what the formula means (SURVEY.md §8f rank 2 notes the bug).
A second one: `updateMatrix` never assigns At_(5,8) (no thrust term in d acc_z / d yaw), it only accumulates into it
(`At_(5,8) += ...`, :689) on the class member At_ -- in the reference that entry keeps growing over the 20 stages of a
replan and across replans.  `update_matrix` below (and the kernel) builds every stage's At from zero.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import scipy.linalg as sla

NX, NU, NW = 9, 4, 3                      # nmpc_utils.h:202-204

KT = np.array([[-2.0, 5.0, 0.0, -1.0, 4.0, 0.0, -8.0, 0.0, 0.0],      # nmpc_solver.cpp:28-31
               [-5.0, -2.0, 0.0, -4.0, -1.0, 0.0, 0.0, -8.0, 0.0],
               [-2.0, -2.0, 0.0, -1.0, -1.0, 0.0, 0.0, 0.0, -8.0],
               [0.0, 0.0, -8.0, 0.0, 0.0, -6.0, 0.0, 0.0, 0.0]])


@dataclass
class EllipsoidConsts:
    mass: float = 0.745319          # rotors_sim.launch:53  (nh.param default 0.74, nmpc_solver.cpp:78)
    drag: float = 0.33              # nmpc/drag_coefficient
    ego_r: float = 0.27             # nmpc/ego_r
    ego_h: float = 0.0425           # nmpc/ego_h
    ext_noise_bound: float = 0.5    # nmpc/ext_noise_bound  -> w_
    epsilon: float = 0.06           # nmpc_utils.h:188
    Ts: float = 0.05


def euler_to_rot(rpy):
    """eulerToRot (:552-564): R = Rz(yaw) Ry(pitch) Rx(roll) through unit quaternions."""
    cr, sr = math.cos(rpy[0]), math.sin(rpy[0])
    cp, sp = math.cos(rpy[1]), math.sin(rpy[1])
    cy, sy = math.cos(rpy[2]), math.sin(rpy[2])
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def update_matrix(euler, vel, thrust_c, c: EllipsoidConsts):
    """updateMatrix (:615-699): A_t, B_t of the linearised model at (euler, vel, thrust); returns (Phi, R_cur)."""
    At = np.zeros((NX, NX)); Bt = np.zeros((NX, NU))
    At[0, 3] = At[1, 4] = At[2, 5] = 1.0                      # constructor, :16-18
    Bt[6, 0] = Bt[7, 1] = Bt[8, 2] = 1.0                      # :20-22
    roll, pitch, yaw = euler
    v1, v2, v3 = vel
    cos, sin = math.cos, math.sin
    comb0 = thrust_c * 1.0 / c.mass
    comb5 = cos(pitch) * sin(pitch)
    comb6 = cos(pitch) * sin(roll)
    comb7 = cos(pitch) * cos(roll)
    comb8 = sin(pitch) * cos(roll)
    comb9 = sin(pitch) * sin(roll)
    comb1 = cos(roll) * sin(yaw) - comb9 * cos(yaw)
    comb2 = sin(roll) * cos(yaw) - comb8 * sin(yaw)
    comb3 = cos(roll) * cos(yaw) + comb9 * sin(yaw)
    comb4 = sin(roll) * sin(yaw) + comb8 * cos(yaw)
    At[3, 6] = comb0 * comb1; At[4, 6] = -comb0 * comb3; At[5, 6] = -comb0 * comb6
    At[3, 7] = comb0 * comb7 * cos(yaw); At[4, 7] = comb0 * comb7 * sin(yaw); At[5, 7] = -comb0 * comb8
    At[3, 8] = comb0 * comb2; At[4, 8] = comb0 * comb4
    R_cur = euler_to_rot(euler)
    Dm = np.diag([c.drag, c.drag, 0.0])
    At[3:6, 3:6] = R_cur @ Dm @ R_cur.T
    drag = c.drag
    cos_pitch_sq = cos(pitch) ** 2; sin_pitch_sq = sin(pitch) ** 2
    cos_yaw_sq = cos(yaw) ** 2; sin_yaw_sq = sin(yaw) ** 2; sin_roll_sq = sin(roll) ** 2
    sq1 = comb3 ** 2; sq2 = comb1 ** 2
    t10 = comb6 * comb4 - comb7 * comb1
    t11 = comb3 * comb4 + comb1 * comb2
    t12 = comb6 * comb2 - comb7 * comb3
    At[3, 6] += drag * (v3 * t10 + v2 * t11 - 2 * v1 * comb4 * comb1)
    At[4, 6] += drag * (v1 * t11 - v3 * t12 - 2 * v2 * comb3 * comb2)
    At[5, 6] += drag * (v1 * t10 - v2 * t12 + 2 * v3 * comb7 * comb6)
    t20 = cos(yaw) * (sin_pitch_sq - cos_pitch_sq + cos_pitch_sq * sin_roll_sq) + comb9 * comb1
    t21 = 2 * comb5 * cos(yaw) * sin(yaw) - comb6 * (cos(yaw) * comb3 + sin(yaw) * comb1)
    t22 = sin(yaw) * (cos_pitch_sq - sin_pitch_sq - cos_pitch_sq * sin_roll_sq) + comb9 * comb3
    At[3, 7] += drag * (v3 * t20 - v2 * t21 - v1 * 2 * (comb5 * cos_yaw_sq + comb6 * comb1 * cos(yaw)))
    At[4, 7] += -drag * (v3 * t22 - v1 * t21 - v2 * 2 * (comb5 * sin_yaw_sq - comb6 * comb3 * sin(yaw)))
    At[5, 7] += drag * (v1 * t20 - v2 * t22 + v3 * 2 * (comb5 - comb5 * sin_roll_sq))
    t30 = 2 * drag * (comb3 * comb1 - cos_pitch_sq * cos(yaw) * sin(yaw))
    t31 = drag * (comb6 * comb3 - comb5 * sin(yaw))
    t32 = drag * (sq1 - sq2 - cos_pitch_sq * cos_yaw_sq + cos_pitch_sq * sin_yaw_sq)
    t33 = drag * (comb6 * comb1 + comb5 * cos(yaw))
    At[3, 8] += v1 * t30 - v3 * t31 - v2 * t32
    At[4, 8] += -v1 * t32 - v3 * t33 - v2 * t30
    At[5, 8] += -v2 * t33 - v1 * t31
    Bt[3, 3] = 1.0 / c.mass * comb4; Bt[4, 3] = -1.0 / c.mass * comb2; Bt[5, 3] = 1.0 / c.mass * comb7
    return At + Bt @ KT, R_cur


def distr_ellipsoid(Phi, t, Q_origin, c: EllipsoidConsts):
    """getDistrEllipsoid (:567-611).  Returns (position block of the propagated shape matrix, Q_update)."""
    Dt = np.zeros((NX, NW)); Dt[3, 0] = Dt[4, 1] = Dt[5, 2] = 1.0          # constructor, :24-26
    w = np.full(NW, c.ext_noise_bound)
    temp = 0.0                                   # the reference leaves this uninitialised (:573), see module docstring
    temp_Q = np.zeros((NX, NX))
    Em = sla.expm(-Phi * t)
    for i in range(NW):
        Nt = t * w[i] * w[i] * np.outer(Dt[:, i], Dt[:, i])
        Array_Q = Nt - Em @ Nt @ Em.T
        X = sla.solve_sylvester(Phi, Phi.T, Array_Q)                       # Phi X + X Phi' = W (Bartels-Stewart)
        temp += math.sqrt(np.trace(X))
        temp_Q += X / math.sqrt(np.trace(X))
    Qd = temp * temp_Q
    beta = math.sqrt(np.trace(Q_origin) / np.trace(Qd))
    Q_update = (1 + 1 / beta) * Q_origin + (1 + beta) * Qd
    Ep = sla.expm(Phi * t)
    position_Q = Ep @ Q_update @ Ep.T
    return position_Q[0:3, 0:3].copy(), Q_update


def sqrtm3(Q):
    """E = V sqrt(Lambda) V^-1, real part (setFORCESParams :511-512, Eigen::EigenSolver)."""
    lam, V = np.linalg.eig(Q)
    return (V @ np.diag(np.sqrt(lam.astype(complex))) @ np.linalg.inv(V)).real


def propagate(z, c: EllipsoidConsts | None = None):
    """The ellipsoid part of setFORCESParams (:484-521) for one plan z [N,17] (the previous mpc_output_):
    returns the N matrices E_i (ellipsoid_matrices_) as [N,3,3]."""
    c = c or EllipsoidConsts()
    N = z.shape[0]
    ego = np.diag([c.ego_r ** 2, c.ego_r ** 2, c.ego_h ** 2])            # ego_size_, :88-90
    Q_init = c.epsilon ** 2 * np.eye(NX)                                  # :487
    out = np.zeros((N, 3, 3))
    Q2 = None
    for i in range(N):
        euler = z[i, 14:17]; vel = z[i, 11:14]
        Phi, R_cur = update_matrix(euler, vel, z[i, 3], c)                 # :501
        Q1 = R_cur @ ego @ R_cur.T                                         # :503
        if i == 0:
            Q = Q1
        else:
            beta = math.sqrt(np.trace(Q1) / np.trace(Q2))                  # :508
            Q = (1 + 1 / beta) * Q1 + (1 + beta) * Q2
        out[i] = sqrtm3(Q)                                                 # :511-512
        Q2, Q_init = distr_ellipsoid(Phi, c.Ts, Q_init, c)                 # :519
    return out


def propagate_batch(z, c: EllipsoidConsts | None = None):
    return np.stack([propagate(z[b], c) for b in range(z.shape[0])])
