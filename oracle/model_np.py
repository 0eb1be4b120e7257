"""numpy restatement of the reference NLP's model layer.  TEST INFRASTRUCTURE ONLY.

Follows (paths relative to /root/reference/src/resilient_planner/plan_manage/):
  * matlab_code/dynamics/nonlinear_dynamics.m:21-40   continuous dynamics
  * matlab_code/dynamics/transit.m:4-8                [x_next; u] stacking of the 13 equalities
  * solver/normal/FORCESNLPsolver_normal_casadi.c:238-240,307-311,383-394  ForcesPro "RK2" = Heun, h=0.05
  * matlab_code/mpc/mpc_objective1.m, mpc/normal/mpc_objective_normal.m,
    mpc/final/mpc_objectiveN_final.m                  stage costs
  * matlab_code/mpc/mpc_corridorconst.m:7-10           corridor rows A*pos - b
  * matlab_code/setup.m:17-43                          constants

Pinned against the reference's CasADi C (oracle/_ref) in tests/test_model_parity.py and against
the committed golden vectors in tests/golden/.
"""
from __future__ import annotations

import numpy as np

N_STAGES = 20
NZ, NX, NU, NXI, NP, NH = 17, 9, 4, 13, 130, 30
DT = 0.05
MASS = 0.745319
GRAV = 9.81
KD = 0.33
HU = 1e-5
RATE_MAX = np.pi / 2

LB = np.array([-RATE_MAX, -RATE_MAX, -RATE_MAX, 0.5 * GRAV * MASS,
               -RATE_MAX, -RATE_MAX, -RATE_MAX, 0.5 * GRAV * MASS,
               -20.0, -20.0, 0.0, -2.0, -2.0, -2.0, -0.4 * np.pi, -0.4 * np.pi, -2 * np.pi])
UB = np.array([RATE_MAX, RATE_MAX, RATE_MAX, 2.0 * GRAV * MASS,
               RATE_MAX, RATE_MAX, RATE_MAX, 2.0 * GRAV * MASS,
               20.0, 20.0, 5.0, 2.0, 2.0, 2.0, 0.4 * np.pi, 0.4 * np.pi, 2 * np.pi])


def zb_and_jac(rpy):
    """Body z axis z_B = R(:,3) and its Jacobian wrt (roll, pitch, yaw)."""
    sr, cr = np.sin(rpy[0]), np.cos(rpy[0])
    sp, cp = np.sin(rpy[1]), np.cos(rpy[1])
    sy, cy = np.sin(rpy[2]), np.cos(rpy[2])
    zb = np.array([cy * sp * cr + sy * sr, sy * sp * cr - cy * sr, cp * cr])
    Z = np.array([[-cy * sp * sr + sy * cr, cy * cp * cr, -sy * sp * cr + cy * sr],
                  [-sy * sp * sr - cy * cr, sy * cp * cr, cy * sp * cr + sy * sr],
                  [-cp * sr, -sp * cr, 0.0]])
    return zb, Z


def acc_and_jac(v, rpy, T, fext):
    """acc = z_B T/m + f_ext - g e3 - kd (v - z_B (z_B.v));  R diag(kd,kd,0) R' = kd (I - z_B z_B')."""
    zb, Z = zb_and_jac(rpy)
    zv = zb @ v
    a = zb * (T / MASS) + fext - np.array([0, 0, GRAV]) - KD * (v - zb * zv)
    Av = -KD * (np.eye(3) - np.outer(zb, zb))
    Ar = (T / MASS + KD * zv) * Z + KD * np.outer(zb, v @ Z)
    AT = zb / MASS
    return a, Av, Ar, AT


def dynamics(z, fext, jac=True):
    """c(z) = [Heun(x,u) (9); u (4)] and its 13x17 Jacobian (transit.m + ForcesPro RK2)."""
    u, x = z[0:4], z[8:17]
    w, T = u[0:3], u[3]
    p, v, r = x[0:3], x[3:6], x[6:9]
    h = DT
    a1, A1v, A1r, A1T = acc_and_jac(v, r, T, fext)
    v2 = v + h * a1
    r2 = r + h * w
    a2, A2v, A2r, A2T = acc_and_jac(v2, r2, T, fext)
    c = np.concatenate([p + h * v + 0.5 * h * h * a1, v + 0.5 * h * (a1 + a2), r + h * w, u])
    if not jac:
        return c
    J = np.zeros((13, 17))
    I3 = np.eye(3)
    # pos+
    J[0:3, 8:11] = I3
    J[0:3, 11:14] = h * I3 + 0.5 * h * h * A1v
    J[0:3, 14:17] = 0.5 * h * h * A1r
    J[0:3, 3] = 0.5 * h * h * A1T
    # vel+
    da2_v = A2v @ (I3 + h * A1v)
    da2_r = A2v @ (h * A1r) + A2r
    da2_T = A2v @ (h * A1T) + A2T
    J[3:6, 11:14] = I3 + 0.5 * h * (A1v + da2_v)
    J[3:6, 14:17] = 0.5 * h * (A1r + da2_r)
    J[3:6, 3] = 0.5 * h * (A1T + da2_T)
    J[3:6, 0:3] = 0.5 * h * h * A2r
    # rpy+
    J[6:9, 14:17] = I3
    J[6:9, 0:3] = h * I3
    # u copy
    J[9:13, 0:4] = np.eye(4)
    return c, J


def objective(z, p, stage, variant="normal", n_stages=N_STAGES):
    """Stage cost, gradient and (constant) Hessian.  stage is 0-indexed."""
    ref, wwp, win, wrate, yawref = p[0:3], p[6], p[7], p[8], p[9]
    u, up, pos, vel, yaw = z[0:4], z[4:8], z[8:11], z[11:14], z[16]
    f = wwp * np.sum((ref - pos) ** 2) + 12 * wwp * (yawref - yaw) ** 2
    f += win * np.sum((u[0:3] / RATE_MAX) ** 2) + wrate * np.sum((u - up) ** 2)
    g = np.zeros(17)
    H = np.zeros((17, 17))
    g[0:3] = 2 * win * u[0:3] / RATE_MAX ** 2
    g[0:4] += 2 * wrate * (u - up)
    g[4:8] = -2 * wrate * (u - up)
    g[8:11] = -2 * wwp * (ref - pos)
    g[16] = -24 * wwp * (yawref - yaw)
    for i in range(4):
        H[i, i] = 2 * wrate + (2 * win / RATE_MAX ** 2 if i < 3 else 0.0)
        H[4 + i, 4 + i] = 2 * wrate
        H[i, 4 + i] = H[4 + i, i] = -2 * wrate
    for i in range(3):
        H[8 + i, 8 + i] = 2 * wwp
    H[16, 16] = 24 * wwp
    if stage == 0:
        f += 10 * win * np.sum(up[0:3] ** 2)
        g[4:7] += 20 * win * up[0:3]
        for i in range(3):
            H[4 + i, 4 + i] += 20 * win
    if stage == n_stages - 1 and variant == "final":
        f += 20 * wwp * np.sum(vel ** 2)
        g[11:14] += 40 * wwp * vel
        for i in range(3):
            H[11 + i, 11 + i] += 40 * wwp
    return f, g, H


def corridor(z, p):
    """h = A pos - b (30 rows), Jacobian = A in cols 8..10 (mpc_corridorconst.m:7-10)."""
    A = p[10:100].reshape(30, 3)
    b = p[100:130]
    hval = A @ z[8:11] - b
    J = np.zeros((30, 17))
    J[:, 8:11] = A
    return hval, J
