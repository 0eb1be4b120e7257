"""numpy restatements of the reference's host code around the solve.  TEST INFRASTRUCTURE ONLY
(checkers for the device kernels of csrc/nmpc_prep.cuh); never imported by the product package.

Reference paths relative to /root/reference/src/resilient_planner/plan_manage/src/.
"""
from __future__ import annotations

import math

import numpy as np

REF_PI = 3.1415926          # nmpc_solver.cpp:3 -- the reference's own (truncated) constant


def pack_params_reference(ref_pos, ref_yaw, ext_acc, ellipsoid, poly_A, poly_b, poly_m, poly_idx, weights5, mcap):
    """FORCESNormal::solveNormal packing loop (forces_normal.cpp:100-136): weights, refs, f_ext, polytope by
    poly_idx, tightening b_j - ||E_i a_j||_2 (:124-125), truncation at mcap (the reference: 30, :114)."""
    B, N, _ = ref_pos.shape
    hdr = np.zeros((B, N, 10))
    rows = np.zeros((B, N, mcap, 4))
    nrows = np.zeros((B, N), np.int32)
    hdr[:, :, 0:3] = ref_pos
    hdr[:, :, 3:6] = ext_acc[:, None, :]
    hdr[:, :, 9] = ref_yaw
    hdr[:, :, 6], hdr[:, :, 7], hdr[:, :, 8] = weights5[0], weights5[1], weights5[2]
    hdr[:, -1, 6], hdr[:, -1, 7] = weights5[3], weights5[4]
    for b in range(B):
        for i in range(N):
            pi = int(poly_idx[b, i])
            m = min(int(poly_m[b, pi]), mcap)
            A = poly_A[b, pi, :m]
            E = ellipsoid[b, i].reshape(3, 3)
            rows[b, i, :m, 0:3] = A
            rows[b, i, :m, 3] = poly_b[b, pi, :m] - np.linalg.norm(A @ E.T, axis=1)
            nrows[b, i] = m
    return hdr, rows, nrows


def sample_reference(kino_path, kino_size, t_off, last_yaw, N, Ts, pos1=None):
    """NMPCSolver::getCurTraj (nmpc_solver.cpp:109-142) + calculate_yaw (:834-862), called for
    index = 0..N-1 as setFORCESParams does (:484-493).  Plain loops, one agent at a time."""
    B = kino_path.shape[0]
    ref_pos = np.zeros((B, N, 3)); ref_yaw = np.zeros((B, N)); far = np.zeros(B, np.int32)
    for b in range(B):
        path, size, last = kino_path[b], int(kino_size[b]), float(last_yaw[b])
        for index in range(N):
            index_time = index * Ts + float(t_off[b])                       # :111
            kino_index = int(index_time / Ts)                                # :112 (truncation)
            if kino_index + 1 < size:                                        # :115-118
                ref = path[kino_index] + math.fmod(index_time, Ts) / Ts * (path[kino_index + 1] - path[kino_index])
            else:
                ref = path[size - 1].copy()
            fwd = path[kino_index + 5] if kino_index + 5 < size else path[size - 1]   # :125-132
            d = fwd - ref                                                    # calculate_yaw, :838-858
            yaw_temp = math.atan2(d[1], d[0]) if math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) > 0.1 else last
            if abs(yaw_temp - last) > REF_PI:
                yaw = yaw_temp - 2 * REF_PI if yaw_temp > 0 else yaw_temp + 2 * REF_PI
            else:
                yaw = yaw_temp
            yaw = 0.2 * last + 0.8 * yaw
            last = yaw
            ref_pos[b, index] = ref; ref_yaw[b, index] = yaw
            if index == 0 and pos1 is not None:                              # :136-140
                far[b] = int(np.linalg.norm(ref - pos1[b]) > 1.0)
    return ref_pos, ref_yaw, far


def wrap_yaw(z):
    """updateFORCESResults' yaw wrap (nmpc_solver.cpp:531-541), with the reference's PI."""
    z = z.copy(); yaw = z[..., 16]
    z[..., 16] = np.where(yaw < -REF_PI, yaw + 2 * REF_PI, np.where(yaw > REF_PI, yaw - 2 * REF_PI, yaw))
    return z
