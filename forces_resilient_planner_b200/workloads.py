"""Synthetic NMPC workloads (BASELINE.json configs 1-5, restated in SURVEY.md §8d).

Everything here is plain numpy on the host: it only *describes* problems in the native batched
layout consumed by `nmpc_solve_batch*` (include/nmpc_b200.h):

    xinit [B, 9]            initial state (pos, vel, rpy)        -> params.xinit
    z0    [B, N, 17]        initial guess / warm start           -> params.x0
    hdr   [B, N, 10]        per stage [ref(3) f_ext(3) w_wp w_in w_rate yaw_ref]
                                                                  -> params.all_parameters[k*130 + 0..9]
    rows  [B, N, MCAP, 4]   corridor half-spaces (a0 a1 a2 b), a.pos <= b   (A | b of all_parameters)
    nrows [B, N] int32      live rows per stage (rows beyond are ignored)

Reference conventions reproduced (paths relative to
/root/reference/src/resilient_planner/plan_manage/):
  * cold start guess                       src/nmpc_solver.cpp:265-286 (thrust 7.3, state replicated)
  * weights                                launch/rotors_sim.launch:56-66 via src/forces_normal.cpp:36-52
  * yaw reference unwrap + 0.2/0.8 LPF     src/nmpc_solver.cpp:834-862
  * corridor tightening b_j - ||E a_j||    src/forces_normal.cpp:124-125
  * DecompROS local bounding box (2,2,1)   src/nmpc_solver.cpp:323,
                                           ThirdParty/.../decomp_util/line_segment.h:47-84
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SEED = 20260101
DT = 0.05
HOVER_THRUST_GUESS = 7.3          # nmpc_solver.cpp real_thrust_c_ default used in initMPCOutput
W_NORMAL = dict(stage=(7.0, 1.0, 80.0), terminal=(12.0, 0.5))
W_FINAL = dict(stage=(12.0, 1.5, 80.0), terminal=(15.0, 0.5))
EGO_E = np.array([0.27, 0.27, 0.0425])      # ego ellipsoid semi-axes (ego_r, ego_r, ego_h)


@dataclass
class Batch:
    xinit: np.ndarray
    z0: np.ndarray
    hdr: np.ndarray
    rows: np.ndarray
    nrows: np.ndarray
    variant: int = 0      # 0 = normal, 1 = final

    @property
    def B(self):
        return self.xinit.shape[0]

    @property
    def N(self):
        return self.z0.shape[1]

    @property
    def mcap(self):
        return self.rows.shape[2]

    def astype(self, dt):
        return Batch(self.xinit.astype(dt), self.z0.astype(dt), self.hdr.astype(dt),
                     self.rows.astype(dt), self.nrows.copy(), self.variant)

    def slice(self, lo, hi):
        return Batch(self.xinit[lo:hi].copy(), self.z0[lo:hi].copy(), self.hdr[lo:hi].copy(),
                     self.rows[lo:hi].copy(), self.nrows[lo:hi].copy(), self.variant)

    def algorithmic_bytes(self, itemsize=None):
        """Compulsory I/O of one solve in the native layout, per problem."""
        it = itemsize or self.xinit.dtype.itemsize
        n = self.N
        words_in = 9 + 17 * n + 10 * n + 4 * int(self.nrows.sum()) / self.B
        words_out = 17 * n + 17
        return (words_in + words_out) * it


def cold_start(xinit, n_stages):
    """z_k = [0,0,0,7.3, 0,0,0,7.3, x_odom] for all k  (nmpc_solver.cpp:272-276)."""
    B = xinit.shape[0]
    z0 = np.zeros((B, n_stages, 17))
    z0[:, :, 3] = HOVER_THRUST_GUESS
    z0[:, :, 7] = HOVER_THRUST_GUESS
    z0[:, :, 8:17] = xinit[:, None, :]
    return z0


def set_weights(hdr, variant=0):
    """setParasNormal/Final: stage weights everywhere, slots 6,7 overridden at the terminal stage."""
    w = W_FINAL if variant else W_NORMAL
    hdr[:, :, 6:9] = w["stage"]
    hdr[:, -1, 6] = w["terminal"][0]
    hdr[:, -1, 7] = w["terminal"][1]


def yaw_reference(yaw0, psi, n_stages):
    """calculate_yaw: unwrap psi to within pi of the running yaw, then y = 0.2*last + 0.8*y."""
    last = yaw0.copy()
    out = np.zeros((yaw0.shape[0], n_stages))
    for k in range(n_stages):
        y = psi.copy()
        far = np.abs(y - last) > np.pi
        y = np.where(far & (y > 0), y - 2 * np.pi, np.where(far, y + 2 * np.pi, y))
        y = 0.2 * last + 0.8 * y
        last = y
        out[:, k] = y
    return out


def tighten(A, b, E=EGO_E):
    """b_j - ||E a_j||_2 with E = diag(ego semi-axes)  (forces_normal.cpp:124-125)."""
    return b - np.linalg.norm(A * E, axis=-1)


def segment_box(p1, p2, bbox=(2.0, 2.0, 1.0)):
    """6 outward half-spaces of DecompROS' local bounding box around segment p1->p2.

    Returns A [B,6,3] (unit normals) and b [B,6] with A x <= b inside.
    """
    d = p2 - p1
    nrm = np.linalg.norm(d, axis=-1, keepdims=True)
    d = d / np.maximum(nrm, 1e-12)
    dh = np.stack([d[:, 1], -d[:, 0], np.zeros_like(d[:, 0])], -1)
    dhn = np.linalg.norm(dh, axis=-1, keepdims=True)
    dh = np.where(dhn > 0, dh / np.maximum(dhn, 1e-300), np.array([-1.0, 0.0, 0.0]))
    dv = np.cross(d, dh)
    normals = np.stack([dh, -dh, d, -d, dv, -dv], 1)
    pts = np.stack([p1 + dh * bbox[1], p1 - dh * bbox[1], p2 + d * bbox[0], p1 - d * bbox[0],
                    p1 + dv * bbox[2], p1 - dv * bbox[2]], 1)
    b = np.einsum("bjk,bjk->bj", normals, pts)
    return normals, b


def _pack_rows(A, b, n_stages, mcap, nrows=None):
    """Replicate one polytope per problem over all stages into the rows/nrows layout."""
    B, m, _ = A.shape
    rows = np.zeros((B, n_stages, mcap, 4))
    rows[:, :, :m, 0:3] = A[:, None]
    rows[:, :, :m, 3] = b[:, None]
    if nrows is None:
        nrows = np.full((B,), m, dtype=np.int32)
    nr = np.repeat(nrows[:, None].astype(np.int32), n_stages, 1)
    mask = np.arange(mcap)[None, None, :] >= nr[:, :, None]
    rows[mask] = 0.0
    return rows, nr


def config1(n_stages=20, variant=0):
    """Single-instance anchor: hover at (0,0,1), 1 m/s reference along +x, axis-aligned box."""
    xinit = np.array([[0, 0, 1.0, 0, 0, 0, 0, 0, 0]], dtype=np.float64)
    z0 = cold_start(xinit, n_stages)
    hdr = np.zeros((1, n_stages, 10))
    hdr[0, :, 0] = DT * (np.arange(n_stages) + 1)
    hdr[0, :, 1] = 0.0
    hdr[0, :, 2] = 1.0
    set_weights(hdr, variant)
    A = np.array([[[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1.0]]])
    b = np.array([[3.0, 2.0, 2.0, 2.0, 2.0, 0.0]])
    rows, nrows = _pack_rows(A, tighten(A, b), n_stages, 8)
    return Batch(xinit, z0, hdr, rows, nrows, variant)


def _random_core(B, n_stages, rng, fext=None):
    pos0 = np.stack([rng.uniform(-5, 5, B), rng.uniform(-5, 5, B), rng.uniform(0.8, 1.6, B)], -1)
    vel0 = rng.uniform(-1, 1, (B, 3))
    rp = rng.uniform(-0.2, 0.2, (B, 2))
    yaw0 = rng.uniform(-np.pi, np.pi, B)
    psi = rng.uniform(-np.pi, np.pi, B)
    speed = rng.uniform(0.5, 1.5, B)
    if fext is None:
        fext = rng.uniform(-2, 2, (B, 3))
    xinit = np.concatenate([pos0, vel0, rp, yaw0[:, None]], -1)
    heading = np.stack([np.cos(psi), np.sin(psi), np.zeros(B)], -1)
    k = (np.arange(n_stages) + 1)[None, :, None]
    ref = pos0[:, None, :] + speed[:, None, None] * DT * k * heading[:, None, :]
    hdr = np.zeros((B, n_stages, 10))
    hdr[:, :, 0:3] = ref
    hdr[:, :, 3:6] = fext[:, None, :]
    hdr[:, :, 9] = yaw_reference(yaw0, psi, n_stages)
    return xinit, hdr, pos0, ref


def config2(B=4096, n_stages=20, seed=SEED, variant=0, mcap=8, fext=None):
    """batch of identical-topology problems, randomised x0/goal/f_ext, 6-plane corridors."""
    rng = np.random.Generator(np.random.PCG64(seed))
    xinit, hdr, pos0, ref = _random_core(B, n_stages, rng, fext)
    set_weights(hdr, variant)
    A, b = segment_box(pos0, ref[:, -1, :])
    rows, nrows = _pack_rows(A, tighten(A, b), n_stages, mcap)
    return Batch(xinit, cold_start(xinit, n_stages), hdr, rows, nrows, variant)


def config3(B=65536, n_stages=20, seed=SEED + 3, variant=0, mcap=12):
    """as config 2 plus per-problem row count m ~ U{4..10} (divergent constraint rows)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    xinit, hdr, pos0, ref = _random_core(B, n_stages, rng)
    set_weights(hdr, variant)
    A6, b6 = segment_box(pos0, ref[:, -1, :])
    m = rng.integers(4, 11, B)
    A = np.zeros((B, 10, 3))
    b = np.zeros((B, 10))
    # box planes in random order, so m<6 drops (6-m) random planes
    perm = np.argsort(rng.random((B, 6)), axis=1)
    A[:, :6] = np.take_along_axis(A6, perm[:, :, None], 1)
    b[:, :6] = np.take_along_axis(b6, perm, 1)
    # extra planes: random unit normal, U(0.4,1.5) m (after tightening) beyond the farthest of
    # {start, end of the reference segment, where the initial velocity carries the vehicle in 0.5 s}
    nrm = rng.normal(size=(B, 4, 3))
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    dist = rng.uniform(0.4, 1.5, (B, 4))
    s1 = np.einsum("bjk,bk->bj", nrm, pos0)
    s2 = np.einsum("bjk,bk->bj", nrm, ref[:, -1, :])
    s3 = np.einsum("bjk,bk->bj", nrm, pos0 + 0.5 * xinit[:, 3:6])
    A[:, 6:] = nrm
    b[:, 6:] = np.maximum(np.maximum(s1, s2), s3) + dist + np.linalg.norm(nrm * EGO_E, axis=-1)
    rows, nrows = _pack_rows(A, tighten(A, b), n_stages, mcap, m)
    return Batch(xinit, cold_start(xinit, n_stages), hdr, rows, nrows, variant)


def config4(side=512, n_stages=40, seed=SEED + 4, variant=0, mcap=8, lo=0, hi=None):
    """N=40 long horizon, constant-wind sweep: |f| = linspace(0,4,side) x azimuth linspace(0,2pi,side)."""
    B = side * side
    mag = np.repeat(np.linspace(0, 4, side), side)
    az = np.tile(np.linspace(0, 2 * np.pi, side, endpoint=False), side)
    fext = np.stack([mag * np.cos(az), mag * np.sin(az), np.zeros(B)], -1)
    batch = config2(B, n_stages, seed, variant, mcap, fext=fext)
    return batch if hi is None and lo == 0 else batch.slice(lo, hi if hi is not None else B)


def to_forces_params(batch: Batch, i: int):
    """Problem i in the reference ABI layout: (xinit[9], x0[N*17], all_parameters[N*130])."""
    n = batch.N
    allp = np.zeros((n, 130))
    allp[:, 0:10] = batch.hdr[i]
    for k in range(n):
        m = min(int(batch.nrows[i, k]), 30)
        allp[k, 10:10 + 3 * m] = batch.rows[i, k, :m, 0:3].reshape(-1)
        allp[k, 100:100 + m] = batch.rows[i, k, :m, 3]
    return batch.xinit[i].copy(), batch.z0[i].reshape(-1).copy(), allp.reshape(-1)


def shift_warm_start(z, xinit_next=None):
    """Reference receding-horizon shift: x0[k] <- z[k+1], last stage duplicated; xinit <- z[1][8:17].

    forces_normal.cpp:62-97 with mpc_output_.at(20) = mpc_output_.at(19) (nmpc_solver.cpp:543).
    """
    z0 = np.concatenate([z[:, 1:], z[:, -1:]], axis=1)
    xinit = z[:, 1, 8:17].copy() if xinit_next is None else xinit_next
    return xinit, z0
