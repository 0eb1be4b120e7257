"""Device-side parameter packing and warm-start shift (nmpc_pack_params_f64, nmpc_shift_warm_start_f64):
the batched, GPU-resident form of FORCESNormal::solveNormal's packing loop
(/root/reference/src/resilient_planner/plan_manage/src/forces_normal.cpp:55-136) and of the
receding-horizon shift (nmpc_solver.cpp:531-543)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .solver import _check


def pack_params(ref_pos, ref_yaw, ext_acc, ellipsoid, poly_A, poly_b, poly_m, poly_idx, weights5, mcap,
                stream=None):
    """All tensor arguments are cuda tensors (float64 / int32); returns hdr, rows, nrows (cuda)."""
    import torch
    lib = _lib.load()
    B, N, _ = ref_pos.shape
    P, M = poly_A.shape[1], poly_A.shape[2]
    dev = ref_pos.device
    hdr = torch.empty((B, N, 10), dtype=torch.float64, device=dev)
    rows = torch.empty((B, N, mcap, 4), dtype=torch.float64, device=dev)
    nrows = torch.empty((B, N), dtype=torch.int32, device=dev)
    w = (ctypes.c_double * 5)(*[float(x) for x in weights5])
    fn = lib.nmpc_pack_params_f64
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p] * 8 + [ctypes.POINTER(ctypes.c_double)] + \
        [ctypes.c_void_p] * 4
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _check(fn(B, N, P, M, mcap, ref_pos.data_ptr(), ref_yaw.data_ptr(), ext_acc.data_ptr(),
                  ellipsoid.data_ptr(), poly_A.data_ptr(), poly_b.data_ptr(), poly_m.data_ptr(),
                  poly_idx.data_ptr(), w, hdr.data_ptr(), rows.data_ptr(), nrows.data_ptr(), st.cuda_stream))
    return hdr, rows, nrows


def shift_warm_start(z_prev, xinit=None, z0=None, wrap_yaw=True, stream=None):
    """z_prev [B,N,17] (cuda, float64) -> (xinit [B,9], z0 [B,N,17])."""
    import torch
    lib = _lib.load()
    B, N, _ = z_prev.shape
    xinit = torch.empty((B, 9), dtype=torch.float64, device=z_prev.device) if xinit is None else xinit
    z0 = torch.empty_like(z_prev) if z0 is None else z0
    fn = lib.nmpc_shift_warm_start_f64
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                   ctypes.c_int, ctypes.c_void_p]
    st = stream if stream is not None else torch.cuda.current_stream(z_prev.device)
    with torch.cuda.device(z_prev.device):
        _check(fn(B, N, z_prev.data_ptr(), xinit.data_ptr(), z0.data_ptr(), int(bool(wrap_yaw)), st.cuda_stream))
    return xinit, z0


def pack_params_reference(ref_pos, ref_yaw, ext_acc, ellipsoid, poly_A, poly_b, poly_m, poly_idx, weights5, mcap):
    """numpy restatement of the same packing loop (host only; the tests' checker for the kernel)."""
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
