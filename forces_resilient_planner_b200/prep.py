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


def adopt_plans(z_new, info_int, z_prev, accept=None, odom=None, wrap_yaw=True, cold=None, stream=None):
    """nmpc_adopt_plans_f64: accepted agents adopt z_new into z_prev (in place), rejected ones get the cold guess at
    their current state (NMPCSolver::solveNMPC result handling, nmpc_solver.cpp:398-427, 363-364).
    cold [B] (int32 cuda tensor, optional) receives 1 for the rejected agents; returned as given."""
    import torch
    lib = _lib.load()
    B, N, _ = z_prev.shape
    fn = lib.nmpc_adopt_plans_f64
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 6 + [ctypes.c_int, ctypes.c_void_p]
    st = stream if stream is not None else torch.cuda.current_stream(z_prev.device)
    with torch.cuda.device(z_prev.device):
        _check(fn(B, N, z_new.data_ptr(), info_int.data_ptr() if info_int is not None else None,
                  accept.data_ptr() if accept is not None else None, odom.data_ptr() if odom is not None else None,
                  z_prev.data_ptr(), cold.data_ptr() if cold is not None else None, int(bool(wrap_yaw)), st.cuda_stream))
    return cold


def rank_longest_first(info_int, order, stream=None):
    """nmpc_rank_longest_first: order [B] (int32, cuda) <- launch order of the next warm solve."""
    import torch
    lib = _lib.load()
    fn = lib.nmpc_rank_longest_first
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    st = stream if stream is not None else torch.cuda.current_stream(info_int.device)
    with torch.cuda.device(info_int.device):
        _check(fn(info_int.shape[0], info_int.data_ptr(), order.data_ptr(), st.cuda_stream))
    return order


def sample_reference(kino_path, kino_size, t_off, last_yaw, N, Ts, pos1=None, stream=None):
    """Batched getCurTraj + calculate_yaw (nmpc_solver.cpp:109-142, 834-862) on the device.

    kino_path [B,P,3], kino_size [B] (int32), t_off [B], last_yaw [B], pos1 [B,3] or None (cuda tensors)
    -> ref_pos [B,N,3], ref_yaw [B,N], hard_to_follow [B] (int32)."""
    import torch
    lib = _lib.load()
    B, P, _ = kino_path.shape
    dev = kino_path.device
    ref_pos = torch.empty((B, N, 3), dtype=torch.float64, device=dev)
    ref_yaw = torch.empty((B, N), dtype=torch.float64, device=dev)
    far = torch.zeros((B,), dtype=torch.int32, device=dev)
    fn = lib.nmpc_sample_reference_f64
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double] + [ctypes.c_void_p] * 9
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _check(fn(B, int(N), P, float(Ts), kino_path.data_ptr(), kino_size.data_ptr(), t_off.data_ptr(),
                  last_yaw.data_ptr(), pos1.data_ptr() if pos1 is not None else None, ref_pos.data_ptr(),
                  ref_yaw.data_ptr(), far.data_ptr(), st.cuda_stream))
    return ref_pos, ref_yaw, far


def propagate_ellipsoids(z, consts=None, out=None, stream=None):
    """Disturbance-ellipsoid propagation along the previous plan (setFORCESParams / updateMatrix /
    getDistrEllipsoid, nmpc_solver.cpp:484-521, 567-699) on the device: z [B,N,17] (cuda, float64)
    -> ellipsoid [B,N,9], the E_i that pack_params takes.  `consts`: dict of nmpc_ellipsoid_consts overrides."""
    import torch
    lib = _lib.load()
    B, N, _ = z.shape
    c = _lib.EllipsoidConsts()
    lib.nmpc_default_ellipsoid_consts(ctypes.byref(c))
    for k, v in (consts or {}).items():
        if not hasattr(c, k):
            raise AttributeError(f"nmpc_ellipsoid_consts has no field {k!r}")
        setattr(c, k, float(v))
    out = torch.empty((B, N, 9), dtype=torch.float64, device=z.device) if out is None else out
    fn = lib.nmpc_propagate_ellipsoids_f64
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(_lib.EllipsoidConsts), ctypes.c_void_p,
                   ctypes.c_void_p]
    st = stream if stream is not None else torch.cuda.current_stream(z.device)
    with torch.cuda.device(z.device):
        _check(fn(B, N, z.data_ptr(), ctypes.byref(c), out.data_ptr(), st.cuda_stream))
    return out


def select_corridors(cloud, cloud_n, ref_pos, ref_yaw, ellipsoid, max_polys=8, max_rows=30, bbox=(2.0, 2.0, 1.0),
                     stream=None):
    """Corridor generation + per-stage polytope selection on the device (getSikangConst / setFORCESParams,
    nmpc_solver.cpp:288-332, 493-516; DecompROS EllipsoidDecomp::dilate).

    cloud [B,M,3] per agent or [M,3] shared (cuda, float64), cloud_n [B] / [1] (int32), ref_pos [B,N,3],
    ref_yaw [B,N], ellipsoid [B,N,9]  ->  poly_A [B,P,R,3], poly_b [B,P,R], poly_m [B,P], poly_idx [B,N],
    n_poly [B], overflow [B]."""
    import torch
    lib = _lib.load()
    B, N, _ = ref_pos.shape
    shared = cloud.dim() == 2
    M = cloud.shape[0] if shared else cloud.shape[1]
    dev = ref_pos.device
    P, R = int(max_polys), int(max_rows)
    poly_A = torch.zeros((B, P, R, 3), dtype=torch.float64, device=dev)      # unused polytope slots stay zero
    poly_b = torch.zeros((B, P, R), dtype=torch.float64, device=dev)
    poly_m = torch.empty((B, P), dtype=torch.int32, device=dev)
    poly_idx = torch.empty((B, N), dtype=torch.int32, device=dev)
    n_poly = torch.empty((B,), dtype=torch.int32, device=dev)
    overflow = torch.empty((B,), dtype=torch.int32, device=dev)
    fn = lib.nmpc_select_corridors_f64
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p, ctypes.c_longlong] + [ctypes.c_void_p] * 4 + \
        [ctypes.POINTER(ctypes.c_double)] + [ctypes.c_void_p] * 7
    bb = (ctypes.c_double * 3)(*[float(x) for x in bbox])
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _check(fn(B, N, M, P, R, cloud.data_ptr(), 0 if shared else 3 * M, cloud_n.data_ptr(), ref_pos.data_ptr(),
                  ref_yaw.data_ptr(), ellipsoid.data_ptr(), bb, poly_A.data_ptr(), poly_b.data_ptr(), poly_m.data_ptr(),
                  poly_idx.data_ptr(), n_poly.data_ptr(), overflow.data_ptr(), st.cuda_stream))
    return poly_A, poly_b, poly_m, poly_idx, n_poly, overflow


def wrap_yaw(z, stream=None):
    """updateFORCESResults' yaw wrap (nmpc_solver.cpp:531-541) on an adopted plan z [B,N,17] (cuda, float64), in place."""
    import torch
    lib = _lib.load()
    B, N, _ = z.shape
    fn = lib.nmpc_wrap_yaw_f64
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    st = stream if stream is not None else torch.cuda.current_stream(z.device)
    with torch.cuda.device(z.device):
        _check(fn(B, N, z.data_ptr(), st.cuda_stream))
    return z
