"""Stand-alone structured KKT factorisation / backsolve (nmpc_riccati_factor_*, nmpc_kkt_backsolve_*).

torch tensors are device buffers only.  Layouts: see include/nmpc_b200.h.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

PHI_WORDS, JC_WORDS = 21, 51


def factor_words() -> int:
    return int(_lib.load().nmpc_backsolve_factor_words())


def algorithmic_bytes(N: int, itemsize: int) -> int:
    lib = _lib.load()
    lib.nmpc_backsolve_algorithmic_bytes.restype = ctypes.c_long
    return int(lib.nmpc_backsolve_algorithmic_bytes(int(N), int(itemsize)))


def _sfx(t):
    import torch
    return "f64" if t.dtype == torch.float64 else "f32"


def _check(rc):
    if rc != 0:
        raise RuntimeError(f"nmpc_b200 call failed (rc={rc}): {_lib.last_error()}")


def riccati_factor(phi, jc, stream=None):
    """phi [B,N,21], jc [B,N,51] (cuda tensors) -> fac [B,N,204] (opaque: N*204 words per problem, see
    include/nmpc_b200.h), status [B] (int32)."""
    import torch
    lib = _lib.load()
    B, N, _ = phi.shape
    fac = torch.empty((B, N, factor_words()), dtype=phi.dtype, device=phi.device)
    status = torch.empty((B,), dtype=torch.int32, device=phi.device)
    fn = getattr(lib, f"nmpc_riccati_factor_{_sfx(phi)}")
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 5
    st = stream if stream is not None else torch.cuda.current_stream(phi.device)
    with torch.cuda.device(phi.device):
        _check(fn(B, N, phi.data_ptr(), jc.data_ptr(), fac.data_ptr(), status.data_ptr(), st.cuda_stream))
    return fac, status


def kkt_backsolve(fac, g, d, dz=None, y=None, stream=None):
    """fac [B,N,204] (as produced by riccati_factor), g [B,N,17], d [B,N,13] -> dz [B,N,17], y [B,N,13]."""
    import torch
    lib = _lib.load()
    B, N, _ = fac.shape
    dz = torch.empty((B, N, 17), dtype=fac.dtype, device=fac.device) if dz is None else dz
    y = torch.empty((B, N, 13), dtype=fac.dtype, device=fac.device) if y is None else y
    fn = getattr(lib, f"nmpc_kkt_backsolve_{_sfx(fac)}")
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 6
    st = stream if stream is not None else torch.cuda.current_stream(fac.device)
    with torch.cuda.device(fac.device):
        _check(fn(B, N, fac.data_ptr(), g.data_ptr(), d.data_ptr(), dz.data_ptr(), y.data_ptr(), st.cuda_stream))
    return dz, y


# ---- host-side helpers shared by tests and bench (problem construction only, no solving) ------
def dense_from_compact(phi: np.ndarray, jc: np.ndarray):
    """Expand one problem's compact blocks: phi [N,21], jc [N,51] -> Phi [N,17,17], C [N-1,13,17]."""
    N = phi.shape[0]
    Phi = np.zeros((N, 17, 17))
    C = np.zeros((N - 1, 13, 17))
    h = 0.05
    for k in range(N):
        Phi[k][np.arange(17), np.arange(17)] = phi[k, :17]
        for (a, b), v in zip(((8, 9), (8, 10), (9, 10)), phi[k, 17:20]):
            Phi[k, a, b] = Phi[k, b, a] = v
        for i in range(4):
            Phi[k, i, 4 + i] = Phi[k, 4 + i, i] = phi[k, 20]
        if k < N - 1:
            j = jc[k]
            C[k, 0:3, 8:11] = np.eye(3)
            C[k, 0:3, 11:14] = j[0:9].reshape(3, 3)
            C[k, 0:3, 14:17] = j[9:18].reshape(3, 3)
            C[k, 0:3, 3] = j[18:21]
            C[k, 3:6, 11:14] = j[21:30].reshape(3, 3)
            C[k, 3:6, 14:17] = j[30:39].reshape(3, 3)
            C[k, 3:6, 3] = j[39:42]
            C[k, 3:6, 0:3] = j[42:51].reshape(3, 3)
            C[k, 6:9, 14:17] = np.eye(3)
            C[k, 6:9, 0:3] = h * np.eye(3)
            C[k, 9:13, 0:4] = np.eye(4)
    return Phi, C


def random_kkt_problems(B: int, N: int, seed: int = 0):
    """Well-conditioned synthetic KKT systems in the compact layout (numpy, fp64)."""
    rng = np.random.default_rng(seed)
    phi = np.zeros((B, N, PHI_WORDS))
    jc = np.zeros((B, N, JC_WORDS))
    phi[:, :, :17] = rng.uniform(1.0, 30.0, (B, N, 17))
    phi[:, :, 0:8] += 160.0                       # 2 w_rate on the u / u_prev diagonal
    phi[:, :, 20] = -160.0                        # H[u_i][uprev_i] = -2 w_rate
    v = rng.normal(size=(B, N, 3)) * 2.0          # pos block: diag += v v', off-diag = v_a v_b (PSD)
    phi[:, :, 8:11] += v ** 2
    phi[:, :, 17] = v[..., 0] * v[..., 1]
    phi[:, :, 18] = v[..., 0] * v[..., 2]
    phi[:, :, 19] = v[..., 1] * v[..., 2]
    I3 = np.eye(3).reshape(-1)
    jc[:, :, 0:9] = 0.05 * I3 + 1e-3 * rng.normal(size=(B, N, 9))
    jc[:, :, 9:18] = 0.02 * rng.normal(size=(B, N, 9))
    jc[:, :, 18:21] = 2e-3 * rng.normal(size=(B, N, 3))
    jc[:, :, 21:30] = I3 + 0.02 * rng.normal(size=(B, N, 9))
    jc[:, :, 30:39] = 0.5 * rng.normal(size=(B, N, 9))
    jc[:, :, 39:42] = 0.07 * rng.normal(size=(B, N, 3))
    jc[:, :, 42:51] = 0.02 * rng.normal(size=(B, N, 9))
    g = rng.normal(size=(B, N, 17)) * 10.0
    d = rng.normal(size=(B, N, 13)) * 0.05
    return phi, jc, g, d
