"""Host-side Python mirror of the batched solver interface.

torch is used for device memory and streams only (plumbing); all arithmetic happens inside
libnmpc_b200.so.  Three ways in, all ending in the same CUDA kernel:

  * `DeviceBatch` + `solve_device`   inputs resident in HBM (what `bench.py` times as `value`)
  * `solve_host`                     host numpy arrays, H2D + solve + D2H inside the C call
                                     (what `bench.py` times as `e2e`)
    float32 arrays select the mixed-precision kernel (nmpc_solve_batch_f32: single-precision Newton
    system, double-precision iterate and residuals, the reference tolerances); `mixed=True` selects it
    for float64 arrays (nmpc_solve_batch_mixed_f64)
  * `forces.FORCESNormal/FORCESFinal` the reference wrapper classes over the reference ABI
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import _lib
from .workloads import Batch

EXIT_OPTIMAL, EXIT_MAXIT, EXIT_FACTOR, EXIT_BADFUNC, EXIT_NOPROGRESS = 1, 0, -5, -6, -7


@dataclass
class Result:
    z: np.ndarray          # [B, N, 17]
    flag: np.ndarray       # [B] reference exit codes (1 optimal, 0 maxit, -5/-6/-7 failures)
    it: np.ndarray         # [B] iterations
    nbt: np.ndarray        # [B] backtracking steps
    info_real: np.ndarray  # [B, 8] res_eq res_ineq rsnorm rcompnorm pobj mu alpha_p alpha_d
    resolved: np.ndarray | None = None   # [B] 1 where the mixed-precision kernel handed the problem to the fp64 kernel


def _check(rc: int):
    if rc != 0:
        raise RuntimeError(f"nmpc_b200 call failed (rc={rc}): {_lib.last_error()}")


def _np_ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def solve_host(batch: Batch, dtype=np.float64, opts: _lib.NmpcOpts | None = None, mixed: bool = False) -> Result:
    """Host buffers in, host buffers out: nmpc_solve_batch_host_{f64,f32,mixed_f64}."""
    lib = _lib.load()
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        fn = lib.nmpc_solve_batch_host_mixed_f64 if mixed else lib.nmpc_solve_batch_host_f64
    else:
        fn = lib.nmpc_solve_batch_host_f32
    B, N, mcap = batch.B, batch.N, batch.mcap
    xinit = np.ascontiguousarray(batch.xinit, dtype)
    z0 = np.ascontiguousarray(batch.z0, dtype)
    hdr = np.ascontiguousarray(batch.hdr, dtype)
    rows = np.ascontiguousarray(batch.rows, dtype)
    nrows = np.ascontiguousarray(batch.nrows, np.int32)
    z = np.empty((B, N, 17), dtype)
    ii = np.empty((B, 4), np.int32)
    ir = np.empty((B, 8), dtype)
    o = opts or _lib.default_opts()
    _check(fn(B, N, mcap, _np_ptr(xinit), _np_ptr(z0), _np_ptr(hdr), _np_ptr(rows), _np_ptr(nrows),
              int(batch.variant), ctypes.byref(o), _np_ptr(z), _np_ptr(ii), _np_ptr(ir)))
    return Result(z, ii[:, 0].copy(), ii[:, 1].copy(), ii[:, 2].copy(), ir, ii[:, 3].copy())


class DeviceBatch:
    """A batch of problems resident in HBM (torch tensors are just the allocator here)."""

    def __init__(self, batch: Batch, dtype=np.float64, device="cuda:0", pinned: bool = False):
        import torch
        self.torch = torch
        self.np_dtype = np.dtype(dtype)
        self.t_dtype = torch.float64 if self.np_dtype == np.float64 else torch.float32
        self.device = torch.device(device)
        self.B, self.N, self.mcap, self.variant = batch.B, batch.N, batch.mcap, int(batch.variant)
        mk = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dt))
        self.h = dict(xinit=mk(batch.xinit, self.np_dtype), z0=mk(batch.z0, self.np_dtype),
                      hdr=mk(batch.hdr, self.np_dtype), rows=mk(batch.rows, self.np_dtype),
                      nrows=mk(batch.nrows, np.int32))
        if pinned:
            self.h = {k: v.pin_memory() for k, v in self.h.items()}
        self.d = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in self.h.items()}
        self.z = torch.empty((self.B, self.N, 17), dtype=self.t_dtype, device=self.device)
        self.info_int = torch.empty((self.B, 4), dtype=torch.int32, device=self.device)
        self.info_real = torch.empty((self.B, 8), dtype=self.t_dtype, device=self.device)
        self.upload()

    def upload(self, non_blocking: bool = True):
        for k in self.h:
            self.d[k].copy_(self.h[k], non_blocking=non_blocking)

    @property
    def h2d_bytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.h.values())

    @property
    def d2h_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.z, self.info_int, self.info_real))

    def result(self) -> Result:
        self.torch.cuda.synchronize(self.device)
        ii = self.info_int.cpu().numpy()
        return Result(self.z.cpu().numpy(), ii[:, 0].copy(), ii[:, 1].copy(), ii[:, 2].copy(),
                      self.info_real.cpu().numpy(), ii[:, 3].copy())


def solve_device(db: DeviceBatch, opts: _lib.NmpcOpts | None = None, stream=None, mixed: bool = False,
                 lowlatency: bool = False) -> None:
    """Enqueue one solve on `stream` (default: torch's current stream): the fp64 kernel for float64 batches, the
    mixed-precision kernel (+ its fp64 re-solve of the few problems it gives up on) for float32 batches or `mixed`;
    `lowlatency` (float64 batches): the warp-group variant of the mixed-precision kernel (nmpc_solve_batch_lowlatency_f64)."""
    lib = _lib.load()
    torch = db.torch
    st = stream if stream is not None else torch.cuda.current_stream(db.device)
    o = opts or _lib.default_opts()
    d = db.d
    args = [db.B, db.N, db.mcap, d["xinit"].data_ptr(), d["z0"].data_ptr(), d["hdr"].data_ptr(),
            d["rows"].data_ptr(), d["nrows"].data_ptr(), db.variant, ctypes.byref(o),
            db.z.data_ptr(), db.info_int.data_ptr(), db.info_real.data_ptr()]
    with torch.cuda.device(db.device):
        if db.np_dtype == np.float32:
            _check(lib.nmpc_solve_batch_f32(*args, ctypes.c_void_p(st.cuda_stream)))
        elif lowlatency:
            _check(lib.nmpc_solve_batch_lowlatency_f64(*args, None, None, None, None, None, ctypes.c_void_p(st.cuda_stream)))
        elif mixed:
            _check(lib.nmpc_solve_batch_mixed_f64(*args, None, None, None, None, None, ctypes.c_void_p(st.cuda_stream)))
        else:
            _check(lib.nmpc_solve_batch_f64(*args, ctypes.c_void_p(st.cuda_stream)))


def solve(batch: Batch, dtype=np.float64, opts=None, device="cuda:0", mixed: bool = False, lowlatency: bool = False) -> Result:
    """Convenience: upload, solve on the device, download."""
    db = DeviceBatch(batch, dtype, device)
    solve_device(db, opts, mixed=mixed, lowlatency=lowlatency)
    return db.result()


def solve_with_multipliers(batch: Batch, opts=None, device="cuda:0", mixed: bool = False, lowlatency: bool = False):
    """fp64-array solve that also returns the multipliers of the KKT point (nmpc_solve_batch_ex_f64, or
    nmpc_solve_batch_mixed_f64 with `mixed`).

    Returns (Result, dict(y, zl, zu, lc)) -- what the KKT-acceptance tests need to re-evaluate
    ForcesPro's stopping test with the reference callbacks."""
    lib = _lib.load()
    db = DeviceBatch(batch, np.float64, device)
    torch = db.torch
    mk = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=db.device)
    y, zl, zu = mk(db.B, db.N, 13), mk(db.B, db.N, 17), mk(db.B, db.N, 17)
    lc = mk(db.B, db.N, max(db.mcap, 1))
    o = opts or _lib.default_opts()
    d = db.d
    with torch.cuda.device(db.device):
        st = torch.cuda.current_stream(db.device)
        args = [db.B, db.N, db.mcap, d["xinit"].data_ptr(), d["z0"].data_ptr(), d["hdr"].data_ptr(),
                d["rows"].data_ptr(), d["nrows"].data_ptr(), db.variant, ctypes.byref(o),
                db.z.data_ptr(), db.info_int.data_ptr(), db.info_real.data_ptr(),
                y.data_ptr(), zl.data_ptr(), zu.data_ptr(), lc.data_ptr() if db.mcap else None]
        if lowlatency:
            _check(lib.nmpc_solve_batch_lowlatency_f64(*args, None, ctypes.c_void_p(st.cuda_stream)))
        elif mixed:
            _check(lib.nmpc_solve_batch_mixed_f64(*args, None, ctypes.c_void_p(st.cuda_stream)))
        else:
            _check(lib.nmpc_solve_batch_ex_f64(*args, ctypes.c_void_p(st.cuda_stream)))
    res = db.result()
    return res, dict(y=y.cpu().numpy(), zl=zl.cpu().numpy(), zu=zu.cpu().numpy(),
                     lc=lc.cpu().numpy()[:, :, :db.mcap])


def model_eval_device(z, p, stage, n_stages=20, variant=0):
    """Device model probe (nmpc_model_eval_host_f64): same outputs as the reference callback."""
    lib = _lib.load()
    z = np.ascontiguousarray(z, np.float64).reshape(-1, 17)
    p = np.ascontiguousarray(p, np.float64).reshape(-1, 130)
    st = np.ascontiguousarray(stage, np.int32).reshape(-1)
    n = z.shape[0]
    f = np.zeros(n); g = np.zeros((n, 17)); c = np.zeros((n, 13)); jc = np.zeros((n, 221))
    h = np.zeros((n, 30)); jh = np.zeros((n, 510))
    lib.nmpc_model_eval_host_f64.restype = ctypes.c_int
    _check(lib.nmpc_model_eval_host_f64(n, _np_ptr(z), _np_ptr(p), _np_ptr(st), int(n_stages), int(variant),
                                        _np_ptr(f), _np_ptr(g), _np_ptr(c), _np_ptr(jc), _np_ptr(h), _np_ptr(jh)))
    return dict(f=f, grad=g, c=c, jc=jc.reshape(n, 17, 13).transpose(0, 2, 1).copy(), h=h,
                jh=jh.reshape(n, 17, 30).transpose(0, 2, 1).copy())
