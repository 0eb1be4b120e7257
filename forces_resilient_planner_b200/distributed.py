"""Multi-GPU driver: contiguous sharding of independent problems, one process per GPU.

The NMPC instances never interact (single-vehicle planner), so the solve itself needs NO collective:
rank r owns problems [r*per, (r+1)*per) with per = ceil(B / world) and runs the same fused kernel on them.  The only
exchange is the end-of-batch collation of the results, and it is copy-free: every rank's kernel writes its z straight
into its slice of ONE buffer of world*per problems and an in-place all-gather (sendbuff = recvbuff + rank*count) on the
same stream completes the other slices -- `nmpc_solve_batch_sharded_{f64,f32}` of the C ABI (include/nmpc_b200.h), which
drives NCCL itself.  torch.distributed is used for one thing only: handing rank 0's 128-byte NCCL id to the other ranks.

`PeerCollator` goes one step further on the GPUs of one node (`nmpc_solve_batch_sharded_p2p_*`): the solve kernel's epilogue
writes every result into all ranks' buffers over NVLink (CUDA IPC mappings), so the exchange overlaps the arithmetic and
only a barrier kernel follows the solve.

`TorchCollator` is the same in-place gather through torch.distributed (any backend): it exists so that the host logic
(shard arithmetic, slice placement, trimming of the padded tail) is testable with `gloo` on CPU, world size 2.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .solver import _check
from .workloads import Batch


def per_rank(B: int, world: int) -> int:
    """Problems per rank: ceil(B / world), rounded up to even so that every rank's slice stays 16-byte aligned
    (TMA bulk stores of the results) in fp32 and fp64 alike."""
    per = -(-B // world)
    return per + (per & 1)


def shard_range(B: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous split in equal blocks of per_rank(B, world); the last rank(s) may be short (or empty)."""
    per = per_rank(B, world)
    lo = min(rank * per, B)
    return lo, min(lo + per, B)


def shard(batch: Batch, rank: int, world: int, pad: bool = False) -> Batch:
    """This rank's problems.  pad=True repeats the last problem up to per_rank(B, world), so that every rank launches
    the same grid and owns an equally sized slice of the collation buffer; the padding rows land beyond index B of the
    gathered result and are never looked at."""
    lo, hi = shard_range(batch.B, rank, world)
    if hi == lo:
        lo, hi = batch.B - 1, batch.B                      # a rank with nothing to do solves a copy of the last problem
    mine = batch.slice(lo, hi)
    if pad:
        need = per_rank(batch.B, world) - mine.B
        if need > 0:
            rep = lambda a: np.concatenate([a, np.repeat(a[-1:], need, axis=0)], axis=0)
            mine = Batch(rep(mine.xinit), rep(mine.z0), rep(mine.hdr), rep(mine.rows), rep(mine.nrows), mine.variant)
    return mine


class NcclCollator:
    """nmpc_comm of the C ABI: NCCL communicator + NCCL-registered result buffers.  One per process (= per GPU)."""

    def __init__(self, rank: int, world: int, device, uid: bytes | None = None, group=None):
        import torch
        self.torch, self.lib = torch, _lib.load()
        self.rank, self.world, self.device = rank, world, torch.device(device)
        lib = self.lib
        lib.nmpc_comm_unique_id.argtypes = [ctypes.c_char_p]
        lib.nmpc_comm_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
        lib.nmpc_comm_destroy.argtypes = [ctypes.c_void_p]
        lib.nmpc_comm_alloc.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]
        lib.nmpc_comm_free.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.nmpc_collate_inplace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        sig = [ctypes.c_void_p] + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.POINTER(_lib.NmpcOpts)] + \
            [ctypes.c_void_p] * 3
        lib.nmpc_solve_batch_sharded_f64.argtypes = sig + [ctypes.c_int, ctypes.c_void_p]
        lib.nmpc_solve_batch_sharded_f32.argtypes = sig + [ctypes.c_void_p]
        if uid is None:
            # the one use of torch.distributed: rank 0's NCCL id to everybody (a 128-byte broadcast over the existing group)
            import torch.distributed as dist
            buf = ctypes.create_string_buffer(128)
            if rank == 0:
                _check(lib.nmpc_comm_unique_id(buf))
            box = [buf.raw]
            dist.broadcast_object_list(box, src=0, group=group)
            uid = box[0]
        self.comm = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _check(lib.nmpc_comm_create(world, rank, uid, ctypes.byref(self.comm)))
        self._bufs = []

    def alloc(self, shape, dtype):
        """A device tensor over NCCL-allocated, NCCL-registered memory (ncclMemAlloc + ncclCommRegister)."""
        torch = self.torch
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        ptr = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _check(self.lib.nmpc_comm_alloc(self.comm, n, ctypes.byref(ptr)))
        self._bufs.append(ptr)
        return _wrap_device_memory(torch, ptr.value, shape, dtype, self.device, self)   # keeps the communicator alive

    def solve_sharded(self, db, z_all, info_int_all, opts=None, mixed: bool = False, stream=None):
        """Enqueue solve + in-place all-gather for this rank's DeviceBatch `db` (every rank: the same db.B)."""
        torch = self.torch
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        o = opts or _lib.default_opts()
        d = db.d
        args = [self.comm, db.B, db.N, db.mcap, d["xinit"].data_ptr(), d["z0"].data_ptr(), d["hdr"].data_ptr(),
                d["rows"].data_ptr(), d["nrows"].data_ptr(), db.variant, ctypes.byref(o),
                z_all.data_ptr(), info_int_all.data_ptr(), db.info_real.data_ptr()]
        with torch.cuda.device(self.device):
            if db.np_dtype == np.float32:
                _check(self.lib.nmpc_solve_batch_sharded_f32(*args, ctypes.c_void_p(st.cuda_stream)))
            else:
                _check(self.lib.nmpc_solve_batch_sharded_f64(*args, int(bool(mixed)), ctypes.c_void_p(st.cuda_stream)))

    def collate(self, buf_all, stream=None):
        torch = self.torch
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        per_bytes = buf_all.numel() * buf_all.element_size() // self.world
        with torch.cuda.device(self.device):
            _check(self.lib.nmpc_collate_inplace(self.comm, buf_all.data_ptr(), per_bytes, ctypes.c_void_p(st.cuda_stream)))

    def close(self):
        if self.comm:
            self.torch.cuda.synchronize(self.device)
            self.lib.nmpc_comm_destroy(self.comm)
            self.comm = ctypes.c_void_p()


def _wrap_device_memory(torch, ptr: int, shape, dtype, device, owner):
    """A torch tensor over device memory this library owns (no copy; `owner` is kept alive by the tensor)."""
    np_dt = {torch.float64: np.float64, torch.float32: np.float32, torch.int32: np.int32}[dtype]

    class _Mem:      # __cuda_array_interface__
        pass
    m = _Mem()
    m.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=np.dtype(np_dt).str, data=(int(ptr), False), version=3)
    t = torch.as_tensor(m, device=device)
    t._nmpc_owner = owner
    return t


class PeerCollator:
    """nmpc_peers of the C ABI: collation by peer stores from inside the solve kernel (GPUs of one node).

    Every rank holds z_all [world * per][N][17] and info_all [world * per][4]; the solve kernel of rank r writes its
    results into slice r of every rank's copy, two barrier kernels bracket the solve.  torch.distributed (any backend)
    is used once, to exchange the 64-byte CUDA IPC handles."""

    def __init__(self, rank: int, world: int, device, per: int, N: int, dtype=np.float64, group=None, handles=None):
        import torch
        self.torch, self.lib = torch, _lib.load()
        self.rank, self.world, self.device = rank, world, torch.device(device)
        self.per, self.N, self.np_dtype = per, N, np.dtype(dtype)
        lib = self.lib
        vp, i = ctypes.c_void_p, ctypes.c_int
        lib.nmpc_peers_create.argtypes = [i, i, ctypes.c_size_t, ctypes.c_size_t, ctypes.POINTER(vp)]
        lib.nmpc_peers_export.argtypes = [vp, ctypes.c_char_p]
        lib.nmpc_peers_connect.argtypes = [vp, ctypes.c_char_p]
        lib.nmpc_peers_z.argtypes = [vp]; lib.nmpc_peers_z.restype = vp
        lib.nmpc_peers_info.argtypes = [vp]; lib.nmpc_peers_info.restype = vp
        for name in ("nmpc_peers_status", "nmpc_peers_destroy"):
            getattr(lib, name).argtypes = [vp]
        lib.nmpc_peers_barrier.argtypes = [vp, vp]
        sig = [vp, i, i, i] + [vp] * 5 + [i, ctypes.POINTER(_lib.NmpcOpts), vp]
        lib.nmpc_solve_batch_sharded_p2p_f64.argtypes = sig + [i, vp]
        lib.nmpc_solve_batch_sharded_p2p_f32.argtypes = sig + [vp]
        self.peers = vp()
        zbytes = world * per * N * 17 * self.np_dtype.itemsize
        with torch.cuda.device(self.device):
            err = None
            mine = ctypes.create_string_buffer(64)
            if lib.nmpc_peers_create(world, rank, zbytes, world * per * 4, ctypes.byref(self.peers)) != 0 or \
                    (world > 1 and lib.nmpc_peers_export(self.peers, mine) != 0):
                err = _lib.last_error()
            if world > 1:
                if handles is None:
                    # every rank takes part in the exchange, also one whose allocation failed (it sends None): nobody is
                    # left waiting in the collective
                    import torch.distributed as dist
                    box = [None] * world
                    dist.all_gather_object(box, None if err else mine.raw, group=group)
                    if any(h is None for h in box):
                        err = err or "a peer rank could not allocate or export its buffers"
                    else:
                        handles = b"".join(box)
                if err is None:
                    _check(lib.nmpc_peers_connect(self.peers, handles))
            if err is not None:
                self.close()
                raise RuntimeError(f"nmpc_peers setup failed: {err}")
        t_dt = torch.float64 if self.np_dtype == np.float64 else torch.float32
        self.z_all = _wrap_device_memory(torch, lib.nmpc_peers_z(self.peers), (world * per, N, 17), t_dt, self.device, self)
        self.info_all = _wrap_device_memory(torch, lib.nmpc_peers_info(self.peers), (world * per, 4), torch.int32, self.device, self)

    def solve_sharded(self, db, opts=None, mode: int = 0, stream=None):
        """Enqueue barrier + solve with peer stores + barrier for this rank's DeviceBatch (db.B == per on every rank).
        mode: 0 fp64 kernel, 1 mixed precision, 2 low-latency warp-group kernel."""
        torch = self.torch
        assert db.B == self.per and db.N == self.N and db.np_dtype == self.np_dtype
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        o = opts or _lib.default_opts()
        d = db.d
        args = [self.peers, db.B, db.N, db.mcap, d["xinit"].data_ptr(), d["z0"].data_ptr(), d["hdr"].data_ptr(),
                d["rows"].data_ptr(), d["nrows"].data_ptr(), db.variant, ctypes.byref(o), db.info_real.data_ptr()]
        with torch.cuda.device(self.device):
            if self.np_dtype == np.float32:
                _check(self.lib.nmpc_solve_batch_sharded_p2p_f32(*args, ctypes.c_void_p(st.cuda_stream)))
            else:
                _check(self.lib.nmpc_solve_batch_sharded_p2p_f64(*args, int(mode), ctypes.c_void_p(st.cuda_stream)))

    def barrier(self, stream=None):
        st = stream if stream is not None else self.torch.cuda.current_stream(self.device)
        with self.torch.cuda.device(self.device):
            _check(self.lib.nmpc_peers_barrier(self.peers, ctypes.c_void_p(st.cuda_stream)))

    def check(self):
        """Synchronous: raises if a barrier timed out (a rank missing for 2 s)."""
        with self.torch.cuda.device(self.device):
            _check(self.lib.nmpc_peers_status(self.peers))

    def close(self):
        """Call on every rank after a host-level barrier: no peer may still be writing into this rank's buffers."""
        if self.peers:
            with self.torch.cuda.device(self.device):
                self.lib.nmpc_peers_destroy(self.peers)
            self.peers = ctypes.c_void_p()


class TorchCollator:
    """The same in-place gather through torch.distributed (gloo on CPU in the tests, nccl on GPUs)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def collate(self, buf_all):
        per = buf_all.shape[0] // self.world
        self.dist.all_gather_into_tensor(buf_all, buf_all[self.rank * per:(self.rank + 1) * per], group=self.group)


def solve_sharded(batch: Batch, local_solver, device, collator, dtype=np.float64):
    """Solve `batch` across the ranks of `collator` and collate the results on every rank.

    `local_solver(shard, z_out, info_out)` writes this rank's results into the two views it is given -- slices of the
    collation buffers, so nothing is copied afterwards (the CUDA solver in production; the gloo tests inject a CPU
    stand-in).  Returns (z [B,N,17], flag [B], it [B]) as views of the gathered buffers, trimmed to the true B."""
    import torch
    rank, world = collator.rank, collator.world
    per = per_rank(batch.B, world)
    mine = shard(batch, rank, world, pad=True)
    t_dt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    z_all = torch.empty((world * per, batch.N, 17), dtype=t_dt, device=device)
    info_all = torch.empty((world * per, 4), dtype=torch.int32, device=device)
    local_solver(mine, z_all[rank * per:(rank + 1) * per], info_all[rank * per:(rank + 1) * per])
    collator.collate(z_all)
    collator.collate(info_all)
    return z_all[:batch.B], info_all[:batch.B, 0], info_all[:batch.B, 1]
