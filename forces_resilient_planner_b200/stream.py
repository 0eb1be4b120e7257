"""Warm-started receding-horizon stream (BASELINE config 5): B agents replanning in lock-step, the
whole replan cycle resident on the device.

One replan = the reference's solveNMPC sequence for every agent at once
(/root/reference/src/resilient_planner/plan_manage/src/nmpc_solver.cpp:346-482):

    shift warm start   x0[i] <- z[i+1], last stage duplicated, xinit <- z[1][8:17]
                       (forces_normal.cpp:62-97, nmpc_solver.cpp:531-543)   nmpc_shift_warm_start_f64
    ellipsoids         (optional) E_i propagated along the previous plan
                       (setFORCESParams, nmpc_solver.cpp:484-521)            nmpc_propagate_ellipsoids_f64
    pack parameters    refs, f_ext, yaw refs, tightened corridor rows
                       (forces_normal.cpp:100-136)                          nmpc_pack_params_f64
    solve              FORCESNLPsolver_normal_solve for every agent          nmpc_solve_batch_ordered_f64
    result handling    exit flag 1 -> the plan is adopted; otherwise the agent keeps nothing of the failed
                       solve and cold-starts next cycle (nmpc_solver.cpp:398-427, 363-364)   nmpc_adopt_plans_f64

With a perfect-model plant the next initial state is the predicted stage-1 state, exactly what the
reference feeds back (`xinit = mpc_output[1][8:17]`, not odometry).  The three launches of a replan
can be captured once into a CUDA graph and replayed.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib, prep, workloads as W
from .solver import _check


class RecedingHorizonStream:
    def __init__(self, batch: W.Batch, device="cuda:0", mu0_warm: float = 0.1, use_graph: bool = True,
                 wrap_yaw: bool = False, dynamic_ellipsoids: bool = False, longest_first: bool = True,
                 mixed: bool = False, lowlatency: bool | None = None):
        import torch
        self.torch = torch
        self.dev = torch.device(device)
        self.B, self.N, self.mcap = batch.B, batch.N, batch.mcap
        t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a if dt is None else a.astype(dt))).to(self.dev)
        # static per-agent corridor (one polytope per agent, all stages), raw (untightened) rows
        A = batch.rows[:, 1, :, 0:3]
        b_raw = batch.rows[:, 1, :, 3] + np.linalg.norm(A * W.EGO_E, axis=-1)
        self.poly_A = t(A[:, None]); self.poly_b = t(b_raw[:, None])
        self.poly_m = t(batch.nrows[:, 1:2], np.int32); self.poly_idx = t(np.zeros((self.B, self.N)), np.int32)
        self.ellipsoid = t(np.tile(np.diag(W.EGO_E).reshape(1, 1, 9), (self.B, self.N, 1)))
        self.weights = (7.0, 1.0, 80.0, 12.0, 0.5)
        # dynamic inputs of a replan (host-generated): ONE pinned staging buffer, ONE H2D copy per cycle, three views
        n0, n1, n2 = self.B * self.N * 3, self.B * self.N, self.B * 3
        self.h_in = torch.empty(n0 + n1 + n2, dtype=torch.float64).pin_memory()
        self.d_in = torch.empty(n0 + n1 + n2, dtype=torch.float64, device=self.dev)
        self.ref_pos = self.d_in[0:n0].view(self.B, self.N, 3); self.ref_yaw = self.d_in[n0:n0 + n1].view(self.B, self.N)
        self.ext_acc = self.d_in[n0 + n1:].view(self.B, 3)
        self._h_views = (self.h_in[0:n0].view(self.B, self.N, 3).numpy(), self.h_in[n0:n0 + n1].view(self.B, self.N).numpy(),
                         self.h_in[n0 + n1:].view(self.B, 3).numpy())
        # results a replan hands back to the host: first commands and (flag, iterations, ...), pinned
        self.h_cmd = torch.empty((self.B, 4), dtype=torch.float64).pin_memory()
        self.h_ii = torch.empty((self.B, 4), dtype=torch.int32).pin_memory()
        # solver state on the device
        self.xinit = t(batch.xinit); self.z0 = t(batch.z0); self.z = torch.empty_like(self.z0)
        self.zprev = self.z0.clone()      # the plan in force (mpc_output_): only ACCEPTED solves ever enter it
        self.hdr = torch.empty((self.B, self.N, 10), dtype=torch.float64, device=self.dev)
        self.rows = torch.empty((self.B, self.N, self.mcap, 4), dtype=torch.float64, device=self.dev)
        self.nrows = torch.empty((self.B, self.N), dtype=torch.int32, device=self.dev)
        self.info_int = torch.zeros((self.B, 4), dtype=torch.int32, device=self.dev)
        self.info_real = torch.zeros((self.B, 8), dtype=torch.float64, device=self.dev)
        self.cold_z0 = self.z0.clone(); self.cold_x = self.xinit.clone()
        self.opts_cold = _lib.default_opts()
        self.opts_warm = _lib.default_opts(mu0=mu0_warm)
        self.lib = _lib.load()
        self.graph = None
        self.use_graph = use_graph
        # The reference wraps each stage's yaw into (-pi, pi] after a solve (nmpc_solver.cpp:531-541) and
        # recomputes the yaw reference relative to the wrapped value (calculate_yaw, :834-862).  A caller
        # that keeps its yaw references unwrapped (as synthetic_refs does) must leave the states unwrapped
        # too, otherwise the warm start asks for a spurious 2*pi rotation.
        self.wrap_yaw = wrap_yaw
        # True: the corridor is tightened with the disturbance ellipsoids propagated along the previous plan
        # (setFORCESParams, nmpc_solver.cpp:484-521 -> nmpc_propagate_ellipsoids_f64) instead of the static
        # ego ellipsoid; the cold start propagates along the cold guess, as the reference does after
        # initMPCOutput (:363-364).
        self.dynamic_ellipsoids = dynamic_ellipsoids
        # True: warm solves are launched longest-first, ranked by the previous replan's iteration counts (failed
        # agents, which restart cold, first).  CTAs start in index order, so the agents that spill into the
        # second wave (1024 agents > 148 x 6 resident warps) are the quick ones and the launch ends sooner.
        self.longest_first = longest_first
        # True: the replans run the mixed-precision kernel (nmpc_solve_batch_mixed_f64: same tolerances, 25.8 KB instead of
        # 36.4 KB of shared memory per agent -> 1184 instead of 888 resident agents per GPU, so a 1024-agent fleet is ONE
        # wave, and every iteration is shorter); its fp64 re-solve of an agent it gives up on rides on the same stream
        self.mixed = mixed
        # the warp-group kernel (nmpc_solve_batch_lowlatency_f64: 256 threads per agent, one agent per SM, ~0.65x the time
        # per iteration) pays when the fleet leaves most of the GPU idle; default: on for mixed streams of at most two
        # agents per SM (N = 20: beyond one agent per SM the library switches to its 128-register build, two CTAs per
        # SM; a second WAVE would cost more than the shorter iterations save)
        if lowlatency is None:
            sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
            lowlatency = mixed and self.B <= (2 * sms if self.N == 20 else sms)
        self.lowlatency = bool(lowlatency)
        self.order = torch.arange(self.B, dtype=torch.int32, device=self.dev)
        self.cycle = 0

    # -- the three launches of a replan, all on `stream` ---------------------------------------------
    def _enqueue(self, stream, warm: bool):
        torch = self.torch
        with torch.cuda.stream(stream):
            self.d_in.copy_(self.h_in, non_blocking=True)          # this cycle's references and f_ext
        if warm:
            # result handling of the previous cycle, per agent: accepted plans are adopted, a failed solve (whose
            # output may be NaN) leaves nothing behind -- that agent restarts from the cold guess at its state
            prep.adopt_plans(self.z, self.info_int, self.zprev, wrap_yaw=False, stream=stream)
            prep.shift_warm_start(self.zprev, self.xinit, self.z0, wrap_yaw=self.wrap_yaw, stream=stream)
        if self.dynamic_ellipsoids:
            prep.propagate_ellipsoids(self.zprev if warm else self.z0, out=self.ellipsoid, stream=stream)
        hdr, rows, nrows = self.hdr, self.rows, self.nrows
        w = (ctypes.c_double * 5)(*self.weights)
        fn = self.lib.nmpc_pack_params_f64
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p] * 8 + [ctypes.POINTER(ctypes.c_double)] + [ctypes.c_void_p] * 4
        _check(fn(self.B, self.N, 1, self.mcap, self.mcap, self.ref_pos.data_ptr(), self.ref_yaw.data_ptr(),
                  self.ext_acc.data_ptr(), self.ellipsoid.data_ptr(), self.poly_A.data_ptr(), self.poly_b.data_ptr(),
                  self.poly_m.data_ptr(), self.poly_idx.data_ptr(), w, hdr.data_ptr(), rows.data_ptr(),
                  nrows.data_ptr(), stream.cuda_stream))
        o = self.opts_warm if warm else self.opts_cold
        order = None
        if warm and self.longest_first:
            prep.rank_longest_first(self.info_int, self.order, stream=stream)
            order = self.order.data_ptr()
        args = [self.B, self.N, self.mcap, self.xinit.data_ptr(), self.z0.data_ptr(), hdr.data_ptr(), rows.data_ptr(),
                nrows.data_ptr(), 0, ctypes.byref(o), self.z.data_ptr(), self.info_int.data_ptr(), self.info_real.data_ptr()]
        if self.lowlatency:
            _check(self.lib.nmpc_solve_batch_lowlatency_f64(*args, None, None, None, None, order, ctypes.c_void_p(stream.cuda_stream)))
        elif self.mixed:
            _check(self.lib.nmpc_solve_batch_mixed_f64(*args, None, None, None, None, order, ctypes.c_void_p(stream.cuda_stream)))
        else:
            fn = self.lib.nmpc_solve_batch_ordered_f64
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.POINTER(_lib.NmpcOpts)] + \
                [ctypes.c_void_p] * 5
            _check(fn(*args, order, ctypes.c_void_p(stream.cuda_stream)))
        with torch.cuda.stream(stream):                             # what the host reads back: commands, flags, iterations
            self.h_cmd.copy_(self.z[:, 0, 0:4], non_blocking=True)
            self.h_ii.copy_(self.info_int, non_blocking=True)

    def replan(self, ref_pos: np.ndarray, ref_yaw: np.ndarray, ext_acc: np.ndarray):
        """One cycle.  Host arrays in (refs of this cycle), host arrays out (first command, flags).

        Cycle 0 is the cold start; cycle 1 launches the warm sequence directly (which also sets the
        kernels' function attributes); before cycle 2 the warm sequence is captured into a CUDA graph
        (capture does not execute) and every later cycle replays it."""
        torch = self.torch
        st = torch.cuda.current_stream(self.dev)
        self._h_views[0][...] = ref_pos; self._h_views[1][...] = ref_yaw; self._h_views[2][...] = ext_acc
        if self.cycle == 0:
            self._enqueue(st, warm=False)
        elif self.cycle == 1 or not self.use_graph:
            self._enqueue(st, warm=True)
        else:
            if self.graph is None:
                torch.cuda.synchronize(self.dev)
                self.graph = torch.cuda.CUDAGraph()
                # thread_local: other threads of the process (NCCL watchdogs, a sampler) may keep calling CUDA while we capture
                with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                    self._enqueue(torch.cuda.current_stream(self.dev), warm=True)
            self.graph.replay()
        self.cycle += 1
        st.synchronize()                                  # the step's results are in the pinned buffers now
        ii = self.h_ii.numpy()
        return self.h_cmd.numpy().copy(), ii[:, 0].copy(), ii[:, 1].copy()

def synthetic_refs(batch: W.Batch, step: int, rng: np.random.Generator, ext_acc: np.ndarray, period: int = 20):
    """References of replan `step` for the config-2 style scenario: the straight-line reference
    shuttles back and forth along its heading (triangle wave, `period` replans each way) so that a
    long stream stays inside its corridor; the yaw reference holds its settled value; f_ext
    random-walks with sigma 0.1 per replan."""
    hdr = batch.hdr
    v = (hdr[:, 1, 0:3] - hdr[:, 0, 0:3])                     # reference displacement per stage
    ph = step % (2 * period)
    off = ph if ph <= period else 2 * period - ph             # 0..period..0
    kk = np.arange(batch.N)[None, :, None]
    # within the horizon the reference keeps moving in the current direction of the shuttle
    direction = 1.0 if ph < period else -1.0
    ref = hdr[:, 0:1, 0:3] + (off + direction * kk) * v[:, None, :]
    yaw = np.repeat(hdr[:, -1:, 9], batch.N, axis=1) if step > 0 else hdr[:, :, 9].copy()
    ext = np.clip(ext_acc + rng.normal(0.0, 0.1, ext_acc.shape), -2.5, 2.5) if step > 0 else ext_acc
    return np.ascontiguousarray(ref), np.ascontiguousarray(yaw), np.ascontiguousarray(ext)


class PlannerPipeline:
    """One replan of the reference's `setFORCESParams` + `solveNormal` for B agents, device-resident end to end
    (SURVEY.md §8f rank 1-4 + the solve):

        yaw wrap + shift        nmpc_wrap_yaw_f64, nmpc_shift_warm_start_f64   nmpc_solver.cpp:531-543, forces_normal.cpp:62-97
        disturbance ellipsoids  nmpc_propagate_ellipsoids_f64  :484-521, 567-699      (along the previous plan)
        references + yaw        nmpc_sample_reference_f64      :109-142, 834-862      (front-end polyline at Ts)
        corridors               nmpc_select_corridors_f64      :288-332, DecompROS    (obstacle cloud -> polytopes)
        parameters              nmpc_pack_params_f64           forces_normal.cpp:100-136
        solve                   nmpc_solve_batch_ordered_f64   FORCESNLPsolver_normal_solve

    The front end stays on the host: it supplies the polyline `kino_path` (sampled at Ts) and the obstacle `cloud`.
    """

    def __init__(self, xinit, kino_path, kino_size, cloud, cloud_n, device="cuda:0", mcap=30, max_polys=20,
                 weights=(7.0, 1.0, 80.0, 12.0, 0.5), Ts=0.05, N=20, mu0_warm=0.1):
        import torch
        self.torch = torch
        self.dev = torch.device(device)
        t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(self.dev)
        self.B, self.N, self.mcap, self.P, self.Ts = xinit.shape[0], N, mcap, max_polys, Ts
        self.kino_path, self.kino_size = t(kino_path), t(kino_size, torch.int32)
        self.cloud, self.cloud_n = t(cloud), t(cloud_n, torch.int32)
        self.weights = weights
        z0 = W.cold_start(xinit, N)                                   # initMPCOutput (nmpc_solver.cpp:265-286)
        self.xinit, self.z0 = t(xinit), t(z0)
        self.z = self.z0.clone()                                      # "previous plan" of the first cycle = the cold guess
        self.info_int = torch.zeros((self.B, 4), dtype=torch.int32, device=self.dev)
        self.info_real = torch.zeros((self.B, 8), dtype=torch.float64, device=self.dev)
        self.opts_cold, self.opts_warm = _lib.default_opts(), _lib.default_opts(mu0=mu0_warm)
        self.lib = _lib.load()
        self.cycle = 0
        self.last = {}
        from . import forces
        self.policy = [forces.SolveAcceptance() for _ in range(self.B)]   # solveNMPC's per-vehicle acceptance state

    def replan(self, ext_acc: np.ndarray, t_off: np.ndarray):
        """ext_acc [B,3], t_off [B] (= mpc_start_time_ - kino_start_time_) on the host -> (commands, flags, iterations)."""
        torch = self.torch
        st = torch.cuda.current_stream(self.dev)
        warm = self.cycle > 0
        ext = torch.from_numpy(np.ascontiguousarray(ext_acc)).to(self.dev)
        toff = torch.from_numpy(np.ascontiguousarray(t_off)).to(self.dev)
        prev = self.z                                                  # mpc_output_ of the last cycle (cold guess at first)
        if warm:
            prep.shift_warm_start(prev, self.xinit, self.z0, wrap_yaw=False, stream=st)   # prev is already wrapped (below)
        E = prep.propagate_ellipsoids(prev, stream=st)
        last_yaw = prev[:, 1, 16].contiguous(); pos1 = prev[:, 1, 8:11].contiguous()
        ref_pos, ref_yaw, far = prep.sample_reference(self.kino_path, self.kino_size, toff, last_yaw, self.N, self.Ts,
                                                      pos1=pos1, stream=st)
        pA, pb, pm, pidx, npoly, ovf = prep.select_corridors(self.cloud, self.cloud_n, ref_pos, ref_yaw, E,
                                                             max_polys=self.P, max_rows=self.mcap, stream=st)
        hdr, rows, nrows = prep.pack_params(ref_pos, ref_yaw, ext, E, pA, pb, pm, pidx, self.weights, self.mcap, stream=st)
        z_new = torch.empty_like(self.z)
        o = self.opts_warm if warm else self.opts_cold
        _check(self.lib.nmpc_solve_batch_f64(self.B, self.N, self.mcap, self.xinit.data_ptr(), self.z0.data_ptr(),
                                             hdr.data_ptr(), rows.data_ptr(), nrows.data_ptr(), 0, ctypes.byref(o),
                                             z_new.data_ptr(), self.info_int.data_ptr(), self.info_real.data_ptr(),
                                             ctypes.c_void_p(st.cuda_stream)))
        cmd = z_new[:, 0, 0:4].cpu().numpy()
        flag = self.info_int[:, 0].cpu().numpy()
        # solveNMPC's result handling (:398-427): the acceptance policy is per-vehicle host state; only accepted plans
        # are adopted (kept yaw-wrapped, updateFORCESResults :531-541), a rejected agent restarts cold next cycle
        accept = np.array([p.consume(int(f)) for p, f in zip(self.policy, flag)], dtype=np.int32)
        self.last_cold = prep.adopt_plans(z_new, self.info_int, self.z, accept=torch.from_numpy(accept).to(self.dev),
                                          wrap_yaw=True, cold=torch.empty((self.B,), dtype=torch.int32, device=self.dev),
                                          stream=st)
        self.cycle += 1
        self.last = dict(ellipsoid=E, ref_pos=ref_pos, ref_yaw=ref_yaw, hard_to_follow=far, poly_idx=pidx, n_poly=npoly,
                         overflow=ovf, hdr=hdr, rows=rows, nrows=nrows)
        return cmd, flag, self.info_int[:, 1].cpu().numpy()
