"""Python mirror of the reference's solver wrappers, over the reference's own C ABI.

Mirrors `resilient_planner::FORCESNormal` / `FORCESFinal`
(/root/reference/src/resilient_planner/plan_manage/include/plan_manage/nmpc_utils.h:49-106,
 src/forces_normal.cpp:36-168, src/forces_final.cpp) with the same method names, argument meaning
and return convention, but calling `FORCESNLPsolver_{normal,final}_solve` exported by
libnmpc_b200.so instead of the ForcesPro archive.  The C++ twin lives in host/forces_wrappers.hpp.

The ctypes structures are the ABI contract: params 23600 B, output 2720 B,
info 136 B; machine-readable original: solver/normal/FORCESNLPsolver_normal/interface/definitions.py:10-60.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

NUM_PRE_PARAMS, NUM_CONST, NUM_ITER, NUM_VAR, HORIZON = 10, 30, 130, 17, 20   # nmpc_utils.h:52-58


class ForcesParams(ctypes.Structure):
    _fields_ = [("xinit", ctypes.c_double * 9), ("x0", ctypes.c_double * 340),
                ("all_parameters", ctypes.c_double * 2600), ("num_of_threads", ctypes.c_uint)]


class ForcesOutput(ctypes.Structure):
    _fields_ = [(f"x{k:02d}", ctypes.c_double * 17) for k in range(1, 21)]


class ForcesInfo(ctypes.Structure):
    _fields_ = [("it", ctypes.c_int), ("it2opt", ctypes.c_int), ("res_eq", ctypes.c_double),
                ("res_ineq", ctypes.c_double), ("rsnorm", ctypes.c_double),
                ("rcompnorm", ctypes.c_double), ("pobj", ctypes.c_double), ("dobj", ctypes.c_double),
                ("dgap", ctypes.c_double), ("rdgap", ctypes.c_double), ("mu", ctypes.c_double),
                ("mu_aff", ctypes.c_double), ("sigma", ctypes.c_double), ("lsit_aff", ctypes.c_int),
                ("lsit_cc", ctypes.c_int), ("step_aff", ctypes.c_double), ("step_cc", ctypes.c_double),
                ("solvetime", ctypes.c_double), ("fevalstime", ctypes.c_double)]


assert ctypes.sizeof(ForcesParams) == 23600 and ForcesParams.x0.offset == 72
assert ForcesParams.all_parameters.offset == 2792 and ForcesParams.num_of_threads.offset == 23592
assert ctypes.sizeof(ForcesOutput) == 2720 and ctypes.sizeof(ForcesInfo) == 136
assert ForcesInfo.lsit_aff.offset == 96 and ForcesInfo.solvetime.offset == 120


class _ForcesWrapper:
    _symbol = ""

    def __init__(self):
        self.params_ = ForcesParams()
        self.output_ = ForcesOutput()
        self.info_ = ForcesInfo()
        self.params_.num_of_threads = 1          # forces_normal.cpp:31
        lib = _lib.load()
        self._solve = getattr(lib, self._symbol)
        self._solve.restype = ctypes.c_int
        self._solve.argtypes = [ctypes.POINTER(ForcesParams), ctypes.POINTER(ForcesOutput),
                                ctypes.POINTER(ForcesInfo), ctypes.c_void_p, ctypes.c_void_p]

    def _set_paras(self, w_stage_wp, w_stage_input, w_input_rate, w_terminal_wp, w_terminal_input):
        p = self.params_.all_parameters
        for i in range(HORIZON):
            p[i * NUM_ITER + 6] = w_stage_wp
            p[i * NUM_ITER + 7] = w_stage_input
            p[i * NUM_ITER + 8] = w_input_rate
        p[(HORIZON - 1) * NUM_ITER + 6] = w_terminal_wp
        p[(HORIZON - 1) * NUM_ITER + 7] = w_terminal_input

    def _solve_impl(self, mpc_output, external_acc, ref_total_pos, ref_total_yaw, ellipsoid_matrices,
                    poly_constraints, poly_indices):
        """forces_normal.cpp:55-140.  mpc_output: sequence of >= 21 stage vectors (the planner's
        MPCDeque); poly_constraints: list of (A [m,3], b [m]); ellipsoid_matrices: list of 3x3."""
        P = self.params_
        for j in range(9):
            P.xinit[j] = mpc_output[1][8 + j]                    # predicted state, not odometry
        for i in range(HORIZON):
            for j in range(NUM_VAR):
                P.x0[i * NUM_VAR + j] = mpc_output[i + 1][j]     # shift warm start
            base = i * NUM_ITER
            for j in range(3):
                P.all_parameters[base + j] = ref_total_pos[i][j]
                P.all_parameters[base + 3 + j] = external_acc[j]
            P.all_parameters[base + 9] = ref_total_yaw[i]
            A, b = poly_constraints[int(poly_indices[i])]
            E = np.asarray(ellipsoid_matrices[i], float)
            for j in range(NUM_CONST):                            # rows beyond 30 are dropped
                if j < len(b):
                    for q in range(3):
                        P.all_parameters[base + NUM_PRE_PARAMS + 3 * j + q] = A[j][q]
                    P.all_parameters[base + NUM_PRE_PARAMS + 3 * NUM_CONST + j] = \
                        b[j] - float(np.linalg.norm(E @ np.asarray(A[j], float)))
                else:
                    for q in range(3):
                        P.all_parameters[base + NUM_PRE_PARAMS + 3 * j + q] = 0.0
                    P.all_parameters[base + NUM_PRE_PARAMS + 3 * NUM_CONST + j] = 0.0
        return self.solve_params()

    def solve_params(self) -> int:
        """Call the solver on whatever is in params_ (the reference passes fs = NULL)."""
        return int(self._solve(ctypes.byref(self.params_), ctypes.byref(self.output_),
                               ctypes.byref(self.info_), None, None))

    def _update(self, mpc_output):
        """forces_normal.cpp:142-168: x01..x20 -> mpc_output[0..19]."""
        for k in range(HORIZON):
            stage = getattr(self.output_, f"x{k + 1:02d}")
            for j in range(NUM_VAR):
                mpc_output[k][j] = stage[j]

    def output_array(self) -> np.ndarray:
        return np.ctypeslib.as_array(
            (ctypes.c_double * 340).from_buffer(self.output_)).reshape(20, 17).copy()


class FORCESNormal(_ForcesWrapper):
    _symbol = "FORCESNLPsolver_normal_solve"

    def setParasNormal(self, w_stage_wp, w_stage_input, w_input_rate, w_terminal_wp, w_terminal_input):
        self._set_paras(w_stage_wp, w_stage_input, w_input_rate, w_terminal_wp, w_terminal_input)

    def solveNormal(self, mpc_output, external_acc, ref_total_pos, ref_total_yaw, ellipsoid_matrices,
                    poly_constraints, poly_indices) -> int:
        return self._solve_impl(mpc_output, external_acc, ref_total_pos, ref_total_yaw,
                                ellipsoid_matrices, poly_constraints, poly_indices)

    def updateNormal(self, mpc_output):
        self._update(mpc_output)


class FORCESFinal(_ForcesWrapper):
    _symbol = "FORCESNLPsolver_final_solve"

    def setParasFinal(self, w_final_stage_wp, w_final_stage_input, w_input_rate, w_final_terminal_wp,
                      w_final_terminal_input):
        self._set_paras(w_final_stage_wp, w_final_stage_input, w_input_rate, w_final_terminal_wp,
                        w_final_terminal_input)

    def solveFinal(self, mpc_output, external_acc, ref_total_pos, ref_total_yaw, ellipsoid_matrices,
                   poly_constraints, poly_indices) -> int:
        return self._solve_impl(mpc_output, external_acc, ref_total_pos, ref_total_yaw,
                                ellipsoid_matrices, poly_constraints, poly_indices)

    def updateFinal(self, mpc_output):
        self._update(mpc_output)


class SolveAcceptance:
    """Exit-code acceptance policy of NMPCSolver::solveNMPC (nmpc_solver.cpp:398-421); twin of
    host/forces_wrappers.hpp::SolveAcceptance.  `consume(exit_code)` returns update_result."""

    def __init__(self):
        self.fail_count = 0
        self.replan_count = 0
        self.last_exit_code = 1
        self.kino_replan = False

    def consume(self, exit_code: int) -> bool:
        update_result = False
        self.last_exit_code = exit_code
        if exit_code == 1:
            self.fail_count = 0
            self.replan_count = 0
            update_result = True
        else:
            self.fail_count += 1
            if self.replan_count > 3 and exit_code == 0:
                self.fail_count = 0
                self.replan_count = 0
                update_result = True
            elif self.fail_count > 2:
                self.fail_count = 0
                self.replan_count += 1
                self.kino_replan = True
        return update_result

    def next_solve_is_cold(self, initialized_output: bool = True) -> bool:
        """nmpc_solver.cpp:363-364: `if (!initialized_output_ || exit_code != 1) initMPCOutput();`"""
        return (not initialized_output) or self.last_exit_code != 1
