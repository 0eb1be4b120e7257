/*
 * FORCESNLPsolver_final.h -- drop-in replacement header for the ForcesPro-generated solver ABI.
 *
 * Replaces /root/reference/src/resilient_planner/plan_manage/solver/final/FORCESNLPsolver_final/include/FORCESNLPsolver_final.h
 * (types :40-58, return codes :110-139, params :153-168, output :173-236, info :241-301,
 * solve prototype :321-323).  Written from the layout contract (sizes/offsets are checked by
 * static asserts in forces_resilient_planner_b200/csrc/nmpc_capi.cu and by tests/test_abi.py):
 *
 *   params  23600 B : xinit[9] @0, x0[340] @72, all_parameters[2600] @2792, num_of_threads @23592
 *   output   2720 B : x01 .. x20, 17 doubles each
 *   info      136 B : it @0, it2opt @4, res_eq @8 ... lsit_aff @96, lsit_cc @100, step_aff @104,
 *                     step_cc @112, solvetime @120, fevalstime @128
 *
 * The implementation behind FORCESNLPsolver_final_solve is the CUDA kernel in libnmpc_b200.so; the
 * callback argument is accepted and ignored (the device model is built in and tested equal to the
 * reference callback), num_of_threads is ignored, exit code -100 (licence) is never produced.
 */
#ifndef FORCESNLPsolver_final_H
#define FORCESNLPsolver_final_H

#ifndef SOLVER_STDIO_H
#define SOLVER_STDIO_H
#include <stdio.h>
#endif

#ifndef SOLVER_STANDARD_TYPES
#define SOLVER_STANDARD_TYPES
typedef signed char solver_int8_signed;
typedef unsigned char solver_int8_unsigned;
typedef char solver_int8_default;
typedef signed short int solver_int16_signed;
typedef unsigned short int solver_int16_unsigned;
typedef short int solver_int16_default;
typedef signed int solver_int32_signed;
typedef unsigned int solver_int32_unsigned;
typedef int solver_int32_default;
typedef signed long long int solver_int64_signed;
typedef unsigned long long int solver_int64_unsigned;
typedef long long int solver_int64_default;
#endif

typedef double FORCESNLPsolver_final_float;
typedef double FORCESNLPsolver_final_callback_float;
typedef double FORCESNLPsolver_finalinterface_float;

/* ---- solver settings the generated header publishes as macros (reference header :62-108); callers may test them.
 * Values are the reference's.  Note SET_ACC_* are ForcesPro's QP-era defaults (1e-6); the NLP stopping test the
 * solver was generated with -- and that this implementation uses -- is 1e-4 (mpc_generator_final.m:76-79). */
#ifndef MISRA_C_FORCESNLPsolver_final
#define MISRA_C_FORCESNLPsolver_final (0)
#endif
#ifndef RESTRICT_CODE_FORCESNLPsolver_final
#define RESTRICT_CODE_FORCESNLPsolver_final (0)
#endif
#ifndef SET_PRINTLEVEL_FORCESNLPsolver_final
#define SET_PRINTLEVEL_FORCESNLPsolver_final (1)      /* one summary line when the FILE* argument is non-NULL */
#endif
#ifndef SET_TIMING_FORCESNLPsolver_final
#define SET_TIMING_FORCESNLPsolver_final (1)          /* info.solvetime is filled (wall clock of the call)    */
#endif
#define SET_MAXIT_FORCESNLPsolver_final (200)
#define SET_FLS_SCALE_FORCESNLPsolver_final (FORCESNLPsolver_final_float)(0.99)
#define MAX_FILTER_SIZE_FORCESNLPsolver_final (200)
#define MAX_SOC_IT_FORCESNLPsolver_final (4)
#define SET_ACC_RDGAP_FORCESNLPsolver_final (FORCESNLPsolver_final_float)(0.0001)
#define SET_ACC_RESEQ_FORCESNLPsolver_final (FORCESNLPsolver_final_float)(1E-06)
#define SET_ACC_RESINEQ_FORCESNLPsolver_final (FORCESNLPsolver_final_float)(1E-06)
#define SET_ACC_KKTCOMPL_FORCESNLPsolver_final (FORCESNLPsolver_final_float)(1E-06)
/* integrator return codes (reference header :141-145) */
#ifndef INTEGRATOR_SUCCESS
#define INTEGRATOR_SUCCESS (11)
#define INTEGRATOR_MAXSTEPS_EXCEEDED (12)
#endif
#define OPTIMAL_FORCESNLPsolver_final (1)
#define MAXITREACHED_FORCESNLPsolver_final (0)
#define TIMEOUT_FORCESNLPsolver_final (2)
#define INVALID_NUM_INEQ_ERROR_FORCESNLPsolver_final (-4)
#define FACTORIZATION_ERROR_FORCESNLPsolver_final (-5)
#define BADFUNCEVAL_FORCESNLPsolver_final (-6)
#define NOPROGRESS_FORCESNLPsolver_final (-7)
#define PARAM_VALUE_ERROR_FORCESNLPsolver_final (-11)
#define INVALID_TIMEOUT_FORCESNLPsolver_final (-12)
#define LICENSE_ERROR_FORCESNLPsolver_final (-100)

typedef struct {
    FORCESNLPsolver_final_float xinit[9];              /* initial state: pos, vel, rpy            */
    FORCESNLPsolver_final_float x0[340];               /* initial guess, 20 stages x 17           */
    FORCESNLPsolver_final_float all_parameters[2600];  /* 20 stages x 130 run-time parameters     */
    solver_int32_unsigned num_of_threads;            /* ignored                                 */
} FORCESNLPsolver_final_params;

#define NMPC_B200_STAGE(n) FORCESNLPsolver_final_float x##n[17];
typedef struct {
    NMPC_B200_STAGE(01) NMPC_B200_STAGE(02) NMPC_B200_STAGE(03) NMPC_B200_STAGE(04) NMPC_B200_STAGE(05)
    NMPC_B200_STAGE(06) NMPC_B200_STAGE(07) NMPC_B200_STAGE(08) NMPC_B200_STAGE(09) NMPC_B200_STAGE(10)
    NMPC_B200_STAGE(11) NMPC_B200_STAGE(12) NMPC_B200_STAGE(13) NMPC_B200_STAGE(14) NMPC_B200_STAGE(15)
    NMPC_B200_STAGE(16) NMPC_B200_STAGE(17) NMPC_B200_STAGE(18) NMPC_B200_STAGE(19) NMPC_B200_STAGE(20)
} FORCESNLPsolver_final_output;
#undef NMPC_B200_STAGE

typedef struct {
    solver_int32_default it;                  /* iterations                                        */
    solver_int32_default it2opt;              /* = it                                              */
    FORCESNLPsolver_final_float res_eq;         /* inf-norm of the equality residuals                */
    FORCESNLPsolver_final_float res_ineq;       /* inf-norm of the inequality residuals              */
    FORCESNLPsolver_final_float rsnorm;         /* inf-norm of the stationarity residual             */
    FORCESNLPsolver_final_float rcompnorm;      /* largest complementarity product                   */
    FORCESNLPsolver_final_float pobj;           /* primal objective                                  */
    FORCESNLPsolver_final_float dobj;           /* pobj - dgap                                       */
    FORCESNLPsolver_final_float dgap;           /* mu * number of inequalities                       */
    FORCESNLPsolver_final_float rdgap;          /* |dgap / pobj|                                     */
    FORCESNLPsolver_final_float mu;             /* duality measure                                   */
    FORCESNLPsolver_final_float mu_aff;         /* = mu (no affine step in this algorithm)           */
    FORCESNLPsolver_final_float sigma;          /* centering parameter                               */
    solver_int32_default lsit_aff;            /* 0                                                 */
    solver_int32_default lsit_cc;             /* total backtracking steps                          */
    FORCESNLPsolver_final_float step_aff;       /* last dual step length                             */
    FORCESNLPsolver_final_float step_cc;        /* last primal step length                           */
    FORCESNLPsolver_final_float solvetime;      /* wall-clock seconds of the call                    */
    FORCESNLPsolver_final_float fevalstime;     /* 0 (model evaluation is fused into the kernel)     */
} FORCESNLPsolver_final_info;

#ifdef __cplusplus
extern "C" {
#endif
typedef void (*FORCESNLPsolver_final_extfunc)(FORCESNLPsolver_final_float *x, FORCESNLPsolver_final_float *y,
    FORCESNLPsolver_final_float *lambda, FORCESNLPsolver_final_float *params, FORCESNLPsolver_final_float *pobj,
    FORCESNLPsolver_final_float *g, FORCESNLPsolver_final_float *c, FORCESNLPsolver_final_float *Jeq,
    FORCESNLPsolver_final_float *h, FORCESNLPsolver_final_float *Jineq, FORCESNLPsolver_final_float *H,
    solver_int32_default stage, solver_int32_default iterations, solver_int32_default threadID);

extern solver_int32_default FORCESNLPsolver_final_solve(FORCESNLPsolver_final_params *params,
    FORCESNLPsolver_final_output *output, FORCESNLPsolver_final_info *info, FILE *fs,
    FORCESNLPsolver_final_extfunc evalextfunctions_FORCESNLPsolver_final);
#ifdef __cplusplus
}
#endif
#endif
