/*
 * nmpc_oracle.h -- CPU restatement of the reference NMPC hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this.  The product (forces_resilient_planner_b200/) never does.
 *
 * PARITY STATUS: the MODEL layer (objective / Heun dynamics / corridor rows and their Jacobians)
 * is PINNED against the reference's own CasADi-generated C (oracle/_ref, built from
 * /root/reference/.../solver/{normal,final}/FORCESNLPsolver_*_casadi.c) and the known-answer
 * vectors of SURVEY.md §8c (tests/golden/).  The SOLVER layer is "parity unpinned": the reference
 * interior-point core is a licence-locked ForcesPro v4.4.0 binary (exit -100 on any machine but
 * the authors'), ships no source and no golden solutions, so it can only be compared at the KKT
 * point: every solution is checked against ForcesPro's own acceptance thresholds
 * (TolStat/TolEq/TolIneq/TolComp = 1e-4, mpc_generator_normal.m:76-79) using the reference
 * callbacks, and cross-checked against scipy SLSQP driving the same callbacks.
 *
 * Build flavours: -DNMPC_REAL=double (default) or -DNMPC_REAL=float.
 */
#ifndef NMPC_ORACLE_H
#define NMPC_ORACLE_H

#ifndef NMPC_REAL
#define NMPC_REAL double
#endif
typedef NMPC_REAL real;

#ifdef __cplusplus
extern "C" {
#endif

/* Solver options: one struct shared (by value and meaning) with include/nmpc_b200.h:nmpc_opts. */
typedef struct {
    double mu0;        /* initial barrier parameter                                   */
    double sigma;      /* centering parameter: mu_target = max(sigma*mu, mu_floor)    */
    double mu_floor;   /* lowest barrier target (TolComp/10)                          */
    double tol_stat;   /* inf-norm stationarity      (codeoptions.nlp.TolStat  1e-4)  */
    double tol_eq;     /* inf-norm equality residual (codeoptions.nlp.TolEq    1e-4)  */
    double tol_ineq;   /* inf-norm inequality resid. (codeoptions.nlp.TolIneq  1e-4)  */
    double tol_comp;   /* max complementarity        (codeoptions.nlp.TolComp  1e-4)  */
    double kappa_push; /* relative push of the initial guess into the bound interior  */
    double s_floor;    /* floor on initial corridor slacks                            */
    int maxit;         /* iteration cap (codeoptions.maxit 200)                       */
    int max_bt;        /* backtracking steps per iteration                            */
    int pc;            /* 1: Mehrotra predictor-corrector (affine solve -> sigma, corrector rhs)  */
    int mixed;         /* 1: delta-form Newton system solved by a SINGLE-precision Riccati recursion, everything
                          else (iterate, residuals, step rule, line search) in `real`; 2: the same recursion in double
                          precision (the product's own factorisation on the CPU); 0: Schur complement + refinement  */
} nmpc_oracle_opts;

void nmpc_oracle_default_opts(nmpc_oracle_opts *o);
int nmpc_oracle_sizeof_real(void);

/* Mirror of FORCESNLPsolver_{normal,final}_casadi2forces for ONE stage: dense column-major
 * outputs (nabla_c 13x17, nabla_h 30x17), *f is overwritten (not accumulated). variant 0/1. */
void nmpc_oracle_model_eval(const double *z, const double *p130, int stage, int n_stages,
                            int variant, double *f, double *grad, double *c, double *jc,
                            double *h, double *jh);

/* Batched solve in the native layout (see forces_resilient_planner_b200/workloads.py).
 * info_int [B][4]  = exitflag, iterations, backtracks, reserved
 * info_real[B][8]  = res_eq, res_ineq, rsnorm, rcompnorm, pobj, mu, alpha_p, alpha_d      */
int nmpc_oracle_solve_batch(int B, int N, int mcap, const real *xinit, const real *z0,
                            const real *hdr, const real *rows, const int *nrows, int variant,
                            const nmpc_oracle_opts *opts, real *z_out, int *info_int,
                            real *info_real, int nthreads);

/* same, also returning the multipliers of the KKT point (any may be NULL):
 * y [B][N][13] (c-ordering, y[0]=0), zl / zu [B][N][17], lc [B][N][mcap]                  */
int nmpc_oracle_solve_batch_ex(int B, int N, int mcap, const real *xinit, const real *z0,
                               const real *hdr, const real *rows, const int *nrows, int variant,
                               const nmpc_oracle_opts *opts, real *z_out, int *info_int,
                               real *info_real, real *y_out, real *zl_out, real *zu_out,
                               real *lc_out, int nthreads);

/* single thread, problem after problem, every solve timed on its own (how the planner runs the reference solver:
 * num_of_threads = 1, forces_normal.cpp:31); seconds [B] */
int nmpc_oracle_solve_batch_timed(int B, int N, int mcap, const real *xinit, const real *z0,
                                  const real *hdr, const real *rows, const int *nrows, int variant,
                                  const nmpc_oracle_opts *opts, real *z_out, int *info_int,
                                  real *info_real, double *seconds);

/* One structured KKT solve (ForcesPro-style Schur complement, dense 17/13 blocks):
 *   min 1/2 dz' Phi dz + g' dz   s.t.  E dz_{k+1} = C_k dz_k + d_k ,  dz_0[8:17] = 0
 * Phi [N][17][17], g [N][17], C [N-1][13][17] (rows in c-ordering [x(9);u(4)]), d [N-1][13]
 * -> dz [N][17], y [N][13] (y[0] unused).  Returns 0, or -5 on a non-positive pivot.      */
int nmpc_oracle_kkt_solve(int N, const double *Phi, const double *g, const double *C,
                          const double *d, double *dz, double *y);

#ifdef __cplusplus
}
#endif
#endif
