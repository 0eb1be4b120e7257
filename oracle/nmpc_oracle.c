/*
 * nmpc_oracle.c -- CPU restatement of the reference NMPC hot path (plain C, OpenMP over the batch).
 *
 * TEST INFRASTRUCTURE ONLY -- see nmpc_oracle.h for who may call this and for the parity status
 * (model layer pinned to the reference CasADi C; solver layer "parity unpinned").
 *
 * What follows the reference (paths relative to
 * /root/reference/src/resilient_planner/plan_manage/):
 *   model constants ............ matlab_code/setup.m:17-43, mpc/normal/mpc_generator_normal.m:13-14,29-46
 *   continuous dynamics ........ matlab_code/dynamics/nonlinear_dynamics.m:21-40
 *   13 equalities [x+; u] ...... matlab_code/dynamics/transit.m:4-8, generator:4-5 (model.E)
 *   Heun ("RK2") discretisation  solver/normal/FORCESNLPsolver_normal_casadi.c:238-240,307-311,383-394
 *   stage costs ................ matlab_code/mpc/mpc_objective1.m:38-48, mpc/normal/mpc_objective_normal.m,
 *                                mpc/final/mpc_objectiveN_final.m:26
 *   corridor rows .............. matlab_code/mpc/mpc_corridorconst.m:7-10
 *   tolerances / maxit ......... mpc/normal/mpc_generator_normal.m:56,76-79
 *   callback output layout ..... solver/normal/FORCESNLPsolver_normal_casadi2forces.c:42-245
 *
 * Linear algebra: the KKT system is solved the way the ForcesPro binary's symbol table says it
 * does: Cholesky of every 17x17 stage block Phi_k, 13x17 triangular
 * solves V = C L^-T / W = D L^-T, 13x13 Schur blocks Y = VV' + WW', block-tridiagonal Cholesky,
 * forward/backward substitution.  The CUDA product uses a Riccati recursion instead, so the two
 * factorizations check each other.
 *
 * Outer algorithm (identical, step for step, to the CUDA kernel -- DESIGN.md "Algorithm"):
 * primal-dual interior point, Gauss-Newton Hessian (the callbacks never provide `hess`), fixed
 * centering sigma with barrier floor, fraction-to-boundary step rule, backtracking line search
 * with a (theta, barrier-objective) sufficient-decrease test.
 */
#include "nmpc_oracle.h"

#include <math.h>
#include <tgmath.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NZ 17
#define NXI 13
#define NS_MAX 64
#define MC_MAX 32

#define DT ((real)0.05)
#define MASS ((real)0.745319)
#define GRAV ((real)9.81)
#define KD ((real)0.33)
#define HU ((real)1e-5)
#define PI_D 3.14159265358979323846
#define RATE_MAX ((real)(PI_D / 2))

static const double LB_D[NZ] = {-PI_D / 2, -PI_D / 2, -PI_D / 2, 0.5 * 9.81 * 0.745319,
                                -PI_D / 2, -PI_D / 2, -PI_D / 2, 0.5 * 9.81 * 0.745319,
                                -20, -20, 0, -2, -2, -2, -0.4 * PI_D, -0.4 * PI_D, -2 * PI_D};
static const double UB_D[NZ] = {PI_D / 2, PI_D / 2, PI_D / 2, 2.0 * 9.81 * 0.745319,
                                PI_D / 2, PI_D / 2, PI_D / 2, 2.0 * 9.81 * 0.745319,
                                20, 20, 5, 2, 2, 2, 0.4 * PI_D, 0.4 * PI_D, 2 * PI_D};

int nmpc_oracle_sizeof_real(void) { return (int)sizeof(real); }

void nmpc_oracle_default_opts(nmpc_oracle_opts *o)
{
    o->mu0 = 1.0;
    o->sigma = 0.1;
    o->mu_floor = 1e-5;
    o->tol_stat = o->tol_eq = o->tol_ineq = o->tol_comp = 1e-4;
    o->kappa_push = 1e-2;
    o->s_floor = 1e-2;
    o->maxit = 200;
    o->max_bt = 6;
    o->pc = 0;
    o->mixed = 0;
}

/* ------------------------------------------------------------------ model ---------------- */

static void zb_jac(const real r[3], real zb[3], real Z[3][3])
{
    real sr = sin(r[0]), cr = cos(r[0]), sp = sin(r[1]), cp = cos(r[1]), sy = sin(r[2]), cy = cos(r[2]);
    zb[0] = cy * sp * cr + sy * sr;
    zb[1] = sy * sp * cr - cy * sr;
    zb[2] = cp * cr;
    if (Z) {
        Z[0][0] = -cy * sp * sr + sy * cr; Z[0][1] = cy * cp * cr; Z[0][2] = -sy * sp * cr + cy * sr;
        Z[1][0] = -sy * sp * sr - cy * cr; Z[1][1] = sy * cp * cr; Z[1][2] = cy * sp * cr + sy * sr;
        Z[2][0] = -cp * sr;                Z[2][1] = -sp * cr;     Z[2][2] = 0;
    }
}

/* acc = z_B T/m + f_ext - g e3 - kd (v - z_B (z_B.v))   [R diag(kd,kd,0) R' = kd (I - z_B z_B')] */
static void acc_jac(const real v[3], const real r[3], real T, const real fe[3], real a[3],
                    real Av[3][3], real Ar[3][3], real AT[3], int jac)
{
    real zb[3], Z[3][3];
    zb_jac(r, zb, jac ? Z : NULL);
    real zv = zb[0] * v[0] + zb[1] * v[1] + zb[2] * v[2];
    for (int i = 0; i < 3; i++) a[i] = zb[i] * (T / MASS) + fe[i] - KD * (v[i] - zb[i] * zv);
    a[2] -= GRAV;
    if (!jac) return;
    real vZ[3];
    for (int j = 0; j < 3; j++) vZ[j] = v[0] * Z[0][j] + v[1] * Z[1][j] + v[2] * Z[2][j];
    real w = T / MASS + KD * zv;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) {
            Av[i][j] = -KD * ((i == j ? (real)1 : (real)0) - zb[i] * zb[j]);
            Ar[i][j] = w * Z[i][j] + KD * zb[i] * vZ[j];
        }
        AT[i] = zb[i] / MASS;
    }
}

/* c(z) = [Heun_h(x,u;f_ext) (9); u (4)], J = dc/dz (13x17, row-major here) */
static void dynamics(const real z[NZ], const real fe[3], real c[NXI], real (*J)[NZ])
{
    const real *w = z, *p = z + 8, *v = z + 11, *r = z + 14;
    real T = z[3], h = DT;
    real a1[3], A1v[3][3], A1r[3][3], A1T[3], a2[3], A2v[3][3], A2r[3][3], A2T[3];
    int jac = (J != NULL);
    acc_jac(v, r, T, fe, a1, A1v, A1r, A1T, jac);
    real v2[3], r2[3];
    for (int i = 0; i < 3; i++) { v2[i] = v[i] + h * a1[i]; r2[i] = r[i] + h * w[i]; }
    acc_jac(v2, r2, T, fe, a2, A2v, A2r, A2T, jac);
    for (int i = 0; i < 3; i++) {
        c[i] = p[i] + h * v[i] + (real)0.5 * h * h * a1[i];
        c[3 + i] = v[i] + (real)0.5 * h * (a1[i] + a2[i]);
        c[6 + i] = r[i] + h * w[i];
    }
    for (int i = 0; i < 4; i++) c[9 + i] = z[i];
    if (!jac) return;
    memset(J, 0, sizeof(real) * NXI * NZ);
    real hh = (real)0.5 * h * h, h2 = (real)0.5 * h;
    for (int i = 0; i < 3; i++) {
        J[i][8 + i] = 1;
        J[6 + i][14 + i] = 1;
        J[6 + i][i] = h;
        for (int j = 0; j < 3; j++) {
            real dv = 0, dr = 0;
            for (int k = 0; k < 3; k++) {
                dv += A2v[i][k] * ((k == j ? (real)1 : (real)0) + h * A1v[k][j]);
                dr += A2v[i][k] * h * A1r[k][j];
            }
            dr += A2r[i][j];
            J[i][11 + j] = (i == j ? h : (real)0) + hh * A1v[i][j];
            J[i][14 + j] = hh * A1r[i][j];
            J[3 + i][11 + j] = (i == j ? (real)1 : (real)0) + h2 * (A1v[i][j] + dv);
            J[3 + i][14 + j] = h2 * (A1r[i][j] + dr);
            J[3 + i][j] = hh * A2r[i][j];
        }
        real dT = A2T[i];
        for (int k = 0; k < 3; k++) dT += A2v[i][k] * h * A1T[k];
        J[i][3] = hh * A1T[i];
        J[3 + i][3] = h2 * (A1T[i] + dT);
    }
    for (int i = 0; i < 4; i++) J[9 + i][i] = 1;
}

/* stage cost and gradient; hdr = [ref(3) f_ext(3) w_wp w_in w_rate yaw_ref] */
static real objective(const real z[NZ], const real *hdr, int stage, int N, int variant, real *g)
{
    real wwp = hdr[6], win = hdr[7], wrate = hdr[8], yawref = hdr[9];
    real q = 1 / (RATE_MAX * RATE_MAX);
    real f = 0;
    if (g) memset(g, 0, sizeof(real) * NZ);
    for (int i = 0; i < 3; i++) {
        real e = hdr[i] - z[8 + i];
        f += wwp * e * e + win * q * z[i] * z[i];
        if (g) { g[8 + i] = -2 * wwp * e; g[i] = 2 * win * q * z[i]; }
    }
    real ey = yawref - z[16];
    f += 12 * wwp * ey * ey;
    if (g) g[16] = -24 * wwp * ey;
    for (int i = 0; i < 4; i++) {
        real du = z[i] - z[4 + i];
        f += wrate * du * du;
        if (g) { g[i] += 2 * wrate * du; g[4 + i] = -2 * wrate * du; }
    }
    if (stage == 0)
        for (int i = 0; i < 3; i++) {
            f += 10 * win * z[4 + i] * z[4 + i];
            if (g) g[4 + i] += 20 * win * z[4 + i];
        }
    if (stage == N - 1 && variant == 1)
        for (int i = 0; i < 3; i++) {
            f += 20 * wwp * z[11 + i] * z[11 + i];
            if (g) g[11 + i] += 40 * wwp * z[11 + i];
        }
    return f;
}

/* constant Gauss-Newton (= exact objective) Hessian */
static void cost_hessian(const real *hdr, int stage, int N, int variant, real H[NZ][NZ])
{
    real wwp = hdr[6], win = hdr[7], wrate = hdr[8];
    real q = 1 / (RATE_MAX * RATE_MAX);
    memset(H, 0, sizeof(real) * NZ * NZ);
    for (int i = 0; i < 4; i++) {
        H[i][i] = 2 * wrate + (i < 3 ? 2 * win * q : (real)0);
        H[4 + i][4 + i] = 2 * wrate + ((stage == 0 && i < 3) ? 20 * win : (real)0);
        H[i][4 + i] = H[4 + i][i] = -2 * wrate;
    }
    for (int i = 0; i < 3; i++) H[8 + i][8 + i] = 2 * wwp;
    H[16][16] = 24 * wwp;
    if (stage == N - 1 && variant == 1)
        for (int i = 0; i < 3; i++) H[11 + i][11 + i] = 40 * wwp;
}

void nmpc_oracle_model_eval(const double *z, const double *p, int stage, int n_stages, int variant,
                            double *f, double *grad, double *c, double *jc, double *h, double *jh)
{
    real zr[NZ], hdr[10], g[NZ], cc[NXI], J[NXI][NZ];
    for (int i = 0; i < NZ; i++) zr[i] = (real)z[i];
    for (int i = 0; i < 10; i++) hdr[i] = (real)p[i];
    *f = (double)objective(zr, hdr, stage, n_stages, variant, g);
    for (int i = 0; i < NZ; i++) grad[i] = g[i];
    if (stage < n_stages - 1) {
        dynamics(zr, hdr + 3, cc, J);
        for (int i = 0; i < NXI; i++) {
            c[i] = cc[i];
            for (int j = 0; j < NZ; j++) jc[j * NXI + i] = J[i][j];
        }
    }
    for (int j = 0; j < 30; j++) {
        real acc = -(real)p[100 + j];
        for (int i = 0; i < 3; i++) acc += (real)p[10 + 3 * j + i] * zr[8 + i];
        h[j] = acc;
        for (int i = 0; i < NZ; i++) jh[i * 30 + j] = (i >= 8 && i < 11) ? p[10 + 3 * j + i - 8] : 0.0;
    }
}

/* ------------------------------------------------------- small dense linear algebra ------ */

/* in-place lower Cholesky of n x n (row-major, leading dim ld); returns 0 or -5 */
static int chol(int n, real *A, int ld)
{
    for (int j = 0; j < n; j++) {
        real d = A[j * ld + j];
        for (int k = 0; k < j; k++) d -= A[j * ld + k] * A[j * ld + k];
        if (!(d > 0)) return -5;
        d = sqrt(d);
        A[j * ld + j] = d;
        for (int i = j + 1; i < n; i++) {
            real s = A[i * ld + j];
            for (int k = 0; k < j; k++) s -= A[i * ld + k] * A[j * ld + k];
            A[i * ld + j] = s / d;
        }
    }
    return 0;
}
/* x <- L^-1 x */
static void fsub(int n, const real *L, int ld, real *x)
{
    for (int i = 0; i < n; i++) {
        real s = x[i];
        for (int k = 0; k < i; k++) s -= L[i * ld + k] * x[k];
        x[i] = s / L[i * ld + i];
    }
}
/* x <- L^-T x */
static void bsub(int n, const real *L, int ld, real *x)
{
    for (int i = n - 1; i >= 0; i--) {
        real s = x[i];
        for (int k = i + 1; k < n; k++) s -= L[k * ld + i] * x[k];
        x[i] = s / L[i * ld + i];
    }
}

/* E selector in c-ordering: row i of (E dz) */
static inline int e_col(int i) { return i < 9 ? 8 + i : i - 9 + 4; }

/*
 * Structured KKT solve.  Phi [N][17][17] SPD, g [N][17], C [N-1][13][17], d [N-1][13].
 *   Phi dz + A' y = -g,   A dz = -d,   row block k (1..N-1):  C_{k-1} dz_{k-1} - E dz_k = -d_{k-1}
 * Stage-0 fixed variables must already be decoupled by the caller (identity rows, zero g, zero C cols).
 */
/* factor storage (per thread) */
static _Thread_local real kL[NS_MAX][NZ][NZ];
static _Thread_local real kV[NS_MAX][NXI][NZ];   /* V_k = C_k L_k^-T            (k = 0..N-2) */
static _Thread_local real kW[NS_MAX][NXI][NZ];   /* W_k = D   L_k^-T, D = -E    (k = 1..N-1) */
static _Thread_local real kYd[NS_MAX][NXI][NXI]; /* Cholesky factors of the Schur diagonal     */
static _Thread_local real kYo[NS_MAX][NXI][NXI]; /* L_{k+1,k}                                  */

static int kkt_factor(int N, real (*Phi)[NZ][NZ], real (*C)[NXI][NZ])
{
    for (int k = 0; k < N; k++) {
        memcpy(kL[k], Phi[k], sizeof(real) * NZ * NZ);
        if (chol(NZ, &kL[k][0][0], NZ)) return -5;
        if (k < N - 1)
            for (int r = 0; r < NXI; r++) {
                memcpy(kV[k][r], C[k][r], sizeof(real) * NZ);
                fsub(NZ, &kL[k][0][0], NZ, kV[k][r]);
            }
        if (k > 0)
            for (int r = 0; r < NXI; r++) {
                memset(kW[k][r], 0, sizeof(real) * NZ);
                kW[k][r][e_col(r)] = -1;
                fsub(NZ, &kL[k][0][0], NZ, kW[k][r]);
            }
    }
    /* Schur complement Y = A Phi^-1 A' (block tridiagonal) */
    for (int k = 1; k < N; k++) {
        for (int i = 0; i < NXI; i++)
            for (int j = 0; j <= i; j++) {
                real s = 0;
                for (int q = 0; q < NZ; q++) s += kV[k - 1][i][q] * kV[k - 1][j][q] + kW[k][i][q] * kW[k][j][q];
                kYd[k][i][j] = kYd[k][j][i] = s;
            }
        if (k < N - 1) /* Y_{k+1,k} = C_k Phi_k^-1 D' = V_k W_k' */
            for (int i = 0; i < NXI; i++)
                for (int j = 0; j < NXI; j++) {
                    real s = 0;
                    for (int q = 0; q < NZ; q++) s += kV[k][i][q] * kW[k][j][q];
                    kYo[k][i][j] = s;
                }
    }
    /* block-tridiagonal Cholesky */
    for (int k = 1; k < N; k++) {
        if (k > 1) {
            for (int i = 0; i < NXI; i++) fsub(NXI, &kYd[k - 1][0][0], NXI, kYo[k - 1][i]);
            for (int i = 0; i < NXI; i++)
                for (int j = 0; j <= i; j++) {
                    real s = 0;
                    for (int q = 0; q < NXI; q++) s += kYo[k - 1][i][q] * kYo[k - 1][j][q];
                    kYd[k][i][j] -= s;
                    if (j != i) kYd[k][j][i] -= s;
                }
        }
        if (chol(NXI, &kYd[k][0][0], NXI)) return -5;
    }
    return 0;
}

/* solve  Phi dz + A' y = -g,  A dz = -d  with the stored factor */
static void kkt_solve_rhs(int N, real (*g)[NZ], real (*d)[NXI], real (*dz)[NZ], real (*y)[NXI])
{
    static _Thread_local real t[NS_MAX][NZ];
    static _Thread_local real beta[NS_MAX][NXI];
    for (int k = 0; k < N; k++) {
        for (int i = 0; i < NZ; i++) t[k][i] = g[k][i];
        fsub(NZ, &kL[k][0][0], NZ, t[k]);
    }
    for (int k = 1; k < N; k++) {
        for (int i = 0; i < NXI; i++) {
            real b = d[k - 1][i];
            for (int q = 0; q < NZ; q++) b -= kV[k - 1][i][q] * t[k - 1][q] + kW[k][i][q] * t[k][q];
            if (k > 1) for (int q = 0; q < NXI; q++) b -= kYo[k - 1][i][q] * beta[k - 1][q];
            beta[k][i] = b;
        }
        fsub(NXI, &kYd[k][0][0], NXI, beta[k]);
    }
    for (int k = N - 1; k >= 1; k--) {
        if (k < N - 1)
            for (int i = 0; i < NXI; i++) {
                real s = 0;
                for (int q = 0; q < NXI; q++) s += kYo[k][q][i] * y[k + 1][q];
                beta[k][i] -= s;
            }
        bsub(NXI, &kYd[k][0][0], NXI, beta[k]);
        for (int i = 0; i < NXI; i++) y[k][i] = beta[k][i];
    }
    for (int i = 0; i < NXI; i++) y[0][i] = 0;
    for (int k = 0; k < N; k++) {
        real r[NZ];
        for (int q = 0; q < NZ; q++) {
            real s = t[k][q];
            if (k < N - 1) for (int i = 0; i < NXI; i++) s += kV[k][i][q] * y[k + 1][i];
            if (k > 0) for (int i = 0; i < NXI; i++) s += kW[k][i][q] * y[k][i];
            r[q] = -s;
        }
        bsub(NZ, &kL[k][0][0], NZ, r);
        for (int q = 0; q < NZ; q++) dz[k][q] = r[q];
    }
}

/*
 * Structured KKT solve.  Phi [N][17][17] SPD, g [N][17], C [N-1][13][17], d [N-1][13].
 *   Phi dz + A' y = -g,   A dz = -d,   row block k (1..N-1):  C_{k-1} dz_{k-1} - E dz_k = -d_{k-1}
 * Stage-0 fixed variables must already be decoupled by the caller (identity rows, zero g, zero C cols).
 * The normal-equations route squares the conditioning of Phi (barrier terms reach 1e8), so the
 * solution is polished by iterative refinement on the original saddle-point system.
 */
static int kkt_solve(int N, real (*Phi)[NZ][NZ], real (*g)[NZ], real (*C)[NXI][NZ], real (*d)[NXI],
                     real (*dz)[NZ], real (*y)[NXI])
{
    static _Thread_local real rg[NS_MAX][NZ], rd[NS_MAX][NXI], cz[NS_MAX][NZ], cy[NS_MAX][NXI];
    int rc = kkt_factor(N, Phi, C);
    if (rc) return rc;
    kkt_solve_rhs(N, g, d, dz, y);
    for (int pass = 0; pass < 2; pass++) {
        /* residuals: rg = g + Phi dz + A' y ,  rd = d + A dz  (both should be 0) */
        for (int k = 0; k < N; k++) {
            for (int i = 0; i < NZ; i++) {
                real s = g[k][i];
                for (int j = 0; j < NZ; j++) s += Phi[k][i][j] * dz[k][j];
                if (k < N - 1) for (int r = 0; r < NXI; r++) s += C[k][r][i] * y[k + 1][r];
                rg[k][i] = s;
            }
            if (k > 0) for (int r = 0; r < NXI; r++) rg[k][e_col(r)] -= y[k][r];
            if (k < N - 1)
                for (int r = 0; r < NXI; r++) {
                    real s = d[k][r] - dz[k + 1][e_col(r)];
                    for (int j = 0; j < NZ; j++) s += C[k][r][j] * dz[k][j];
                    rd[k][r] = s;
                }
        }
        kkt_solve_rhs(N, rg, rd, cz, cy);
        for (int k = 0; k < N; k++) {
            for (int i = 0; i < NZ; i++) dz[k][i] += cz[k][i];
            for (int i = 0; i < NXI; i++) y[k][i] += cy[k][i];
        }
    }
    return 0;
}

int nmpc_oracle_kkt_solve(int N, const double *Phi, const double *g, const double *C, const double *d,
                          double *dz, double *y)
{
    if (N > NS_MAX) return -11;
    real(*P)[NZ][NZ] = malloc(sizeof(real) * N * NZ * NZ);
    real(*G)[NZ] = malloc(sizeof(real) * N * NZ);
    real(*CC)[NXI][NZ] = malloc(sizeof(real) * N * NXI * NZ);
    real(*D)[NXI] = malloc(sizeof(real) * N * NXI);
    real(*DZ)[NZ] = malloc(sizeof(real) * N * NZ);
    real(*Y)[NXI] = malloc(sizeof(real) * N * NXI);
    for (int i = 0; i < N * NZ * NZ; i++) (&P[0][0][0])[i] = (real)Phi[i];
    for (int i = 0; i < N * NZ; i++) (&G[0][0])[i] = (real)g[i];
    for (int i = 0; i < (N - 1) * NXI * NZ; i++) (&CC[0][0][0])[i] = (real)C[i];
    for (int i = 0; i < (N - 1) * NXI; i++) (&D[0][0])[i] = (real)d[i];
    /* decouple the fixed stage-0 states */
    for (int i = 8; i < NZ; i++) {
        for (int j = 0; j < NZ; j++) P[0][i][j] = P[0][j][i] = 0;
        P[0][i][i] = 1;
        G[0][i] = 0;
        for (int r = 0; r < NXI; r++) CC[0][r][i] = 0;
    }
    int rc = kkt_solve(N, P, G, CC, D, DZ, Y);
    for (int i = 0; i < N * NZ; i++) dz[i] = (&DZ[0][0])[i];
    for (int i = 0; i < N * NXI; i++) y[i] = (&Y[0][0])[i];
    free(P); free(G); free(CC); free(D); free(DZ); free(Y);
    return rc;
}

/* ------------------------------------------------ mixed precision: single-precision Riccati ---- */
/*
 * opts.mixed = 1 (the CUDA product's *_mixed / *_f32 entry points; 2 = the same in double precision, i.e. the
 * fp64 product's own Riccati algorithm on the CPU): the Newton system is written in
 * DELTA form -- the right-hand side is the full KKT residual, evaluated in double precision at the
 * current (z, y), so the unknowns are (dz, dy) -- and solved by a Riccati recursion over
 * xi = [x(9); u_prev(4)] carried out entirely in SINGLE precision.  The iterate, the residuals, the
 * step rule and the line search stay in double precision, so the outer Newton iteration plays the role
 * of iterative refinement: the stopping test is still the reference's 1e-4 (mpc_generator_normal.m:76-79).
 * Dense restatement of what csrc/nmpc_ipm_mixed.cuh does with the structured Jacobian.
 *   min 1/2 dz'Phi dz + g'dz  s.t.  E dz_{k+1} = C_k dz_k + d_k,  dz_0[8:17] = 0;   dy = QP multipliers
 */
#define sreal float
#define RIC_NAME(x) x##_s
#define RIC_SQRT sqrtf
#include "riccati_dense.inc"
#undef sreal
#undef RIC_NAME
#undef RIC_SQRT
#define sreal double
#define RIC_NAME(x) x##_d
#define RIC_SQRT sqrt
#include "riccati_dense.inc"
#undef sreal
#undef RIC_NAME
#undef RIC_SQRT

/* ------------------------------------------------------------------ IPM ------------------ */

typedef struct {
    real f, theta, logsum;
    real g[NS_MAX][NZ];
    real J[NS_MAX][NXI][NZ];
    real d[NS_MAX][NXI];       /* defect d_k = c(z_k) - E z_{k+1}, k = 0..N-2 */
    real rc[NS_MAX][MC_MAX];   /* corridor residual a.pos - (b+hu) + s        */
} eval_t;

typedef struct {
    int N, mcap, variant;
    const real *hdr, *rows;
    const int *nrows;
    real lb[NZ], ub[NZ];
    real z[NS_MAX][NZ], y[NS_MAX][NXI], zl[NS_MAX][NZ], zu[NS_MAX][NZ];
    real s[NS_MAX][MC_MAX], lc[NS_MAX][MC_MAX];
    real zt[NS_MAX][NZ], st[NS_MAX][MC_MAX];
    real dz[NS_MAX][NZ], yn[NS_MAX][NXI], dzl[NS_MAX][NZ], dzu[NS_MAX][NZ];
    real ds[NS_MAX][MC_MAX], dlc[NS_MAX][MC_MAX];
    real ccl[NS_MAX][NZ], ccu[NS_MAX][NZ], ccr[NS_MAX][MC_MAX];   /* predictor-corrector second-order terms */
    real Phi[NS_MAX][NZ][NZ], gt[NS_MAX][NZ], Cm[NS_MAX][NXI][NZ];
    eval_t ev[2];
} work_t;

static inline int is_free(int k, int i) { return k > 0 || i < 8; }
static inline int live_rows(const work_t *w, int k) { return k == 0 ? 0 : (w->nrows[k] < w->mcap ? w->nrows[k] : w->mcap); }
static inline const real *row(const work_t *w, int k, int j) { return w->rows + ((size_t)k * w->mcap + j) * 4; }

static void evaluate(const work_t *w, real (*z)[NZ], real (*s)[MC_MAX], eval_t *e)
{
    int N = w->N;
    real f = 0, th = 0, ls = 0;
    for (int k = 0; k < N; k++) {
        const real *hdr = w->hdr + k * 10;
        f += objective(z[k], hdr, k, N, w->variant, e->g[k]);
        if (k < N - 1) {
            real c[NXI];
            dynamics(z[k], hdr + 3, c, e->J[k]);
            for (int i = 0; i < NXI; i++) {
                e->d[k][i] = c[i] - z[k + 1][e_col(i)];
                th += fabs(e->d[k][i]);
            }
        }
        for (int i = 0; i < NZ; i++)
            if (is_free(k, i)) ls += log(z[k][i] - w->lb[i]) + log(w->ub[i] - z[k][i]);
        int m = live_rows(w, k);
        for (int j = 0; j < m; j++) {
            const real *a = row(w, k, j);
            real r = a[0] * z[k][8] + a[1] * z[k][9] + a[2] * z[k][10] - (a[3] + HU) + s[k][j];
            e->rc[k][j] = r;
            th += fabs(r);
            ls += log(s[k][j]);
        }
    }
    e->f = f; e->theta = th; e->logsum = ls;
}

static int solve_one(work_t *w, const nmpc_oracle_opts *o, const real *xinit, const real *z0,
                     real *z_out, int *iinfo, real *rinfo, real *y_out, real *zl_out, real *zu_out, real *lc_out)
{
    const int N = w->N;
    const real eps = (sizeof(real) == 8) ? (real)2.220446049250313e-16 : (real)1.1920929e-07;
    int ncomp = 0;
    for (int i = 0; i < NZ; i++) { w->lb[i] = (real)LB_D[i]; w->ub[i] = (real)UB_D[i]; }
    /* ---- initial point */
    for (int k = 0; k < N; k++) {
        for (int i = 0; i < NZ; i++) {
            real v = z0[k * NZ + i];
            if (k == 0 && i >= 8) v = xinit[i - 8];
            if (is_free(k, i)) {
                real lb = w->lb[i], ub = w->ub[i], kp = (real)o->kappa_push;
                real pl = fmin(kp * fmax((real)1, fabs(lb)), kp * (ub - lb));
                real pu = fmin(kp * fmax((real)1, fabs(ub)), kp * (ub - lb));
                v = fmin(fmax(v, lb + pl), ub - pu);
                w->zl[k][i] = (real)o->mu0 / (v - lb);
                w->zu[k][i] = (real)o->mu0 / (ub - v);
                ncomp += 2;
            } else {
                w->zl[k][i] = w->zu[k][i] = 0;
            }
            w->z[k][i] = v;
        }
        for (int i = 0; i < NXI; i++) w->y[k][i] = 0;
        int m = live_rows(w, k);
        for (int j = 0; j < m; j++) {
            const real *a = row(w, k, j);
            real sl = (a[3] + HU) - (a[0] * w->z[k][8] + a[1] * w->z[k][9] + a[2] * w->z[k][10]);
            w->s[k][j] = fmax(sl, (real)o->s_floor);
            w->lc[k][j] = (real)o->mu0 / w->s[k][j];
            ncomp++;
        }
    }
    int cur = 0, flag = 0, it = 0, nbt_total = 0;
    real alpha_p = 0, alpha_d = 0, rs_n = 0, req_n = 0, rin_n = 0, rcomp = 0, mu = 0;
    evaluate(w, w->z, w->s, &w->ev[cur]);
    /* The stage-0 states are fixed by the xinit equality, so their bounds and the stage-0 corridor rows are not part of
     * the barrier problem.  If xinit VIOLATES one of them (beyond TolIneq) the reference's NLP -- which does carry
     * them (mpc_generator_normal.m:29-50) -- has no feasible point; say so in the reference's vocabulary:
     * NOPROGRESS (-7, header :128), zero iterations, the violation in res_ineq. */
    real v0 = 0;
    for (int i = 8; i < NZ; i++) v0 = fmax(v0, fmax(w->lb[i] - w->z[0][i], w->z[0][i] - w->ub[i]));
    {
        int m0 = w->nrows[0] < w->mcap ? w->nrows[0] : w->mcap;
        for (int j = 0; j < m0; j++) {
            const real *a = row(w, 0, j);
            v0 = fmax(v0, a[0] * w->z[0][8] + a[1] * w->z[0][9] + a[2] * w->z[0][10] - (a[3] + HU));
        }
    }
    const int infeasible0 = v0 > (real)o->tol_ineq;
    if (infeasible0) { flag = -7; rin_n = v0; }
    for (it = 0; !infeasible0; it++) {
        eval_t *e = &w->ev[cur];
        /* ---- residuals */
        rs_n = req_n = rin_n = rcomp = 0;
        real csum = 0, cmin = (real)1e30;
        for (int k = 0; k < N; k++) {
            int m = live_rows(w, k);
            for (int i = 0; i < NZ; i++) {
                if (!is_free(k, i)) continue;
                real r = e->g[k][i] - w->zl[k][i] + w->zu[k][i];
                if (k < N - 1) for (int q = 0; q < NXI; q++) r += e->J[k][q][i] * w->y[k + 1][q];
                if (k > 0) {
                    if (i >= 8) r -= w->y[k][i - 8];
                    else if (i >= 4) r -= w->y[k][9 + i - 4];
                }
                if (i >= 8 && i < 11) for (int j = 0; j < m; j++) r += row(w, k, j)[i - 8] * w->lc[k][j];
                rs_n = fmax(rs_n, fabs(r));
                real cl = (w->z[k][i] - w->lb[i]) * w->zl[k][i], cu = (w->ub[i] - w->z[k][i]) * w->zu[k][i];
                csum += cl + cu;
                rcomp = fmax(rcomp, fmax(cl, cu));
                cmin = fmin(cmin, fmin(cl, cu));
            }
            if (k < N - 1) for (int i = 0; i < NXI; i++) req_n = fmax(req_n, fabs(e->d[k][i]));
            for (int j = 0; j < m; j++) {
                real cc = w->s[k][j] * w->lc[k][j];
                csum += cc;
                rcomp = fmax(rcomp, cc);
                cmin = fmin(cmin, cc);
                rin_n = fmax(rin_n, fmax(fabs(e->rc[k][j]), e->rc[k][j] - w->s[k][j]));
            }
        }
        mu = csum / (real)ncomp;
        if (!(isfinite(rs_n) && isfinite(req_n) && isfinite(mu) && isfinite(e->f))) { flag = (it == 0) ? -6 : -7; break; }
        if (rs_n <= o->tol_stat && req_n <= o->tol_eq && rin_n <= o->tol_ineq && rcomp <= o->tol_comp) { flag = 1; break; }
        if (it >= o->maxit) { flag = 0; break; }
        real sigma = (real)o->sigma;
        if (sigma <= 0) { /* LOQO centrality rule: xi = min(s.lambda)/mu */
            real xi = cmin / mu;
            real q = fmin((real)0.05 * (1 - xi) / xi, (real)2);
            sigma = (real)0.1 * q * q * q;
        }
        real mu_t = fmax(sigma * mu, (real)o->mu_floor);
        /* Mehrotra predictor-corrector (o->pc): pass 0 solves with mu = 0 (affine direction), measures how far it can
         * go (mu_aff) and sets sigma = (mu_aff / mu)^3; pass 1 solves with that target and the second-order term
         * ds_aff * dlambda_aff in the complementarity rows (ccl / ccu / ccr).  Without pc only pass 1 runs, with zero terms. */
        for (int k = 0; k < N; k++) { for (int i = 0; i < NZ; i++) w->ccl[k][i] = w->ccu[k][i] = 0; for (int j = 0; j < w->mcap; j++) w->ccr[k][j] = 0; }
        for (int pass = o->pc ? 0 : 1; pass < 2 && flag == 0; pass++) {
        const real mu_use = (pass == 0) ? (real)0 : mu_t;
        /* ---- KKT blocks */
        for (int k = 0; k < N; k++) {
            const real *hdr = w->hdr + k * 10;
            int m = live_rows(w, k);
            cost_hessian(hdr, k, N, w->variant, w->Phi[k]);
            for (int i = 0; i < NZ; i++) {
                if (is_free(k, i)) {
                    real sl = w->z[k][i] - w->lb[i], su = w->ub[i] - w->z[k][i];
                    w->Phi[k][i][i] += w->zl[k][i] / sl + w->zu[k][i] / su;
                    real gi = e->g[k][i];
                    if (o->mixed) {   /* delta form: the gradient of the Lagrangian at the current y */
                        if (k < N - 1) for (int q = 0; q < NXI; q++) gi += e->J[k][q][i] * w->y[k + 1][q];
                        if (k > 0) {
                            if (i >= 8) gi -= w->y[k][i - 8];
                            else if (i >= 4) gi -= w->y[k][9 + i - 4];
                        }
                    }
                    w->gt[k][i] = gi - (mu_use - w->ccl[k][i]) / sl + (mu_use - w->ccu[k][i]) / su;
                } else {
                    for (int j = 0; j < NZ; j++) w->Phi[k][i][j] = w->Phi[k][j][i] = 0;
                    w->Phi[k][i][i] = 1;
                    w->gt[k][i] = 0;
                }
            }
            for (int j = 0; j < m; j++) {
                const real *a = row(w, k, j);
                real sg = w->lc[k][j] / w->s[k][j];
                real tt = (mu_use - w->ccr[k][j] + w->lc[k][j] * e->rc[k][j]) / w->s[k][j];
                for (int p = 0; p < 3; p++) {
                    w->gt[k][8 + p] += a[p] * tt;
                    for (int q = 0; q < 3; q++) w->Phi[k][8 + p][8 + q] += a[p] * sg * a[q];
                }
            }
            if (k < N - 1) {
                memcpy(w->Cm[k], e->J[k], sizeof(real) * NXI * NZ);
                if (k == 0) for (int r = 0; r < NXI; r++) for (int i = 8; i < NZ; i++) w->Cm[0][r][i] = 0;
            }
        }
        if (o->mixed) {
            if ((o->mixed == 2 ? riccati_solve_d : riccati_solve_s)(N, w->Phi, w->gt, w->Cm, e->d, w->dz, w->yn)) { flag = -5; break; }
            for (int k = 0; k < N; k++) for (int i = 0; i < NXI; i++) w->yn[k][i] += w->y[k][i];   /* y + dy */
        } else if (kkt_solve(N, w->Phi, w->gt, w->Cm, e->d, w->dz, w->yn)) { flag = -5; break; }
        if (flag != 0) break;
        if (pass == 0) {
            real ap = 1, ad = 1;
            for (int k = 0; k < N; k++) {
                int m = live_rows(w, k);
                for (int i = 0; i < NZ; i++) {
                    if (!is_free(k, i)) continue;
                    real sl = w->z[k][i] - w->lb[i], su = w->ub[i] - w->z[k][i], dzi = w->dz[k][i];
                    real dl = (-w->zl[k][i] * dzi) / sl - w->zl[k][i], du = (w->zu[k][i] * dzi) / su - w->zu[k][i];
                    if (dzi < 0) ap = fmin(ap, -sl / dzi);
                    if (dzi > 0) ap = fmin(ap, su / dzi);
                    if (dl < 0) ad = fmin(ad, -w->zl[k][i] / dl);
                    if (du < 0) ad = fmin(ad, -w->zu[k][i] / du);
                    w->ccl[k][i] = dzi * dl; w->ccu[k][i] = -dzi * du;
                }
                for (int j = 0; j < m; j++) {
                    const real *a = row(w, k, j);
                    real dsj = -e->rc[k][j] - (a[0] * w->dz[k][8] + a[1] * w->dz[k][9] + a[2] * w->dz[k][10]);
                    real dl = (-w->lc[k][j] * dsj) / w->s[k][j] - w->lc[k][j];
                    if (dsj < 0) ap = fmin(ap, -w->s[k][j] / dsj);
                    if (dl < 0) ad = fmin(ad, -w->lc[k][j] / dl);
                    w->ccr[k][j] = dsj * dl;
                }
            }
            /* mu_aff = sum (s + ap ds)(lam + ad dlam) / n = (S0 + ap S1 + ad S2 + ap ad S3) / n, S3 = sum of the cc terms */
            real S1 = 0, S2 = 0, S3 = 0;
            for (int k = 0; k < N; k++) {
                int m = live_rows(w, k);
                for (int i = 0; i < NZ; i++) {
                    if (!is_free(k, i)) continue;
                    real sl = w->z[k][i] - w->lb[i], su = w->ub[i] - w->z[k][i], dzi = w->dz[k][i];
                    real dl = (-w->zl[k][i] * dzi) / sl - w->zl[k][i], du = (w->zu[k][i] * dzi) / su - w->zu[k][i];
                    S1 += dzi * w->zl[k][i] - dzi * w->zu[k][i];
                    S2 += sl * dl + su * du;
                    S3 += w->ccl[k][i] + w->ccu[k][i];
                }
                for (int j = 0; j < m; j++) {
                    const real *a = row(w, k, j);
                    real dsj = -e->rc[k][j] - (a[0] * w->dz[k][8] + a[1] * w->dz[k][9] + a[2] * w->dz[k][10]);
                    real dl = (-w->lc[k][j] * dsj) / w->s[k][j] - w->lc[k][j];
                    S1 += dsj * w->lc[k][j]; S2 += w->s[k][j] * dl; S3 += w->ccr[k][j];
                }
            }
            real mu_aff = (csum + ap * S1 + ad * S2 + ap * ad * S3) / (real)ncomp, sg = mu_aff / mu;
            mu_t = fmax(sg * sg * sg * mu, (real)o->mu_floor);
        }
        }
        if (flag != 0) break;
        /* ---- multiplier / slack steps and fraction to the boundary */
        real tau = fmin(fmax((real)0.995, 1 - mu), (real)0.99999);
        real ap = 1, ad = 1;
        for (int k = 0; k < N; k++) {
            int m = live_rows(w, k);
            for (int i = 0; i < NZ; i++) {
                if (!is_free(k, i)) { w->dzl[k][i] = w->dzu[k][i] = 0; continue; }
                real sl = w->z[k][i] - w->lb[i], su = w->ub[i] - w->z[k][i], dzi = w->dz[k][i];
                w->dzl[k][i] = (mu_t - w->ccl[k][i] - w->zl[k][i] * dzi) / sl - w->zl[k][i];
                w->dzu[k][i] = (mu_t - w->ccu[k][i] + w->zu[k][i] * dzi) / su - w->zu[k][i];
                if (dzi < 0) ap = fmin(ap, -tau * sl / dzi);
                if (dzi > 0) ap = fmin(ap, tau * su / dzi);
                if (w->dzl[k][i] < 0) ad = fmin(ad, -tau * w->zl[k][i] / w->dzl[k][i]);
                if (w->dzu[k][i] < 0) ad = fmin(ad, -tau * w->zu[k][i] / w->dzu[k][i]);
            }
            for (int j = 0; j < m; j++) {
                const real *a = row(w, k, j);
                real dsj = -e->rc[k][j] - (a[0] * w->dz[k][8] + a[1] * w->dz[k][9] + a[2] * w->dz[k][10]);
                real dl = (mu_t - w->ccr[k][j] - w->lc[k][j] * dsj) / w->s[k][j] - w->lc[k][j];
                w->ds[k][j] = dsj;
                w->dlc[k][j] = dl;
                if (dsj < 0) ap = fmin(ap, -tau * w->s[k][j] / dsj);
                if (dl < 0) ad = fmin(ad, -tau * w->lc[k][j] / dl);
            }
        }
        /* ---- backtracking line search on (theta, barrier objective) */
        real th0 = e->theta, ph0 = e->f - mu_t * e->logsum;
        /* theta below 1% of TolEq counts as feasible (also absorbs the rounding floor of theta) */
        real th_noise = fmax(10 * eps * (real)(N * NXI) * 20, (real)0.01 * (real)o->tol_eq);
        real a = ap;
        int nbt = 0;
        eval_t *tr = &w->ev[1 - cur];
        for (;;) {
            for (int k = 0; k < N; k++) {
                for (int i = 0; i < NZ; i++) w->zt[k][i] = w->z[k][i] + a * w->dz[k][i];
                int m = live_rows(w, k);
                for (int j = 0; j < m; j++) w->st[k][j] = w->s[k][j] + a * w->ds[k][j];
            }
            evaluate(w, w->zt, w->st, tr);
            real pht = tr->f - mu_t * tr->logsum;
            int ok = (tr->theta <= fmax((1 - (real)1e-5) * th0, th_noise)) ||
                     (pht <= ph0 - (real)1e-5 * th0 + 10 * eps * fabs(ph0));
            if (ok || nbt >= o->max_bt) break;
            nbt++;
            a *= (real)0.5;
        }
        nbt_total += nbt;
        alpha_p = a; alpha_d = ad;
        for (int k = 0; k < N; k++) {
            int m = live_rows(w, k);
            for (int i = 0; i < NZ; i++) {
                w->z[k][i] = w->zt[k][i];
                w->zl[k][i] += ad * w->dzl[k][i];
                w->zu[k][i] += ad * w->dzu[k][i];
            }
            for (int j = 0; j < m; j++) { w->s[k][j] = w->st[k][j]; w->lc[k][j] += ad * w->dlc[k][j]; }
            for (int i = 0; i < NXI; i++) w->y[k][i] += a * (w->yn[k][i] - w->y[k][i]);
        }
        cur = 1 - cur;
    }
    for (int k = 0; k < N; k++) for (int i = 0; i < NZ; i++) z_out[k * NZ + i] = w->z[k][i];
    for (int k = 0; k < N; k++) {
        if (y_out) for (int i = 0; i < NXI; i++) y_out[k * NXI + i] = k ? w->y[k][i] : 0;
        if (zl_out) for (int i = 0; i < NZ; i++) zl_out[k * NZ + i] = w->zl[k][i];
        if (zu_out) for (int i = 0; i < NZ; i++) zu_out[k * NZ + i] = w->zu[k][i];
        if (lc_out) for (int j = 0; j < w->mcap; j++) lc_out[k * w->mcap + j] = j < live_rows(w, k) ? w->lc[k][j] : 0;
    }
    iinfo[0] = flag; iinfo[1] = it; iinfo[2] = nbt_total; iinfo[3] = 0;
    rinfo[0] = req_n; rinfo[1] = rin_n; rinfo[2] = rs_n; rinfo[3] = rcomp;
    rinfo[4] = w->ev[cur].f; rinfo[5] = mu; rinfo[6] = alpha_p; rinfo[7] = alpha_d;
    return flag;
}

int nmpc_oracle_solve_batch(int B, int N, int mcap, const real *xinit, const real *z0, const real *hdr,
                            const real *rows, const int *nrows, int variant, const nmpc_oracle_opts *opts,
                            real *z_out, int *info_int, real *info_real, int nthreads)
{
    return nmpc_oracle_solve_batch_ex(B, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_out, info_int,
                                      info_real, 0, 0, 0, 0, nthreads);
}

int nmpc_oracle_solve_batch_ex(int B, int N, int mcap, const real *xinit, const real *z0, const real *hdr,
                               const real *rows, const int *nrows, int variant, const nmpc_oracle_opts *opts,
                               real *z_out, int *info_int, real *info_real, real *y_out, real *zl_out,
                               real *zu_out, real *lc_out, int nthreads)
{
    if (N < 2 || N > NS_MAX || mcap < 0 || mcap > MC_MAX) return -11;
    nmpc_oracle_opts o;
    if (opts) o = *opts; else nmpc_oracle_default_opts(&o);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        work_t *w = (work_t *)malloc(sizeof(work_t));
#pragma omp for schedule(dynamic, 4)
        for (int b = 0; b < B; b++) {
            w->N = N; w->mcap = mcap; w->variant = variant;
            w->hdr = hdr + (size_t)b * N * 10;
            w->rows = rows + (size_t)b * N * mcap * 4;
            w->nrows = nrows + (size_t)b * N;
            solve_one(w, &o, xinit + (size_t)b * 9, z0 + (size_t)b * N * NZ, z_out + (size_t)b * N * NZ,
                      info_int + (size_t)b * 4, info_real + (size_t)b * 8,
                      y_out ? y_out + (size_t)b * N * NXI : 0, zl_out ? zl_out + (size_t)b * N * NZ : 0,
                      zu_out ? zu_out + (size_t)b * N * NZ : 0, lc_out ? lc_out + (size_t)b * N * mcap : 0);
        }
        free(w);
    }
    return 0;
}

/* One thread, one problem after the other, each solve timed on its own (CLOCK_MONOTONIC): how the planner runs the
 * reference solver (num_of_threads = 1, plan_manage/src/forces_normal.cpp:31).  seconds [B]. */
int nmpc_oracle_solve_batch_timed(int B, int N, int mcap, const real *xinit, const real *z0, const real *hdr,
                                  const real *rows, const int *nrows, int variant, const nmpc_oracle_opts *opts,
                                  real *z_out, int *info_int, real *info_real, double *seconds)
{
    if (N < 2 || N > NS_MAX || mcap < 0 || mcap > MC_MAX || !seconds) return -11;
    nmpc_oracle_opts o;
    if (opts) o = *opts; else nmpc_oracle_default_opts(&o);
    work_t *w = (work_t *)malloc(sizeof(work_t));
    for (int b = 0; b < B; b++) {
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        w->N = N; w->mcap = mcap; w->variant = variant;
        w->hdr = hdr + (size_t)b * N * 10;
        w->rows = rows + (size_t)b * N * mcap * 4;
        w->nrows = nrows + (size_t)b * N;
        solve_one(w, &o, xinit + (size_t)b * 9, z0 + (size_t)b * N * NZ, z_out + (size_t)b * N * NZ,
                  info_int + (size_t)b * 4, info_real + (size_t)b * 8, 0, 0, 0, 0);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        seconds[b] = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    }
    free(w);
    return 0;
}

/* The restated model behind the reference callback's own signature (FORCESNLPsolver_normal_extfunc): what kkt_check.c is
 * driven with where oracle/_ref is not available.  normal variant / final variant. */
static void compat_cb(double *x, double *p, double *f, double *nabla_f, double *c, double *nabla_c, double *h, double *nabla_h,
                      int stage, int variant)
{
    double ff;
    nmpc_oracle_model_eval(x, p, stage, 20, variant, &ff, nabla_f, c, nabla_c, h, nabla_h);
    *f += ff;
}
void nmpc_oracle_casadi2forces_normal(double *x, double *y, double *l, double *p, double *f, double *nabla_f, double *c, double *nabla_c,
                                      double *h, double *nabla_h, double *hess, int stage, int iteration, int thread)
{
    (void)y; (void)l; (void)hess; (void)iteration; (void)thread;
    compat_cb(x, p, f, nabla_f, c, nabla_c, h, nabla_h, stage, 0);
}
void nmpc_oracle_casadi2forces_final(double *x, double *y, double *l, double *p, double *f, double *nabla_f, double *c, double *nabla_c,
                                     double *h, double *nabla_h, double *hess, int stage, int iteration, int thread)
{
    (void)y; (void)l; (void)hess; (void)iteration; (void)thread;
    compat_cb(x, p, f, nabla_f, c, nabla_c, h, nabla_h, stage, 1);
}
