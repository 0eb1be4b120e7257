/*
 * kkt_check.c -- ForcesPro's acceptance test, re-evaluated for a whole batch with a model CALLBACK.
 * TEST INFRASTRUCTURE ONLY (tests/helpers.py); never part of the product.
 *
 * The callback has the reference's own signature (FORCESNLPsolver_normal_extfunc,
 * /root/reference/src/resilient_planner/plan_manage/solver/normal/FORCESNLPsolver_normal/include/FORCESNLPsolver_normal.h:321)
 * and is, in the tests, the reference's own FORCESNLPsolver_{normal,final}_casadi2forces out of oracle/_ref (compiled from
 * the reference sources where they lie).  For every problem of the batch this computes, at a candidate (z, y, z_l, z_u,
 * lambda), the four inf-norms ForcesPro stops on (TolStat / TolEq / TolIneq / TolComp = 1e-4,
 * matlab_code/mpc/normal/mpc_generator_normal.m:76-79) plus the two facts a sign-blind residual would miss:
 *   res[b] = { stationarity, equality, inequality, complementarity (max |slack * multiplier|), smallest multiplier, cost }
 * Parameter slots follow matlab_code/setup.m:60-66 (130 per stage); N = 20 only (the callbacks dispatch on stage 0 / 1-18 / 19).
 * The stage-0 states are fixed by the xinit equality (its multiplier absorbs their stationarity rows).
 */
#include <math.h>
#include <string.h>

#define NZ 17
#define NXI 13
#define NH 30
#define HU 1e-5
#define PI_D 3.14159265358979323846

typedef void (*extfunc)(double *x, double *y, double *l, double *p, double *f, double *nabla_f, double *c, double *nabla_c,
                        double *h, double *nabla_h, double *hess, int stage, int iteration, int thread);

static const double LB[NZ] = {-PI_D / 2, -PI_D / 2, -PI_D / 2, 0.5 * 9.81 * 0.745319, -PI_D / 2, -PI_D / 2, -PI_D / 2, 0.5 * 9.81 * 0.745319,
                              -20, -20, 0, -2, -2, -2, -0.4 * PI_D, -0.4 * PI_D, -2 * PI_D};
static const double UB[NZ] = {PI_D / 2, PI_D / 2, PI_D / 2, 2.0 * 9.81 * 0.745319, PI_D / 2, PI_D / 2, PI_D / 2, 2.0 * 9.81 * 0.745319,
                              20, 20, 5, 2, 2, 2, 0.4 * PI_D, 0.4 * PI_D, 2 * PI_D};

static inline int e_col(int i) { return i < 9 ? 8 + i : i - 5; }

int nmpc_kkt_check(extfunc fn, int B, int N, int mcap, const double *xinit, const double *hdr, const double *rows, const int *nrows,
                   const double *z, const double *y, const double *zl, const double *zu, const double *lc, double *res)
{
    if (!fn || N != 20 || mcap < 0 || mcap > NH) return -11;
    static double g[20][NZ], c[20][NXI], J[20][NXI * NZ], h[20][NH], Jh[20][NH * NZ];
    for (int b = 0; b < B; b++) {
        const double *zb = z + (size_t)b * N * NZ, *yb = y + (size_t)b * N * NXI, *zlb = zl + (size_t)b * N * NZ, *zub = zu + (size_t)b * N * NZ;
        const double *lcb = lc + (size_t)b * N * mcap;
        double rs = 0, req = 0, rin = 0, rcomp = 0, mmin = 1e300, cost = 0;
        int m_of[20];
        for (int k = 0; k < N; k++) {
            double p[130], x[NZ], yy[NXI] = {0}, ll[NH] = {0}, f = 0;
            memset(p, 0, sizeof p);
            memcpy(p, hdr + ((size_t)b * N + k) * 10, 10 * sizeof(double));
            int m = nrows[(size_t)b * N + k];
            m = m < mcap ? m : mcap;
            m_of[k] = m;
            for (int j = 0; j < m; j++) {
                const double *r = rows + (((size_t)b * N + k) * mcap + j) * 4;
                p[10 + 3 * j] = r[0]; p[11 + 3 * j] = r[1]; p[12 + 3 * j] = r[2]; p[100 + j] = r[3];
            }
            memcpy(x, zb + k * NZ, sizeof x);
            memset(g[k], 0, sizeof g[k]); memset(c[k], 0, sizeof c[k]); memset(J[k], 0, sizeof J[k]);
            memset(h[k], 0, sizeof h[k]); memset(Jh[k], 0, sizeof Jh[k]);
            fn(x, yy, ll, p, &f, g[k], c[k], J[k], h[k], Jh[k], 0, k, 0, 0);
            cost += f;
        }
        for (int i = 0; i < 9; i++) req = fmax(req, fabs(zb[8 + i] - xinit[(size_t)b * 9 + i]));
        for (int k = 0; k < N; k++) {
            const int m = k > 0 ? m_of[k] : 0;
            for (int i = 0; i < NZ; i++) {
                if (k == 0 && i >= 8) continue;
                double r = g[k][i] - zlb[k * NZ + i] + zub[k * NZ + i];
                if (k < N - 1) for (int q = 0; q < NXI; q++) r += J[k][i * NXI + q] * yb[(k + 1) * NXI + q];   /* column-major 13x17 */
                if (k > 0) {
                    if (i >= 8) r -= yb[k * NXI + i - 8];
                    else if (i >= 4) r -= yb[k * NXI + 5 + i];
                    for (int j = 0; j < m; j++) r += Jh[k][i * NH + j] * lcb[k * mcap + j];                    /* column-major 30x17 */
                }
                rs = fmax(rs, fabs(r));
                const double v = zb[k * NZ + i];
                rin = fmax(rin, fmax(LB[i] - v, v - UB[i]));
                rcomp = fmax(rcomp, fmax(fabs((v - LB[i]) * zlb[k * NZ + i]), fabs((UB[i] - v) * zub[k * NZ + i])));
                mmin = fmin(mmin, fmin(zlb[k * NZ + i], zub[k * NZ + i]));
            }
            if (k < N - 1) for (int q = 0; q < NXI; q++) req = fmax(req, fabs(c[k][q] - zb[(k + 1) * NZ + e_col(q)]));
            for (int j = 0; j < m; j++) {
                rin = fmax(rin, h[k][j] - HU);
                rcomp = fmax(rcomp, fabs((HU - h[k][j]) * lcb[k * mcap + j]));
                mmin = fmin(mmin, lcb[k * mcap + j]);
            }
        }
        double *o = res + (size_t)b * 6;
        o[0] = rs; o[1] = req; o[2] = fmax(rin, 0.0); o[3] = rcomp; o[4] = mmin; o[5] = cost;
    }
    return 0;
}
