// nmpc_ipm_mixed.cuh -- mixed-precision fused interior-point NMPC solve, one warp (= one CTA) per problem.
//
// Same drop-in role as nmpc_ipm.cuh (the whole of FORCESNLPsolver_{normal,final}_solve,
// /root/reference/src/resilient_planner/plan_manage/solver/normal/FORCESNLPsolver_normal/include/FORCESNLPsolver_normal.h:321-323),
// for BASELINE configs 3 and 4 ("fp32"): it meets the REFERENCE tolerances (1e-4 inf-norms,
// matlab_code/mpc/normal/mpc_generator_normal.m:76-79) that a pure single-precision solver cannot resolve
// (stationarity 1e-4 absolute on gradients of order 1e2..1e3), at close to single-precision cost:
//
//   double precision  the iterate (z, y, z_l, z_u, s, lambda), the model evaluation, the KKT residuals,
//                     the fraction-to-boundary rule, the line-search merit -- everything "lanes = stages";
//   single precision  the Newton system: stage Hessians, compact Jacobians, right-hand side, the whole Riccati
//                     recursion (cost-to-go, gains), the forward rollout and the costate sweep.
//
// The Newton system is written in DELTA form: its right-hand side is the full KKT residual at the current
// point -- the stationarity residual r = grad f + J'y - E'y - z_l + z_u + A'lambda is evaluated in double
// precision while the Jacobian of the accepted trial point is still in registers, its inf-norm taken there, and
// only then is it rounded (a residual may be rounded: the error is relative to r itself) -- so the unknowns are
// (dz, dy) and the rounding of the single-precision solve is relative to the STEP, not to the multipliers.
// For r to be available at the trial point the multiplier steps (which do not depend on the primal step
// length) are taken BEFORE the line search; nothing else of the algorithm changes.  The outer Newton iteration
// thereby acts as iterative refinement; on BASELINE configs 2 / 3 the iteration counts equal the fp64
// solver's (oracle restatement: oracle/nmpc_oracle.c, opts.mixed = 1).
//
// Single precision loses the cost-to-go's positive definiteness on rare, badly scaled instances (barrier
// terms lambda/s ~ 1e8 on a state that the previous stage's control can fully absorb: catastrophic
// cancellation in P = Q_xx - Y'Y).  Those warps stop with the reference's factorisation code (-5) and the
// host entry point re-solves exactly those problems with the fp64 kernel (nmpc_capi.cu: solve_mixed).
//
// Shared memory per problem (N = 20, 8 rows): 15.2 KB of fp64 state + 10.6 KB of fp32 Newton data; the
// fp32 sweep-private arrays (gains, Riccati scratch: 7.8 KB) are overlaid on the fp64 state, which is parked
// in 31 registers per lane across the sweeps -- the same trick as the fp64 kernel at half its register cost.
#pragma once
#include "nmpc_ipm.cuh"

namespace nmpc {

struct MixedParams {
    int B, mcap, variant, io32;   // io32: problem data and results are float arrays in HBM (else double)
    const void* xinit;    // [B][9]
    const void* z0;       // [B][N][17]
    const void* hdr;      // [B][N][10]
    const void* rows;     // [B][N][mcap][4]
    const int* nrows;     // [B][N]
    const int* order;     // [B] or nullptr
    void* z_out;          // [B][N][17]
    int* info_int;        // [B][4]  exitflag, iterations, backtracks, 0 (1 = re-solved by the fp64 kernel)
    void* info_real;      // [B][8]
    void *y_out, *zl_out, *zu_out, *lc_out;   // optional multipliers, same element type as the problem data
    PeerOut peers;        // multi-GPU collation by peer stores (nmpc_ipm.cuh); n = 0: single GPU
    Opts o;
};

constexpr int MIXED_BAIL_IT = 60;   // a mixed solve still running after this many iterations is handed to the fp64 kernel

template <int N> struct MLayout {
    static_assert(N % 4 == 0 && N >= 4 && N <= 64, "horizon must be a multiple of 4 (TMA 16-byte granules)");
    using L32 = Layout<float, N, false>;
    static constexpr int HDR_S = L32::HDR_S, PHI_S = L32::PHI_S;
    static constexpr int HEAD_BYTES = 16 + N * 4;   // mbarrier + nrows
    // ---- fp64 state R, offsets in doubles ----
    static constexpr int Z = 0;
    static constexpr int ZL = Z + N * NZ;
    static constexpr int ZU = ZL + N * NZ;
    static constexpr int Y = ZU + N * NZ;
    static constexpr int HDR = Y + N * NXI;
    static constexpr int BND = HDR + N * HDR_S;
    static constexpr int R_FIXED = BND + 2 * NZ;
    __host__ __device__ static constexpr int s_stride(int mcap) { return mcap | 1; }
    __host__ __device__ static constexpr int s_off(int) { return R_FIXED; }
    __host__ __device__ static constexpr int lc_off(int mcap) { return R_FIXED + N * s_stride(mcap); }
    __host__ __device__ static constexpr int r_end(int mcap) { return R_FIXED + 2 * N * s_stride(mcap); }
    // ---- fp32 overlay at the start of R: the sweep-private arrays of Solver<float, N> ----
    static constexpr int KG32 = L32::KFF + N * 4;            // gains always overlaid here (floats)
    static constexpr int O_END32 = KG32 + N * 52;
    static constexpr int NPARK_LANE = (O_END32 + 63) / 64;   // doubles parked per lane
    static constexpr int NPARK = NPARK_LANE * 32;
    static_assert(NPARK_LANE <= 64, "overlay does not fit the parking registers");
    // ---- fp32 Newton data SH: after max(R, parked overlay), 16-byte aligned ----
    __host__ __device__ static constexpr int sh_off_d(int mcap)
    {
        return ((r_end(mcap) > NPARK ? r_end(mcap) : NPARK) + 1) & ~1;
    }
    static constexpr int SH_DZ = 0;
    static constexpr int SH_G = SH_DZ + N * NZ;      // stationarity residual of the trial point, then the right-hand side
    static constexpr int SH_DY = SH_G + N * NZ;      // p_k during the backward sweep, then dy (costates of the QP)
    static constexpr int SH_D = SH_DY + N * NXI;
    static constexpr int SH_JC = SH_D + N * NXI;
    static constexpr int SH_PHID = SH_JC + N * NJC;
    static constexpr int SH_END = SH_PHID + N * PHI_S;
    // TMA staging inside SH (dead until the first evaluation): z0 then hdr, in the I/O element type
    static constexpr int STG_HDR_BYTES = N * NZ * 8;
    static_assert(SH_END * 4 >= N * (NZ + 10) * 8, "staging area");
    __host__ __device__ static constexpr size_t bytes(int mcap)
    {
        return (size_t)HEAD_BYTES + (size_t)sh_off_d(mcap) * 8 + (size_t)SH_END * 4;
    }
    static constexpr int NT = (N * NZ + 31) / 32;
};

template <int N> struct MixedSolver {
    using ML = MLayout<N>;
    using L32 = typename ML::L32;
    using C = Const<double>;
    double* r64;   // base of the fp64 state
    int* nr;
    int lane, mcap, SS;
    const void* rows_g;
    bool io32, final_variant;
    double *Z, *ZL, *ZU, *Y, *HDR, *BND, *S, *LC;
    float *DZ, *G, *DY, *D, *JC, *PHID;
    Solver<float, N, false> sw;   // the single-precision sweeps, bound to the overlay + SH

    __device__ __forceinline__ void bind(unsigned char* smem_raw, int lane_, int mcap_)
    {
        r64 = reinterpret_cast<double*>(smem_raw + ML::HEAD_BYTES);
        nr = reinterpret_cast<int*>(smem_raw + 16);
        lane = lane_; mcap = mcap_; SS = ML::s_stride(mcap);
        Z = r64 + ML::Z; ZL = r64 + ML::ZL; ZU = r64 + ML::ZU; Y = r64 + ML::Y; HDR = r64 + ML::HDR; BND = r64 + ML::BND;
        S = r64 + ML::s_off(mcap); LC = r64 + ML::lc_off(mcap);
        float* sh = reinterpret_cast<float*>(r64 + ML::sh_off_d(mcap));
        DZ = sh + ML::SH_DZ; G = sh + ML::SH_G; DY = sh + ML::SH_DY; D = sh + ML::SH_D; JC = sh + ML::SH_JC; PHID = sh + ML::SH_PHID;
        float* ovl = reinterpret_cast<float*>(r64);
        sw.sm = ovl; sw.nr = nr; sw.lane = lane; sw.mcap = mcap; sw.SS = SS; sw.rows_g = nullptr; sw.final_variant = false;
        sw.DZ = DZ; sw.G = G; sw.P = DY; sw.D = D; sw.JC = JC; sw.PHID = PHID;
        sw.KG = ovl + ML::KG32; sw.KFF = ovl + L32::KFF;
        sw.Z = sw.ZL = sw.ZU = sw.Y = sw.HDR = sw.S = sw.LC = sw.BND = nullptr;
        sw.QINV = sw.PQQ0 = sw.DZAP = nullptr;
        sw.fac_out = nullptr;
    }
    __device__ __forceinline__ void park(double (&regs)[ML::NPARK_LANE]) const
    {
#pragma unroll
        for (int t = 0; t < ML::NPARK_LANE; t++) regs[t] = r64[lane + 32 * t];
    }
    __device__ __forceinline__ void unpark(const double (&regs)[ML::NPARK_LANE]) const
    {
#pragma unroll
        for (int t = 0; t < ML::NPARK_LANE; t++) r64[lane + 32 * t] = regs[t];
    }
    __device__ __forceinline__ int live(int k) const { return k == 0 ? 0 : min(nr[k], mcap); }
    __device__ __forceinline__ void load_row(int k, int j, double (&r)[4]) const
    {
        const size_t idx = (size_t)(k * mcap + j) * 4;
        if (io32) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(rows_g) + idx));
            r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w;
        } else {
            const double2* p = reinterpret_cast<const double2*>(static_cast<const double*>(rows_g) + idx);
            const double2 a = __ldg(p), c = __ldg(p + 1);
            r[0] = a.x; r[1] = a.y; r[2] = c.x; r[3] = c.y;
        }
    }
    __device__ __forceinline__ static void frac_max(double& bn, double& bd, double n, double d)
    {
        if (n * bd > bn * d) { bn = n; bd = d; }
    }

    // ---------------------------------------------------------------- model evaluation ---
    // At (z + a dz, y + a dy) with the multipliers already stepped, all in double precision ("lanes = stages"):
    // cost, defects, theta, barrier log-sum, and the stationarity residual formed while the Jacobian is in
    // registers.  What the single-precision Newton solve needs (compact Jacobian, defects, residual) is rounded on
    // the way out; the norms that decide termination are taken before the rounding.
    __device__ void evaluate(double a, double& f_out, double& th_out, double& ls_out, double& req_out, double& rs_out)
    {
        double f = 0.0, th = 0.0, ls = 0.0, rq = 0.0, rs = 0.0;
        for (int k = lane; k < N; k += 32) {
            double zk[NZ], g[NZ];
#pragma unroll
            for (int i = 0; i < NZ; i++) zk[i] = Z[k * NZ + i] + a * (double)DZ[k * NZ + i];
            const double* hdr = HDR + k * ML::HDR_S;
            f += objective<double, true>(zk, hdr, k == 0, final_variant && k == N - 1, g);
            if (k < N - 1) {
                double c[NXI], jc[NJC], yn[NXI];
                dynamics<double, true>(zk, hdr + 3, c, jc);
#pragma unroll
                for (int i = 0; i < NXI; i++) {
                    const int zi = (k + 1) * NZ + e_col(i);
                    const double d = c[i] - (Z[zi] + a * (double)DZ[zi]);
                    th += fabs(d);
                    rq = fmax(rq, fabs(d));
                    D[k * NXI + i] = (float)d;
                    yn[i] = Y[(k + 1) * NXI + i] + a * (double)DY[(k + 1) * NXI + i];
                }
#pragma unroll
                for (int e = 0; e < NJC; e++) JC[k * NJC + e] = (float)jc[e];
#pragma unroll
                for (int i = 0; i < NZ; i++) g[i] += jt_y<double>(jc, yn, i);
            }
            if (k > 0) {
#pragma unroll
                for (int i = 0; i < NXI; i++) g[i < 9 ? 8 + i : i - 5] -= Y[k * NXI + i] + a * (double)DY[k * NXI + i];
            }
            double prod = 1.0;
#pragma unroll
            for (int i = 0; i < NZ; i++) {
                double sl = zk[i] - lower_bound<double>(i), su = upper_bound<double>(i) - zk[i];
                if (i >= 8 && k == 0) { sl = 1.0; su = 1.0; }
                prod *= sl * su;
                if (i % 9 == 8 || i == NZ - 1) { ls += log(prod); prod = 1.0; }     // <= 18 slacks per product: no underflow
            }
            const int m = live(k);
            double al0 = 0.0, al1 = 0.0, al2 = 0.0;
            for (int j = 0; j < m; j++) {
                double r[4]; load_row(k, j, r);
                double sj = S[k * SS + j];
                const double lj = LC[k * SS + j];
                al0 += r[0] * lj; al1 += r[1] * lj; al2 += r[2] * lj;
                double rc = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + sj;
                const double adz = r[0] * (double)DZ[k * NZ + 8] + r[1] * (double)DZ[k * NZ + 9] + r[2] * (double)DZ[k * NZ + 10];
                sj += a * (-rc - adz);
                rc *= (1.0 - a);
                th += fabs(rc);
                prod *= sj;
                if ((j & 7) == 7) { ls += log(prod); prod = 1.0; }
            }
            ls += log(prod);
            g[8] += al0; g[9] += al1; g[10] += al2;
#pragma unroll
            for (int i = 0; i < NZ; i++) {
                double r = g[i] - ZL[k * NZ + i] + ZU[k * NZ + i];
                if (i >= 8 && k == 0) r = 0.0;                    // stage-0 states are fixed by the xinit equality
                rs = fmax(rs, fabs(r));
                G[k * NZ + i] = (float)r;
            }
        }
        f_out = warp_sum(f); th_out = warp_sum(th); ls_out = warp_sum(ls); req_out = warp_max(rq); rs_out = warp_max(rs);
    }

    // ------------------- accept the primal step, measure the new point, assemble the next Newton system ------
    // One pass over the corridor rows and one over the (stage, variable) pairs does what used to be three phases:
    //   * the accepted primal step: s += a ds, z += a dz, y += a dy (the multipliers were stepped before the line search);
    //   * the residuals of the NEW point that are not evaluate()'s: inequality residual, complementarity products, their sum / max / min;
    //   * the barrier-augmented stage Hessians, and the right-hand side in a form that does not need the barrier target
    //     yet:   rhs = [r + (z_l - z_u) + A' lambda (r_c - s)/s]  +  mu_t [1/s_u - 1/s_l + A'(1/s)]  =  G + mu_t * T.
    //     (mu_t follows from the complementarity sum this very pass produces.)  T is parked in the dead step array DZ,
    //     the rows' partial sums in the dead dy array; finish_rhs(mu_t) adds mu_t * T afterwards.
    // Both brackets are large and cancel only along directions whose Hessian entry (z/s, lambda/s) is larger still, so
    // rounding them separately to single precision moves the step by far less than the stopping tolerance.
    // a = 0 with dz = dy = 0 is the initial point.
    __device__ void post_step(double a, double& rin_n, double& rcomp, double& csum, double& cmin)
    {
        double rin = 0.0, cmx = 0.0, cs = 0.0, cmn = 1e30;
        for (int e = NXI + lane; e < N * NXI; e += 32) Y[e] += a * (double)DY[e];
        __syncwarp();                                          // dy is dead from here: rows park their sums in it
        for (int k = lane; k < N; k += 32) {
            float* phi = PHID + k * ML::PHI_S;
            float* tmp = DY + k * NXI;                         // d0 d1 d2 | t0 t1 t2 (coefficient of mu_t) | g0 g1 g2
            double o01 = 0.0, o02 = 0.0, o12 = 0.0, d0 = 0.0, d1 = 0.0, d2 = 0.0, t0 = 0.0, t1 = 0.0, t2 = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0;
            const int m = live(k);
            for (int j = 0; j < m; j++) {
                double r[4]; load_row(k, j, r);
                const double so = S[k * SS + j], lj = LC[k * SS + j];
                const double rco = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + so;
                const double ds = -rco - (r[0] * (double)DZ[k * NZ + 8] + r[1] * (double)DZ[k * NZ + 9] + r[2] * (double)DZ[k * NZ + 10]);
                const double sj = so + a * ds;
                // residual of the row at the new point from the very numbers that will be in memory (not (1 - a) * rco: with
                // slacks of 1e-9 the rounding of z + a dz and s + a ds is part of what the next Newton step has to remove)
                const double rc = r[0] * (Z[k * NZ + 8] + a * (double)DZ[k * NZ + 8]) + r[1] * (Z[k * NZ + 9] + a * (double)DZ[k * NZ + 9]) +
                                  r[2] * (Z[k * NZ + 10] + a * (double)DZ[k * NZ + 10]) - (r[3] + C::hu) + sj;
                S[k * SS + j] = sj;
                const double cc = sj * lj;
                cs += cc; cmx = fmax(cmx, cc); cmn = fmin(cmn, cc);
                rin = fmax(rin, fmax(fabs(rc), rc - sj));
                const double is = rcp_t(sj), sg = lj * is, tt = lj * (rc - sj) * is;
                d0 += r[0] * r[0] * sg; d1 += r[1] * r[1] * sg; d2 += r[2] * r[2] * sg;
                o01 += r[0] * r[1] * sg; o02 += r[0] * r[2] * sg; o12 += r[1] * r[2] * sg;
                t0 += r[0] * is; t1 += r[1] * is; t2 += r[2] * is;
                g0 += r[0] * tt; g1 += r[1] * tt; g2 += r[2] * tt;
            }
            tmp[0] = (float)d0; tmp[1] = (float)d1; tmp[2] = (float)d2;
            tmp[3] = (float)t0; tmp[4] = (float)t1; tmp[5] = (float)t2;
            tmp[6] = (float)g0; tmp[7] = (float)g1; tmp[8] = (float)g2;
            phi[17] = (float)o01; phi[18] = (float)o02; phi[19] = (float)o12;
            phi[20] = (float)(-2.0 * HDR[k * ML::HDR_S + 8]);
        }
        __syncwarp();
        for (int e = lane; e < N * NZ; e += 32) {
            const int k = e / NZ, i = e - k * NZ;
            float* phi = PHID + k * ML::PHI_S;
            if (e < 8 || e >= NZ) {
                const double zi = Z[e] + a * (double)DZ[e], zl = ZL[e], zu = ZU[e];
                Z[e] = zi;
                const double sl = zi - BND[i], su = BND[NZ + i] - zi;
                const double cl = sl * zl, cu = su * zu;
                cs += cl + cu;
                cmx = fmax(cmx, fmax(cl, cu));
                cmn = fmin(cmn, fmin(cl, cu));
                const double isl = rcp_t(sl), isu = rcp_t(su);
                double ph = cost_hess_diag<double>(i, HDR + k * ML::HDR_S, k == 0, final_variant && k == N - 1) + zl * isl + zu * isu;
                double gp = (double)G[e] + (zl - zu), tc = isu - isl;
                if (i >= 8 && i < 11) {                        // k > 0 here: the rows' sums of this stage's position entries
                    const float* tmp = DY + k * NXI;
                    ph += (double)tmp[i - 8]; tc += (double)tmp[i - 5]; gp += (double)tmp[i - 2];
                }
                phi[i] = (float)ph;
                G[e] = (float)gp;
                DZ[e] = (float)tc;
            } else {                                           // stage-0 states: fixed by the xinit equality
                phi[i] = 1.0f;
                G[e] = 0.0f;
                DZ[e] = 0.0f;
            }
        }
        rin_n = warp_max(rin); rcomp = warp_max(cmx);
        csum = warp_sum(cs); cmin = warp_min(cmn);
    }
    // rhs = G + mu_t * T, in single precision (both are single-precision arrays by now)
    __device__ void finish_rhs(double mu_t)
    {
        const float m = (float)mu_t;
        for (int e = lane; e < N * NZ; e += 32) G[e] = fmaf(m, DZ[e], G[e]);
    }

    // ------------------------------------------- multiplier steps, fraction to boundary ---
    __device__ void step_lengths(double mu_t, double tau, double& ap_out, double& ad_out)
    {
        double pn = 0.0, pd = 1.0, dn = 0.0, dd = 1.0;
        for (int e = lane; e < N * NZ; e += 32) {
            if (!(e < 8 || e >= NZ)) continue;
            const int i = e % NZ;
            const double zi = Z[e], dzi = (double)DZ[e], zl = ZL[e], zu = ZU[e];
            const double sl = zi - BND[i], su = BND[NZ + i] - zi;
            frac_max(pn, pd, -dzi, sl);
            frac_max(pn, pd, dzi, su);
            frac_max(dn, dd, zl * (sl + dzi) - mu_t, sl * zl);
            frac_max(dn, dd, zu * (su - dzi) - mu_t, su * zu);
        }
        for (int k = lane; k < N; k += 32) {
            const int m = live(k);
            for (int j = 0; j < m; j++) {
                double r[4]; load_row(k, j, r);
                const double sj = S[k * SS + j], lj = LC[k * SS + j];
                const double rc = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + sj;
                const double ds = -rc - (r[0] * (double)DZ[k * NZ + 8] + r[1] * (double)DZ[k * NZ + 9] + r[2] * (double)DZ[k * NZ + 10]);
                frac_max(pn, pd, -ds, sj);
                frac_max(dn, dd, lj * (sj + ds) - mu_t, sj * lj);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double n2 = __shfl_xor_sync(0xffffffffu, pn, o), d2 = __shfl_xor_sync(0xffffffffu, pd, o);
            frac_max(pn, pd, n2, d2);
            const double n3 = __shfl_xor_sync(0xffffffffu, dn, o), d3 = __shfl_xor_sync(0xffffffffu, dd, o);
            frac_max(dn, dd, n3, d3);
        }
        ap_out = (pn > 0.0) ? fmin(1.0, tau * pd / pn) : 1.0;
        ad_out = (dn > 0.0) ? fmin(1.0, tau * dd / dn) : 1.0;
    }

    // ---------------------------------------------- multiplier steps (before the line search) ----
    __device__ void update_duals(double mu_t, double ad)
    {
        for (int k = lane; k < N; k += 32) {
            const int m = live(k);
            for (int j = 0; j < m; j++) {
                double r[4]; load_row(k, j, r);
                const double sj = S[k * SS + j], lj = LC[k * SS + j];
                const double rc = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + sj;
                const double ds = -rc - (r[0] * (double)DZ[k * NZ + 8] + r[1] * (double)DZ[k * NZ + 9] + r[2] * (double)DZ[k * NZ + 10]);
                LC[k * SS + j] = lj + ad * ((mu_t - lj * ds) * rcp_t(sj) - lj);
            }
        }
        for (int e = lane; e < N * NZ; e += 32) {
            if (!(e < 8 || e >= NZ)) continue;
            const int i = e % NZ;
            const double zi = Z[e], dzi = (double)DZ[e], zl = ZL[e], zu = ZU[e];
            const double isl = rcp_t(zi - BND[i]), isu = rcp_t(BND[NZ + i] - zi);
            ZL[e] = zl + ad * ((mu_t - zl * dzi) * isl - zl);
            ZU[e] = zu + ad * ((mu_t + zu * dzi) * isu - zu);
        }
    }

};

// =====================================================================================
// the kernel: grid = B CTAs of one warp; dynamic smem = MLayout::bytes(mcap)
// =====================================================================================
template <int N>
__global__ void __launch_bounds__(32) nmpc_ipm_mixed_kernel(const MixedParams prm)
{
    using ML = MLayout<N>;
    using C = Const<double>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= prm.B) return;
    const int b = prm.order ? prm.order[blockIdx.x] : (int)blockIdx.x;
    const int mcap = prm.mcap;
    const bool io32 = prm.io32 != 0;
    const size_t esz = io32 ? 4 : 8;

    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    int* nr = reinterpret_cast<int*>(smem_raw + 16);
    MixedSolver<N> s;
    s.bind(smem_raw, lane, mcap);
    s.io32 = io32;
    s.final_variant = (prm.variant == 1);
    const Opts& o = prm.o;

    // ---- stage the problem into shared memory with TMA bulk copies (SH is dead until the first evaluation) ----
    unsigned char* stg = reinterpret_cast<unsigned char*>(s.DZ);
    const uint32_t bytes_z = (uint32_t)(N * NZ * esz), bytes_h = (uint32_t)(N * 10 * esz), bytes_n = N * 4;
    s.rows_g = static_cast<const unsigned char*>(prm.rows) + (size_t)b * N * mcap * 4 * esz;
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, bytes_z + bytes_h + bytes_n);
        tma_load(stg, static_cast<const unsigned char*>(prm.z0) + (size_t)b * N * NZ * esz, bytes_z, bar);
        tma_load(stg + ML::STG_HDR_BYTES, static_cast<const unsigned char*>(prm.hdr) + (size_t)b * N * 10 * esz, bytes_h, bar);
        tma_load(nr, prm.nrows + (size_t)b * N, bytes_n, bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    for (int e = lane; e < N * NZ; e += 32)
        s.Z[e] = io32 ? (double)reinterpret_cast<const float*>(stg)[e] : reinterpret_cast<const double*>(stg)[e];
    for (int e = lane; e < N * 10; e += 32)
        s.HDR[(e / 10) * ML::HDR_S + (e % 10)] = io32 ? (double)reinterpret_cast<const float*>(stg + ML::STG_HDR_BYTES)[e]
                                                      : reinterpret_cast<const double*>(stg + ML::STG_HDR_BYTES)[e];
    __syncwarp();

    // ---- initial point -------------------------------------------------------------------
    if (lane < NZ) { s.BND[lane] = lower_bound<double>(lane); s.BND[NZ + lane] = upper_bound<double>(lane); }
    for (int e = lane; e < N * NZ; e += 32) s.DZ[e] = 0.0f;
    for (int e = lane; e < N * NXI; e += 32) s.DY[e] = 0.0f;
    int ncomp = 0;
    for (int k = lane; k < N; k += 32) {
#pragma unroll
        for (int i = 0; i < NZ; i++) {
            double v = s.Z[k * NZ + i];
            if (k == 0 && i >= 8)
                v = io32 ? (double)static_cast<const float*>(prm.xinit)[(size_t)b * 9 + i - 8] : static_cast<const double*>(prm.xinit)[(size_t)b * 9 + i - 8];
            if (is_free(k, i)) {
                const double lb = lower_bound<double>(i), ub = upper_bound<double>(i), kp = o.kappa_push;
                const double pl = fmin(kp * fmax(1.0, fabs(lb)), kp * (ub - lb));
                const double pu = fmin(kp * fmax(1.0, fabs(ub)), kp * (ub - lb));
                v = fmin(fmax(v, lb + pl), ub - pu);
                s.ZL[k * NZ + i] = o.mu0 / (v - lb);
                s.ZU[k * NZ + i] = o.mu0 / (ub - v);
                ncomp += 2;
            } else {
                s.ZL[k * NZ + i] = 0.0;
                s.ZU[k * NZ + i] = 0.0;
            }
            s.Z[k * NZ + i] = v;
        }
#pragma unroll
        for (int i = 0; i < NXI; i++) s.Y[k * NXI + i] = 0.0;
        const int m = s.live(k);
        for (int j = 0; j < m; j++) {
            double r[4]; s.load_row(k, j, r);
            double sl = (r[3] + C::hu) - (r[0] * s.Z[k * NZ + 8] + r[1] * s.Z[k * NZ + 9] + r[2] * s.Z[k * NZ + 10]);
            sl = fmax(sl, o.s_floor);
            s.S[k * s.SS + j] = sl;
            s.LC[k * s.SS + j] = o.mu0 / sl;
            ncomp++;
        }
    }
    ncomp = warp_sum(ncomp);
    __syncwarp();

    // ---- interior-point iterations ---------------------------------------------------------
    int flag = 0, it = 0, nbt_total = 0;
    double alpha_p = 0.0, alpha_d = 0.0, rs_n = 0.0, req_n = 0.0, rin_n = 0.0, rcomp = 0.0, mu = 0.0;
    double f_cur, th_cur, ls_cur;
    s.evaluate(0.0, f_cur, th_cur, ls_cur, req_n, rs_n);
    __syncwarp();
    const int it_cap = min(o.maxit, MIXED_BAIL_IT);
    // xinit outside a stage-0 bound or stage-0 corridor row (beyond TolIneq): the reference's NLP is infeasible -> NOPROGRESS (-7),
    // zero iterations, the violation in res_ineq (see nmpc_ipm.cuh)
    bool infeasible0;
    {
        double v0 = 0.0;
        if (lane >= 8 && lane < NZ) v0 = fmax(lower_bound<double>(lane) - s.Z[lane], s.Z[lane] - upper_bound<double>(lane));
        const int m0 = min(nr[0], mcap);
        for (int j = lane; j < m0; j += 32) {
            double r[4]; s.load_row(0, j, r);
            v0 = fmax(v0, r[0] * s.Z[8] + r[1] * s.Z[9] + r[2] * s.Z[10] - (r[3] + C::hu));
        }
        v0 = warp_max(v0);
        infeasible0 = v0 > o.tol_ineq;
        if (infeasible0) { flag = -7; rin_n = v0; }
    }
    double a_acc = 0.0;                       // primal step accepted by the last line search (0: the initial point)
    for (it = 0; !infeasible0; it++) {
        double csum, cmin;
        s.post_step(a_acc, rin_n, rcomp, csum, cmin);     // take the step, measure the new point, start the next system
        __syncwarp();
        mu = csum / (double)ncomp;
        const bool finite = isfinite(rs_n) && isfinite(req_n) && isfinite(mu) && isfinite(f_cur) && isfinite(th_cur);
        if (!finite) { flag = (it == 0) ? -6 : -7; break; }
        if (rs_n <= o.tol_stat && req_n <= o.tol_eq && rin_n <= o.tol_ineq && rcomp <= o.tol_comp) { flag = 1; break; }
        if (it >= it_cap) { flag = 0; break; }
        double sigma = o.sigma;
        if (sigma <= 0.0) {
            const double xi = cmin / mu;
            const double q = fmin(0.05 * (1.0 - xi) / xi, 2.0);
            sigma = 0.1 * q * q * q;
        }
        const double mu_t = fmax(sigma * mu, o.mu_floor);
        s.finish_rhs(mu_t);
        __syncwarp();
        bool ok;
        {
            double parked[ML::NPARK_LANE];
            s.park(parked);                 // the fp64 state leaves shared memory for the single-precision sweeps
            __syncwarp();
            ok = s.sw.riccati_backward();
            ok &= s.sw.rollout();
            s.sw.costates();
            __syncwarp();
            s.unpark(parked);
            __syncwarp();
        }
        if (!ok) { flag = -5; break; }
        const double tau = fmin(fmax(0.995, 1.0 - mu), 0.99999);
        double ap, ad;
        s.step_lengths(mu_t, tau, ap, ad);
        s.update_duals(mu_t, ad);       // independent of the primal step length; the trial evaluations see the new multipliers
        __syncwarp();
        const double ph0 = f_cur - mu_t * ls_cur;
        const double th_noise = fmax(10.0 * Eps<double>::v * double(N * NXI) * 20.0, 0.01 * o.tol_eq);
        double a = ap;
        int nbt = 0;
        double ft, tht, lst, reqt, rst;
        for (;;) {
            s.evaluate(a, ft, tht, lst, reqt, rst);
            __syncwarp();
            const double pht = ft - mu_t * lst;
            const bool acc = (tht <= fmax((1.0 - 1e-5) * th_cur, th_noise)) ||
                             (pht <= ph0 - 1e-5 * th_cur + 10.0 * Eps<double>::v * fabs(ph0));
            if (acc || nbt >= o.max_bt) break;
            nbt++;
            a *= 0.5;
        }
        nbt_total += nbt;
        alpha_p = a; alpha_d = ad;
        a_acc = a;                      // taken by post_step() at the top of the next pass
        f_cur = ft; th_cur = tht; ls_cur = lst; req_n = reqt; rs_n = rst;
        __syncwarp();
    }

    // ---- results -----------------------------------------------------------------------------
    __syncwarp();
    store_solution<32>(prm.peers, prm.z_out, prm.info_int, (size_t)b, s.Z, N * NZ, io32, lane, flag, it, nbt_total, 0);
    if (lane == 0) {
        const double v[8] = {req_n, rin_n, rs_n, rcomp, f_cur, mu, alpha_p, alpha_d};
#pragma unroll
        for (int q = 0; q < 8; q++) {
            if (io32) static_cast<float*>(prm.info_real)[(size_t)b * 8 + q] = (float)v[q];
            else static_cast<double*>(prm.info_real)[(size_t)b * 8 + q] = v[q];
        }
    }
    auto put = [&](void* base, size_t idx, double v) {
        if (io32) static_cast<float*>(base)[idx] = (float)v; else static_cast<double*>(base)[idx] = v;
    };
    if (prm.y_out)
        for (int e = lane; e < N * NXI; e += 32) put(prm.y_out, (size_t)b * N * NXI + e, (e < NXI) ? 0.0 : s.Y[e]);
    if (prm.zl_out)
        for (int e = lane; e < N * NZ; e += 32) put(prm.zl_out, (size_t)b * N * NZ + e, s.ZL[e]);
    if (prm.zu_out)
        for (int e = lane; e < N * NZ; e += 32) put(prm.zu_out, (size_t)b * N * NZ + e, s.ZU[e]);
    if (prm.lc_out)
        for (int e = lane; e < N * mcap; e += 32) {
            const int k = e / mcap, j = e - k * mcap;
            put(prm.lc_out, (size_t)b * N * mcap + e, (j < s.live(k)) ? s.LC[k * s.SS + j] : 0.0);
        }
}

}  // namespace nmpc
