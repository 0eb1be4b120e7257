// nmpc_ipm_group.cuh -- low-latency variant of the mixed-precision solve: one WARP-GROUP (8 warps = one CTA of 256
// threads) owns one MPC instance.
//
// Same drop-in role, algorithm, tolerances and exit codes as nmpc_ipm_mixed.cuh (FORCESNLPsolver_{normal,final}_solve,
// /root/reference/src/resilient_planner/plan_manage/solver/normal/FORCESNLPsolver_normal/include/FORCESNLPsolver_normal.h:321-323;
// fp64 iterate / model / residuals / line search, fp32 Newton system in delta form).  The one-warp kernels are built for
// THROUGHPUT: thousands of problems, every SM full, the latency of one problem's iteration (~57 us) hidden behind its
// neighbours.  A receding-horizon fleet of a few hundred vehicles (BASELINE config 5 at 128 agents per GPU) leaves the
// SMs nearly empty, and then the latency of ONE solve is all that counts.  Here a problem's iteration is spread over 256
// threads wherever the work allows:
//
//   * "flat" phases (bounds, multipliers, steps: 17 N independent (stage, variable) pairs) -- 2 rounds instead of 11;
//   * corridor-row phases -- four threads per stage, partial sums joined with two shuffles;
//   * the Riccati backward sweep (the dominant serial chain): every matrix entry of a stage's products is one work
//     item -- P+ F (169 entries + 13 for tv), F'(P+ F) (182), the Q blocks (81), the rank-4 update of the cost-to-go (104)
//     -- described by per-thread offset tables built once, so all threads run ONE instruction stream (no divergent
//     formula branches); five CTA barriers per stage replace the ~110-instruction serial chains of the one-warp sweep,
//     and the back-substitution that only the rollout needs (the gains) runs on warp 0 beside the rank-4 update;
//   * model evaluation: one stage per lane, the stage's work dealt to four warps by kind (dynamics + Jacobian + J'y |
//     objective | bound barriers | corridor rows), the stationarity residual assembled by all threads afterwards;
//   * forward rollout as a closed-loop recursion (transition matrices built by all threads, 13 multiply-adds per stage on
//     warp 0), costates from the stored cost-to-go in one parallel step.
//
// No register parking and no overlay: with few resident problems shared memory is not what limits anything.
// Two builds: MINB = 1 (255 registers, one CTA per SM) for fleets of at most one problem per SM, MINB = 2 (128 registers,
// some spills in the model evaluation) so that up to two problems per SM still run as ONE wave (296 problems: 0.53 ms
// instead of 0.83 ms in two waves, 0.79 ms for the one-warp kernel).
#pragma once
#include "nmpc_ipm_mixed.cuh"

namespace nmpc {

constexpr int GROUP_THREADS = 256;

template <int N> struct GLayout {
    using L32 = Layout<float, N, false>;
    using ML = MLayout<N>;
    static constexpr int HDR_S = L32::HDR_S, PHI_S = L32::PHI_S;
    static constexpr int HEAD_BYTES = 16 + N * 4;
    // ---- fp64 state (doubles) ----
    static constexpr int Z = 0;
    static constexpr int ZL = Z + N * NZ;
    static constexpr int ZU = ZL + N * NZ;
    static constexpr int Y = ZU + N * NZ;
    static constexpr int HDR = Y + N * NXI;
    static constexpr int BND = HDR + N * HDR_S;
    static constexpr int RED = BND + 2 * NZ;            // CTA-reduction scratch: 8 warps x 8 values
    static constexpr int JTY = RED + 64;                // evaluate(): J'y per (stage, variable), written by the dynamics warp
    static constexpr int GP = JTY + N * NZ;             //             cost gradient - E'y - z_l + z_u, written by the objective warp
    static constexpr int AL = GP + N * NZ;              //             A'lambda per stage (3), written by the corridor-row warp
    static constexpr int R_FIXED = AL + N * 3 + (N & 1);
    __host__ __device__ static constexpr int s_stride(int mcap) { return mcap | 1; }
    __host__ __device__ static constexpr int s_off(int) { return R_FIXED; }
    __host__ __device__ static constexpr int lc_off(int mcap) { return R_FIXED + N * s_stride(mcap); }
    __host__ __device__ static constexpr int r_end(int mcap) { return (R_FIXED + 2 * N * s_stride(mcap) + 1) & ~1; }
    // ---- fp32 region (floats), after the fp64 state: sweep-private arrays of Solver<float, N>, then the Newton data ----
    static constexpr int KG32 = L32::KFF + N * 4;
    static constexpr int O_END32 = (KG32 + N * 52 + 3) & ~3;
    static constexpr int SH_DZ = O_END32;
    static constexpr int SH_G = SH_DZ + N * NZ;
    static constexpr int SH_DY = SH_G + N * NZ;
    static constexpr int SH_D = SH_DY + N * NXI;
    static constexpr int SH_JC = SH_D + N * NXI;
    static constexpr int SH_PHID = SH_JC + N * NJC;
    static constexpr int SH_GF = SH_PHID + N * PHI_S;    // 13 x 14: F'(P+ F) and F' tv of the current stage
    static constexpr int SH_PK = SH_GF + 13 * 14 + 2;    // the cost-to-go matrices P_k of all stages (costates in one parallel step)
    static constexpr int SH_AK = SH_PK + N * 169;        // closed-loop transition [A_k | b_k], 13 x 14 per stage (rollout)
    static constexpr int SH_XI = SH_AK + N * 182;        // the rolled-out dxi_k, 13 per stage
    static constexpr int SH_END = SH_XI + N * NXI + 3;
    static constexpr int STG_HDR_BYTES = N * NZ * 8;     // TMA staging (z0, then hdr) at SH_DZ
    __host__ __device__ static constexpr size_t bytes(int mcap)
    {
        return (size_t)HEAD_BYTES + (size_t)r_end(mcap) * 8 + (size_t)SH_END * 4;
    }
};

// one term of a "gather-sum": value += mult * f32[off + k * kstride]
struct GTerm { int off, kstride; float mult; };

template <int N> struct GroupSolver {
    using GL = GLayout<N>;
    using L32 = typename GL::L32;
    using C = Const<double>;
    static constexpr int NT = GROUP_THREADS, NW = NT / 32, RP = 4;      // RP threads share one stage's corridor rows
    double* r64;
    float* f32;      // base of the fp32 region
    int* nr;
    int tid, lane, warp, mcap, SS;
    const void* rows_g;
    bool io32, final_variant;
    double *Z, *ZL, *ZU, *Y, *HDR, *BND, *S, *LC, *RED;
    float *DZ, *G, *DY, *D, *JC, *PHID, *GF;
    Solver<float, N, false> sw;

    __device__ __forceinline__ void bind(unsigned char* smem_raw, int tid_, int mcap_)
    {
        r64 = reinterpret_cast<double*>(smem_raw + GL::HEAD_BYTES);
        nr = reinterpret_cast<int*>(smem_raw + 16);
        tid = tid_; lane = tid_ & 31; warp = tid_ >> 5; mcap = mcap_; SS = GL::s_stride(mcap);
        Z = r64 + GL::Z; ZL = r64 + GL::ZL; ZU = r64 + GL::ZU; Y = r64 + GL::Y; HDR = r64 + GL::HDR; BND = r64 + GL::BND; RED = r64 + GL::RED;
        S = r64 + GL::s_off(mcap); LC = r64 + GL::lc_off(mcap);
        f32 = reinterpret_cast<float*>(r64 + GL::r_end(mcap));
        DZ = f32 + GL::SH_DZ; G = f32 + GL::SH_G; DY = f32 + GL::SH_DY; D = f32 + GL::SH_D; JC = f32 + GL::SH_JC;
        PHID = f32 + GL::SH_PHID; GF = f32 + GL::SH_GF;
        sw.sm = f32; sw.nr = nr; sw.lane = lane; sw.mcap = mcap; sw.SS = SS; sw.rows_g = nullptr; sw.final_variant = false;
        sw.DZ = DZ; sw.G = G; sw.P = DY; sw.D = D; sw.JC = JC; sw.PHID = PHID;
        sw.KG = f32 + GL::KG32; sw.KFF = f32 + L32::KFF;
        sw.Z = sw.ZL = sw.ZU = sw.Y = sw.HDR = sw.S = sw.LC = sw.BND = nullptr;
        sw.QINV = sw.PQQ0 = sw.DZAP = nullptr;
        sw.fac_out = nullptr;
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

    // ---- CTA reductions: warp shuffles, one shared-memory hop, every thread combines the four partials in the same order ----
    // ops: 0 = sum, 1 = max, 2 = min.  v[] is replaced by the CTA-wide result on every thread.
    template <int NV> __device__ void reduce(double (&v)[NV], const int (&ops)[NV])
    {
#pragma unroll
        for (int q = 0; q < NV; q++) v[q] = ops[q] == 0 ? warp_sum(v[q]) : (ops[q] == 1 ? warp_max(v[q]) : warp_min(v[q]));
        __syncthreads();                               // RED is free again
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < NV; q++) RED[warp * 8 + q] = v[q];
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < NV; q++) {
            double r = RED[q];
#pragma unroll
            for (int w = 1; w < NW; w++) {
                const double x = RED[w * 8 + q];
                r = ops[q] == 0 ? r + x : (ops[q] == 1 ? fmax(r, x) : fmin(r, x));
            }
            v[q] = r;
        }
    }

    // ---------------------------------------------------------------- model evaluation, four warps ---
    // One stage per lane as in the one-warp kernels, but the stage's work is dealt to four warps by kind, so that the
    // longest instruction stream -- the Heun step with its Jacobian and J'y -- no longer carries the rest behind it:
    //   warp 0  dynamics, Jacobian (kept in registers for J'y in double precision, then rounded into the Newton data),
    //           defects, J'y                                   -> JTY
    //   warp 1  objective and its gradient, - E'y - z_l + z_u  -> GP
    //   warp 2  log-barrier terms of the bounds
    //   warp 3  corridor rows: A'lambda, their share of theta and of the log-barrier   -> AL
    // then all threads: stationarity residual r = GP + JTY + A'lambda per (stage, variable), its inf-norm, rounded into G.
    __device__ void evaluate(double a, double& f_out, double& th_out, double& ls_out, double& req_out, double& rs_out)
    {
        double f = 0.0, th = 0.0, ls = 0.0, rq = 0.0, rs = 0.0;
        double* JTY = r64 + GL::JTY;
        double* GP = r64 + GL::GP;
        double* AL = r64 + GL::AL;
        if (warp < 4) {
            for (int k = lane; k < N; k += 32) {
                const double* hdr = HDR + k * GL::HDR_S;
                if (warp == 0) {
                    if (k < N - 1) {
                        double zk[NZ], c[NXI], jc[NJC], yn[NXI];
#pragma unroll
                        for (int i = 0; i < NZ; i++) zk[i] = Z[k * NZ + i] + a * (double)DZ[k * NZ + i];
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
                        for (int i = 0; i < NZ; i++) JTY[k * NZ + i] = jt_y<double>(jc, yn, i);
                    } else {
#pragma unroll
                        for (int i = 0; i < NZ; i++) JTY[k * NZ + i] = 0.0;
                    }
                } else if (warp == 1) {
                    double zk[NZ], g[NZ];
#pragma unroll
                    for (int i = 0; i < NZ; i++) zk[i] = Z[k * NZ + i] + a * (double)DZ[k * NZ + i];
                    f += objective<double, true>(zk, hdr, k == 0, final_variant && k == N - 1, g);
                    if (k > 0) {
#pragma unroll
                        for (int i = 0; i < NXI; i++) g[i < 9 ? 8 + i : i - 5] -= Y[k * NXI + i] + a * (double)DY[k * NXI + i];
                    }
#pragma unroll
                    for (int i = 0; i < NZ; i++) GP[k * NZ + i] = g[i] - ZL[k * NZ + i] + ZU[k * NZ + i];
                } else if (warp == 2) {
                    double prod = 1.0;
#pragma unroll
                    for (int i = 0; i < NZ; i++) {
                        const double zi = Z[k * NZ + i] + a * (double)DZ[k * NZ + i];
                        double sl = zi - lower_bound<double>(i), su = upper_bound<double>(i) - zi;
                        if (i >= 8 && k == 0) { sl = 1.0; su = 1.0; }
                        prod *= sl * su;
                        if (i % 9 == 8 || i == NZ - 1) { ls += log(prod); prod = 1.0; }
                    }
                } else {
                    const int m = live(k);
                    double al0 = 0.0, al1 = 0.0, al2 = 0.0, prod = 1.0;
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
                    AL[k * 3] = al0; AL[k * 3 + 1] = al1; AL[k * 3 + 2] = al2;
                }
            }
        }
        __syncthreads();
        for (int e = tid; e < N * NZ; e += NT) {
            const int k = e / NZ, i = e - k * NZ;
            double r = GP[e] + JTY[e];
            if (i >= 8 && i < 11) r += AL[k * 3 + i - 8];
            if (i >= 8 && k == 0) r = 0.0;
            rs = fmax(rs, fabs(r));
            G[e] = (float)r;
        }
        double v[5] = {f, th, ls, rq, rs};
        const int ops[5] = {0, 0, 0, 1, 1};
        reduce<5>(v, ops);
        f_out = v[0]; th_out = v[1]; ls_out = v[2]; req_out = v[3]; rs_out = v[4];
    }

    // per-thread slice of the corridor rows: stage ks = tid / RP (+ NT / RP per round), rows j = tid % RP, + RP, ...
    template <typename F> __device__ __forceinline__ void for_rows(F&& body) const
    {
        for (int k = tid / RP; k < N; k += NT / RP) {
            const int m = live(k);
            for (int j = tid % RP; j < m; j += RP) {
                double r[4]; load_row(k, j, r);
                const double sj = S[k * SS + j], lj = LC[k * SS + j];
                const double rc = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + sj;
                body(k, j, r, sj, lj, rc);
            }
        }
    }

    __device__ void residuals(double& rin_n, double& rcomp, double& csum, double& cmin)
    {
        double rin = 0.0, cmx = 0.0, cs = 0.0, cmn = 1e30;
        for_rows([&](int, int, const double (&)[4], double sj, double lj, double rc) {
            const double cc = sj * lj;
            cs += cc; cmx = fmax(cmx, cc); cmn = fmin(cmn, cc);
            rin = fmax(rin, fmax(fabs(rc), rc - sj));
        });
        for (int e = tid; e < N * NZ; e += NT) {
            if (!(e < 8 || e >= NZ)) continue;
            const int i = e % NZ;
            const double zi = Z[e];
            const double cl = (zi - BND[i]) * ZL[e], cu = (BND[NZ + i] - zi) * ZU[e];
            cs += cl + cu;
            cmx = fmax(cmx, fmax(cl, cu));
            cmn = fmin(cmn, fmin(cl, cu));
        }
        double v[4] = {rin, cmx, cs, cmn};
        const int ops[4] = {1, 1, 0, 2};
        reduce<4>(v, ops);
        rin_n = v[0]; rcomp = v[1]; csum = v[2]; cmin = v[3];
    }

    __device__ void assemble(double mu_t)
    {
        for (int e = tid; e < N * NZ; e += NT) {
            const int k = e / NZ, i = e - k * NZ;
            float* phi = PHID + k * GL::PHI_S;
            if (e < 8 || e >= NZ) {
                const double zi = Z[e], zl = ZL[e], zu = ZU[e];
                const double isl = rcp_t(zi - BND[i]), isu = rcp_t(BND[NZ + i] - zi);
                phi[i] = (float)(cost_hess_diag<double>(i, HDR + k * GL::HDR_S, k == 0, final_variant && k == N - 1) + zl * isl + zu * isu);
                G[e] = (float)((double)G[e] + ((zl - mu_t * isl) - (zu - mu_t * isu)));
            } else {
                phi[i] = 1.0f;
                G[e] = 0.0f;
            }
        }
        __syncthreads();
        for (int k0 = 0; k0 < N; k0 += NT / RP) {            // uniform trip count: the shuffles below need every lane
            const int k = k0 + tid / RP;
            double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};      // d0 d1 d2 o01 o02 o12 g0 g1 g2
            if (k < N) {
                const int m = live(k);
                for (int j = tid % RP; j < m; j += RP) {
                    double r[4]; load_row(k, j, r);
                    const double sj = S[k * SS + j], lj = LC[k * SS + j], is = rcp_t(sj);
                    const double rc = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + sj;
                    const double sg = lj * is, tt = (mu_t + lj * (rc - sj)) * is;
                    acc[0] += r[0] * r[0] * sg; acc[1] += r[1] * r[1] * sg; acc[2] += r[2] * r[2] * sg;
                    acc[3] += r[0] * r[1] * sg; acc[4] += r[0] * r[2] * sg; acc[5] += r[1] * r[2] * sg;
                    acc[6] += r[0] * tt; acc[7] += r[1] * tt; acc[8] += r[2] * tt;
                }
            }
#pragma unroll
            for (int q = 0; q < 9; q++) {
                acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 1);
                acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 2);
            }
            if (k < N && tid % RP == 0) {
                float* phi = PHID + k * GL::PHI_S;
                phi[8] = (float)((double)phi[8] + acc[0]); phi[9] = (float)((double)phi[9] + acc[1]); phi[10] = (float)((double)phi[10] + acc[2]);
                phi[17] = (float)acc[3]; phi[18] = (float)acc[4]; phi[19] = (float)acc[5];
                phi[20] = (float)(-2.0 * HDR[k * GL::HDR_S + 8]);
                if (k > 0) {
                    G[k * NZ + 8] = (float)((double)G[k * NZ + 8] + acc[6]);
                    G[k * NZ + 9] = (float)((double)G[k * NZ + 9] + acc[7]);
                    G[k * NZ + 10] = (float)((double)G[k * NZ + 10] + acc[8]);
                }
            }
        }
    }

    __device__ void step_lengths(double mu_t, double tau, double& ap_out, double& ad_out)
    {
        double pn = 0.0, pd = 1.0, dn = 0.0, dd = 1.0;
        for (int e = tid; e < N * NZ; e += NT) {
            if (!(e < 8 || e >= NZ)) continue;
            const int i = e % NZ;
            const double zi = Z[e], dzi = (double)DZ[e], zl = ZL[e], zu = ZU[e];
            const double sl = zi - BND[i], su = BND[NZ + i] - zi;
            frac_max(pn, pd, -dzi, sl);
            frac_max(pn, pd, dzi, su);
            frac_max(dn, dd, zl * (sl + dzi) - mu_t, sl * zl);
            frac_max(dn, dd, zu * (su - dzi) - mu_t, su * zu);
        }
        for_rows([&](int k, int, const double (&r)[4], double sj, double lj, double rc) {
            const double ds = -rc - (r[0] * (double)DZ[k * NZ + 8] + r[1] * (double)DZ[k * NZ + 9] + r[2] * (double)DZ[k * NZ + 10]);
            frac_max(pn, pd, -ds, sj);
            frac_max(dn, dd, lj * (sj + ds) - mu_t, sj * lj);
        });
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double n2 = __shfl_xor_sync(0xffffffffu, pn, o), d2 = __shfl_xor_sync(0xffffffffu, pd, o);
            frac_max(pn, pd, n2, d2);
            const double n3 = __shfl_xor_sync(0xffffffffu, dn, o), d3 = __shfl_xor_sync(0xffffffffu, dd, o);
            frac_max(dn, dd, n3, d3);
        }
        __syncthreads();
        if (lane == 0) { RED[warp * 8] = pn; RED[warp * 8 + 1] = pd; RED[warp * 8 + 2] = dn; RED[warp * 8 + 3] = dd; }
        __syncthreads();
        pn = RED[0]; pd = RED[1]; dn = RED[2]; dd = RED[3];
#pragma unroll
        for (int w = 1; w < NW; w++) {
            frac_max(pn, pd, RED[w * 8], RED[w * 8 + 1]);
            frac_max(dn, dd, RED[w * 8 + 2], RED[w * 8 + 3]);
        }
        ap_out = (pn > 0.0) ? fmin(1.0, tau * pd / pn) : 1.0;
        ad_out = (dn > 0.0) ? fmin(1.0, tau * dd / dn) : 1.0;
    }

    __device__ void update_duals(double mu_t, double ad)
    {
        for_rows([&](int k, int j, const double (&r)[4], double sj, double lj, double rc) {
            const double ds = -rc - (r[0] * (double)DZ[k * NZ + 8] + r[1] * (double)DZ[k * NZ + 9] + r[2] * (double)DZ[k * NZ + 10]);
            LC[k * SS + j] = lj + ad * ((mu_t - lj * ds) * rcp_t(sj) - lj);
        });
        for (int e = tid; e < N * NZ; e += NT) {
            if (!(e < 8 || e >= NZ)) continue;
            const int i = e % NZ;
            const double zi = Z[e], dzi = (double)DZ[e], zl = ZL[e], zu = ZU[e];
            const double isl = rcp_t(zi - BND[i]), isu = rcp_t(BND[NZ + i] - zi);
            ZL[e] = zl + ad * ((mu_t - zl * dzi) * isl - zl);
            ZU[e] = zu + ad * ((mu_t + zu * dzi) * isu - zu);
        }
    }

    __device__ void update_primal(double a)
    {
        for_rows([&](int k, int j, const double (&r)[4], double sj, double, double rc) {
            const double ds = -rc - (r[0] * (double)DZ[k * NZ + 8] + r[1] * (double)DZ[k * NZ + 9] + r[2] * (double)DZ[k * NZ + 10]);
            S[k * SS + j] = sj + a * ds;
        });
        __syncthreads();
        for (int e = tid; e < N * NZ; e += NT) Z[e] += a * (double)DZ[e];
        for (int e = NXI + tid; e < N * NXI; e += NT) Y[e] += a * (double)DY[e];
    }

    // ------------------------------------------------------------- Riccati backward, 256 threads ----
    // Work items and their offset tables (offsets relative to f32, the base of the fp32 region; Jacobian words relative
    // to the stage's 51-word block).
    //   structured product  out = mA (j[a0] x0 + j[a1] x1 + j[a2] x2) + mB (j[b0] x3 + j[b1] x4 + j[b2] x5) + cc x[ic]
    //   with x = (xp, xv, xa): a row of P+ (phase A) or a column of P+ F / the vector tv (phase B); the output type
    //   (v-ordering w0 w1 w2 T p0 p1 p2 v0 v1 v2 r0 r1 r2) fixes the Jacobian words -- same formulas as Solver::ft_times.
    struct SP { int ja[3], jb[3], x[6], xc, dst; float mA, mB, cc; bool on; };
    __device__ static SP sp_desc(int o, int xoff, int xs, int dst, bool on)
    {
        int oa = 0, sa = 0, ob = 0, sb = 0, ic = 0;
        SP d;
        d.mA = 0.f; d.mB = 0.f; d.cc = 0.f; d.on = on; d.dst = dst;
        if (o < 3) { d.mB = 1.f; ob = JVW + o; sb = 3; d.cc = (float)C::h; ic = 6 + o; }
        else if (o == 3) { d.mA = 1.f; oa = JPT; sa = 1; d.mB = 1.f; ob = JVT; sb = 1; }
        else if (o < 7) { d.cc = 1.f; ic = o - 4; }
        else if (o < 10) { d.mA = 1.f; oa = JPV + o - 7; sa = 3; d.mB = 1.f; ob = JVV + o - 7; sb = 3; }
        else { d.mA = 1.f; oa = JPR + o - 10; sa = 3; d.mB = 1.f; ob = JVR + o - 10; sb = 3; d.cc = 1.f; ic = 6 + o - 10; }
#pragma unroll
        for (int q = 0; q < 3; q++) { d.ja[q] = oa + q * sa; d.jb[q] = ob + q * sb; }
#pragma unroll
        for (int q = 0; q < 6; q++) d.x[q] = xoff + q * xs;
        d.xc = xoff + ic * xs;
        return d;
    }
    __device__ __forceinline__ float sp_eval(const SP& d, const float* __restrict__ jc) const
    {
        const float a = (jc[d.ja[0]] * f32[d.x[0]] + jc[d.ja[1]] * f32[d.x[1]]) + jc[d.ja[2]] * f32[d.x[2]];
        const float b = (jc[d.jb[0]] * f32[d.x[3]] + jc[d.jb[1]] * f32[d.x[4]]) + jc[d.jb[2]] * f32[d.x[5]];
        return (d.mA * a + d.mB * b) + d.cc * f32[d.xc];
    }

    // ------------------------------------------------- forward rollout and costates, 256 threads ----
    // The one-warp kernels roll the step out stage by stage (du_k = K_k dxi_k + kff_k, dxi_k+1 = d_k + M_k dxi_k + B_k du_k:
    // a gain product, shuffles and a Jacobian product on the dependent chain, ~500 cycles per stage) and recover the
    // costates by a second dependent sweep.  Here everything that does not depend on the rolled-out state is taken off the
    // chain:
    //   1. all threads, one (stage, row) each: the closed-loop transition A_k = M_k + B_k K_k, b_k = d_k + B_k kff_k
    //      (M_k, B_k: the structured Jacobian blocks); warp 0 meanwhile solves stage 0 (dq_0 = -P_qq^-1 p_q);
    //   2. warp 0: dxi_k+1 = b_k + A_k dxi_k, lane = row, the state passed by shuffles (13 multiply-adds per stage);
    //   3. all threads: du_k = K_k dxi_k + kff_k and the step dz_k; costates y_k = P_k dxi_k + p_k from the stored
    //      cost-to-go (p_k sits where y_k goes) -- both read only dxi, no barrier between them.
    // Returns false (on warp 0) if the stage-0 block is not positive definite.
    __device__ bool rollout_and_costates()
    {
        float* AK = f32 + GL::SH_AK;
        float* XI = f32 + GL::SH_XI;
        bool ok = true;
        if (warp == 0) {                                             // stage 0: x fixed (dx = 0), u_prev free
            const float* PN = f32 + L32::PN;
            float a[16], l[10], li[4], x[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
#pragma unroll
                for (int c = 0; c <= r; c++) a[4 * r + c] = PN[(9 + r) * 13 + 9 + c];
                x[r] = -DY[9 + r];
            }
            ok = chol4<float>(a, l, li);
            fsub4<float>(l, li, x);
            bsub4<float>(l, li, x);
            if (lane < 13) XI[lane] = (lane < 9) ? 0.f : (lane == 9 ? x[0] : (lane == 10 ? x[1] : (lane == 11 ? x[2] : x[3])));
        }
        for (int t = tid; t < (N - 1) * NXI; t += NT) {
            const int k = t / NXI, i = t - k * NXI;
            const int rt = i < 3 ? 0 : (i < 6 ? 1 : (i < 9 ? 2 : 3));
            const int r = i - (rt == 0 ? 0 : (rt == 1 ? 3 : (rt == 2 ? 6 : 9)));
            const float* jc = JC + k * NJC;
            const float* kg = sw.KG + k * 52;
            const float* kff = sw.KFF + k * 4;
            float b4[4];
#pragma unroll
            for (int c = 0; c < 3; c++) b4[c] = rt == 1 ? jc[JVW + 3 * r + c] : (rt >= 2 && c == r ? (rt == 2 ? (float)C::h : 1.f) : 0.f);
            b4[3] = rt == 0 ? jc[JPT + r] : (rt == 1 ? jc[JVT + r] : (rt == 3 && r == 3 ? 1.f : 0.f));
            const int offV = rt == 0 ? JPV + 3 * r : JVV + 3 * r, offR = rt == 0 ? JPR + 3 * r : JVR + 3 * r;
            float* arow = AK + k * 182 + i * 14;
#pragma unroll
            for (int j = 0; j < NXI; j++) {
                float m = 0.f;
                if (j >= 3 && j < 6) m = rt <= 1 ? jc[offV + j - 3] : 0.f;
                if (j >= 6 && j < 9) m = rt <= 1 ? jc[offR + j - 6] : 0.f;
                if ((rt == 0 || rt == 2) && j == i) m += 1.f;
                arow[j] = m + ((b4[0] * kg[j] + b4[1] * kg[13 + j]) + (b4[2] * kg[26 + j] + b4[3] * kg[39 + j]));
            }
            arow[13] = D[t] + ((b4[0] * kff[0] + b4[1] * kff[1]) + (b4[2] * kff[2] + b4[3] * kff[3]));
        }
        __syncthreads();
        if (warp == 0) {
            const int row = lane < 13 ? lane : 0;
            float xi = XI[row];
            for (int k = 0; k < N - 1; k++) {
                const float* arow = AK + k * 182 + row * 14;
                float c0 = arow[13], c1 = 0.f, c2 = 0.f;
#pragma unroll
                for (int j = 0; j < 12; j += 3) {
                    c0 += arow[j] * __shfl_sync(0xffffffffu, xi, j);
                    c1 += arow[j + 1] * __shfl_sync(0xffffffffu, xi, j + 1);
                    c2 += arow[j + 2] * __shfl_sync(0xffffffffu, xi, j + 2);
                }
                c0 += arow[12] * __shfl_sync(0xffffffffu, xi, 12);
                xi = (c0 + c1) + c2;
                if (lane < 13) XI[(k + 1) * NXI + lane] = xi;
            }
        }
        __syncthreads();
        for (int t = tid; t < N * NZ; t += NT) {
            const int k = t / NZ, i = t - k * NZ;
            const float* xk = XI + k * NXI;
            float v;
            if (i < 4) {
                const float* kg = sw.KG + k * 52 + i * 13;
                float c0 = sw.KFF[k * 4 + i], c1 = 0.f, c2 = 0.f;
#pragma unroll
                for (int j = 0; j < 12; j += 3) { c0 += kg[j] * xk[j]; c1 += kg[j + 1] * xk[j + 1]; c2 += kg[j + 2] * xk[j + 2]; }
                v = (c0 + c1) + (c2 + kg[12] * xk[12]);
            } else {
                v = xk[i < 8 ? 5 + i : i - 8];
            }
            DZ[t] = v;
        }
        for (int t = NXI + tid; t < N * NXI; t += NT) {
            const int k = t / NXI, i = t - k * NXI;
            const float* pk = f32 + GL::SH_PK + k * 169 + i * 13;
            const float* xk = XI + k * NXI;
            float c0 = DY[t], c1 = 0.f, c2 = 0.f;
#pragma unroll
            for (int j = 0; j < 12; j += 3) { c0 += pk[j] * xk[j]; c1 += pk[j + 1] * xk[j + 1]; c2 += pk[j + 2] * xk[j + 2]; }
            DY[t] = (c0 + c1) + (c2 + pk[12] * xk[12]);
        }
        return ok;
    }

    __device__ bool riccati_backward()
    {
        constexpr int PN = L32::PN, PF = L32::PF, TV = L32::TV, QUU = L32::QUU, QUR = L32::QUR, QV = L32::QV, QXI = L32::QXI,
                      YS = L32::YS, Y0 = L32::Y0, oGF = GL::SH_GF, oG = GL::SH_G, oPHI = GL::SH_PHID, oDY = GL::SH_DY, oD = GL::SH_D;
        static_assert(NT >= 32 + 182 && NT >= 32 + 104, "work items of a phase must fit one round");
        // phase A on warps 1.. (warp 0 may still be finishing the previous stage's back-substitution):
        //   w < 169 -> PF[r][o] from row r of P+;  169 <= w < 182 -> tv[w - 169]
        const int wA = tid - 32;
        const bool onA = wA >= 0 && wA < 169;
        const SP pa = sp_desc(onA ? wA % 13 : 0, PN + (onA ? wA / 13 : 0) * 13, 1, PF + (onA ? wA : 0), onA);
        const int tvrow = (wA >= 169 && wA < 182) ? wA - 169 : -1;
        // phase B on all warps: w < 182 -> GF[o][j] from column j of P+ F (j < 13) or from tv (j = 13)
        const bool onB = tid < 182;
        const int jB = onB ? tid / 13 : 0, oB = onB ? tid % 13 : 0;
        const SP pb = sp_desc(oB, jB == 13 ? TV : PF + jB, jB == 13 ? 1 : 13, oGF + oB * 14 + jB, onB);
        // phase B2: the Q blocks as gather-sums of at most five terms (81 work items)
        GTerm b2[5];
        int b2dst = -1;
        {
#pragma unroll
            for (int q = 0; q < 5; q++) b2[q] = GTerm{PN, 0, 0.f};
            const int id = tid;
            if (id < 16) {                                    // Q_uu[i][j]
                const int i = id >> 2, j = id & 3;
                b2[0] = GTerm{oGF + i * 14 + j, 0, 1.f};
                b2[1] = GTerm{PN + (9 + i) * 13 + 9 + j, 0, 1.f};
                b2[2] = GTerm{PF + (9 + i) * 13 + j, 0, 1.f};
                b2[3] = GTerm{PF + (9 + j) * 13 + i, 0, 1.f};
                if (i == j) b2[4] = GTerm{oPHI + i, GL::PHI_S, 1.f};
                b2dst = QUU + i * 4 + j;
            } else if (id < 52) {                             // Q_ux[j][i - 4]
                const int e = id - 16, j = e / 9, i = 4 + e % 9;
                b2[0] = GTerm{oGF + i * 14 + j, 0, 1.f};
                b2[1] = GTerm{PF + (9 + j) * 13 + i, 0, 1.f};
                b2dst = QUR + j * 13 + i - 4;
            } else if (id < 68) {                             // Q_uq = -2 w_rate I
                const int e = id - 52, r = e >> 2, c = e & 3;
                if (r == c) b2[0] = GTerm{oPHI + 20, GL::PHI_S, 1.f};
                b2dst = QUR + r * 13 + 9 + c;
            } else if (id < 72) {                             // q~_u
                const int w = id - 68;
                b2[0] = GTerm{oGF + w * 14 + 13, 0, 1.f};
                b2[1] = GTerm{oG + w, NZ, 1.f};
                b2[2] = GTerm{TV + 9 + w, 0, 1.f};
                b2dst = QV + w;
            } else if (id < 81) {                             // q~_x
                const int w = 4 + id - 72;
                b2[0] = GTerm{oGF + w * 14 + 13, 0, 1.f};
                b2[1] = GTerm{oG + w + 4, NZ, 1.f};
                b2dst = QXI + w - 4;
            }
        }
        // phase D on warps 1..: 91 entries of P_k (i >= j) + 13 of p_k
        const int wD = tid - 32;
        int di = 0, dj = 0, dgf = -1, dphi = -1;
        const bool dmat = wD >= 0 && wD < 91, dvec = wD >= 91 && wD < 104;
        if (dmat) {
            int i = (int)((sqrtf(8.0f * (float)wD + 1.0f) - 1.0f) * 0.5f);
            i += ((i + 1) * (i + 2) / 2 <= wD) - (i * (i + 1) / 2 > wD);
            const int j = wD - i * (i + 1) / 2;
            di = i; dj = j;
            dgf = (i < 9) ? oGF + (4 + i) * 14 + 4 + j : -1;
            dphi = (i < 9) ? (i == j ? 8 + i : (i < 3 ? 17 + i + j - 1 : -1)) : (i == j ? 4 + i - 9 : -1);
        } else if (dvec) {
            di = wD - 91;
        }
        // the terminal stage is the generic stage with P+ = 0: no products, the Q blocks come out of zeros
        for (int e = tid; e < 169; e += NT) { f32[PN + e] = 0.f; f32[PF + e] = 0.f; }
        for (int e = tid; e < 13 * 14; e += NT) f32[oGF + e] = 0.f;
        if (tid < 13) f32[TV + tid] = 0.f;
        bool ok = true;
        __syncthreads();
        for (int k = N - 1; k >= 0; k--) {
            if (k < N - 1) {
                const float* jc = JC + k * NJC;
                // ---- phase A ----
                if (pa.on) f32[pa.dst] = sp_eval(pa, jc);
                if (tvrow >= 0) {
                    const float* pr = f32 + PN + tvrow * 13;
                    const float* dk = f32 + oD + k * NXI;
                    float c0 = f32[oDY + (k + 1) * NXI + tvrow], c1 = 0.f, c2 = 0.f;
#pragma unroll
                    for (int q = 0; q < 13; q += 3) {
                        c0 += pr[q] * dk[q];
                        if (q + 1 < 13) c1 += pr[q + 1] * dk[q + 1];
                        if (q + 2 < 13) c2 += pr[q + 2] * dk[q + 2];
                    }
                    f32[TV + tvrow] = (c0 + c1) + c2;
                }
                __syncthreads();
                // ---- phase B ----
                if (pb.on) f32[pb.dst] = sp_eval(pb, jc);
                __syncthreads();
            }
            // ---- phase B2: Q blocks ----
            if (b2dst >= 0) {
                float v = 0.f;
#pragma unroll
                for (int q = 0; q < 5; q++) v += b2[q].mult * f32[b2[q].off + k * b2[q].kstride];
                f32[b2dst] = v;
            }
            __syncthreads();
            // ---- phase C1 (warp 0): pivot block factorised redundantly in registers, forward substitution of 13 + 1 columns ----
            float l[10], li[4], x[4];
            if (warp == 0) {
                float a[16];
                a[0] = f32[QUU + 0]; a[4] = f32[QUU + 4]; a[5] = f32[QUU + 5]; a[8] = f32[QUU + 8]; a[9] = f32[QUU + 9]; a[10] = f32[QUU + 10];
                a[12] = f32[QUU + 12]; a[13] = f32[QUU + 13]; a[14] = f32[QUU + 14]; a[15] = f32[QUU + 15];
                ok &= chol4<float>(a, l, li);
                if (lane < 14) {
#pragma unroll
                    for (int r = 0; r < 4; r++) x[r] = (lane < 13) ? f32[QUR + r * 13 + lane] : f32[QV + r];
                    fsub4<float>(l, li, x);
                    float* ys = (lane < 13) ? f32 + YS + lane : f32 + Y0;
                    const int ystr = (lane < 13) ? 13 : 1;
#pragma unroll
                    for (int r = 0; r < 4; r++) ys[r * ystr] = x[r];
                }
            }
            __syncthreads();
            // ---- phase C2 (warp 0: back-substitution -> gains, needed only by the rollout)  ||  phase D (warps 1..):
            //      P_k = blkdiag(Q_xx, Phi_qq) - Y'Y,  p_k = q_xi - Y' y0 ----
            if (warp == 0) {
                if (lane < 14) {
                    bsub4<float>(l, li, x);
                    float* kg = (lane < 13) ? sw.KG + k * 52 + lane : sw.KFF + k * 4;
                    const int ystr = (lane < 13) ? 13 : 1;
#pragma unroll
                    for (int r = 0; r < 4; r++) kg[r * ystr] = -x[r];
                }
            } else if (dmat) {
                float v = (dgf >= 0) ? f32[dgf] : 0.f;
                if (dphi >= 0) v += f32[oPHI + k * GL::PHI_S + dphi];
                const float* ys = f32 + YS;
                v -= (ys[di] * ys[dj] + ys[13 + di] * ys[13 + dj]) + (ys[26 + di] * ys[26 + dj] + ys[39 + di] * ys[39 + dj]);
                f32[PN + di * 13 + dj] = v;
                f32[PN + dj * 13 + di] = v;
                f32[GL::SH_PK + k * 169 + di * 13 + dj] = v;
                f32[GL::SH_PK + k * 169 + dj * 13 + di] = v;
            } else if (dvec) {
                const float* ys = f32 + YS;
                const float* y0 = f32 + Y0;
                const float qxi = (di < 9) ? f32[QXI + di] : f32[oG + k * NZ + di - 5];
                f32[oDY + k * NXI + di] = qxi - ((ys[di] * y0[0] + ys[13 + di] * y0[1]) + (ys[26 + di] * y0[2] + ys[39 + di] * y0[3]));
            }
            __syncthreads();
        }
        return ok;
    }
};

// =====================================================================================
// the kernel: grid = B CTAs of 256 threads; dynamic smem = GLayout::bytes(mcap)
// =====================================================================================
template <int N, int MINB = 1>
__global__ void __launch_bounds__(GROUP_THREADS, MINB) nmpc_ipm_group_kernel(const MixedParams prm)
{
    using GL = GLayout<N>;
    using C = Const<double>;
    constexpr int NT = GROUP_THREADS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    if ((int)blockIdx.x >= prm.B) return;
    const int b = prm.order ? prm.order[blockIdx.x] : (int)blockIdx.x;
    const int mcap = prm.mcap;
    const bool io32 = prm.io32 != 0;
    const size_t esz = io32 ? 4 : 8;

    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    int* nr = reinterpret_cast<int*>(smem_raw + 16);
    GroupSolver<N> s;
    s.bind(smem_raw, tid, mcap);
    s.io32 = io32;
    s.final_variant = (prm.variant == 1);
    const Opts& o = prm.o;

    unsigned char* stg = reinterpret_cast<unsigned char*>(s.DZ);
    const uint32_t bytes_z = (uint32_t)(N * NZ * esz), bytes_h = (uint32_t)(N * 10 * esz), bytes_n = N * 4;
    s.rows_g = static_cast<const unsigned char*>(prm.rows) + (size_t)b * N * mcap * 4 * esz;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, bytes_z + bytes_h + bytes_n);
        tma_load(stg, static_cast<const unsigned char*>(prm.z0) + (size_t)b * N * NZ * esz, bytes_z, bar);
        tma_load(stg + GL::STG_HDR_BYTES, static_cast<const unsigned char*>(prm.hdr) + (size_t)b * N * 10 * esz, bytes_h, bar);
        tma_load(nr, prm.nrows + (size_t)b * N, bytes_n, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    for (int e = tid; e < N * NZ; e += NT)
        s.Z[e] = io32 ? (double)reinterpret_cast<const float*>(stg)[e] : reinterpret_cast<const double*>(stg)[e];
    for (int e = tid; e < N * 10; e += NT)
        s.HDR[(e / 10) * GL::HDR_S + (e % 10)] = io32 ? (double)reinterpret_cast<const float*>(stg + GL::STG_HDR_BYTES)[e]
                                                      : reinterpret_cast<const double*>(stg + GL::STG_HDR_BYTES)[e];
    __syncthreads();

    // ---- initial point ----
    if (tid < NZ) { s.BND[tid] = lower_bound<double>(tid); s.BND[NZ + tid] = upper_bound<double>(tid); }
    for (int e = tid; e < N * NZ; e += NT) s.DZ[e] = 0.0f;
    for (int e = tid; e < N * NXI; e += NT) { s.DY[e] = 0.0f; s.Y[e] = 0.0; }
    int ncomp = 0;
    for (int e = tid; e < N * NZ; e += NT) {
        const int k = e / NZ, i = e - k * NZ;
        double v = s.Z[e];
        if (k == 0 && i >= 8)
            v = io32 ? (double)static_cast<const float*>(prm.xinit)[(size_t)b * 9 + i - 8] : static_cast<const double*>(prm.xinit)[(size_t)b * 9 + i - 8];
        if (is_free(k, i)) {
            const double lb = lower_bound<double>(i), ub = upper_bound<double>(i), kp = o.kappa_push;
            const double pl = fmin(kp * fmax(1.0, fabs(lb)), kp * (ub - lb));
            const double pu = fmin(kp * fmax(1.0, fabs(ub)), kp * (ub - lb));
            v = fmin(fmax(v, lb + pl), ub - pu);
            s.ZL[e] = o.mu0 / (v - lb);
            s.ZU[e] = o.mu0 / (ub - v);
            ncomp += 2;
        } else {
            s.ZL[e] = 0.0;
            s.ZU[e] = 0.0;
        }
        s.Z[e] = v;
    }
    __syncthreads();
    for (int k = tid / 4; k < N; k += NT / 4) {
        const int m = s.live(k);
        for (int j = tid % 4; j < m; j += 4) {
            double r[4]; s.load_row(k, j, r);
            double sl = (r[3] + C::hu) - (r[0] * s.Z[k * NZ + 8] + r[1] * s.Z[k * NZ + 9] + r[2] * s.Z[k * NZ + 10]);
            sl = fmax(sl, o.s_floor);
            s.S[k * s.SS + j] = sl;
            s.LC[k * s.SS + j] = o.mu0 / sl;
            ncomp++;
        }
    }
    {
        double v[1] = {(double)ncomp};
        const int ops[1] = {0};
        s.template reduce<1>(v, ops);
        ncomp = (int)(v[0] + 0.5);
    }

    // ---- interior-point iterations ----
    int flag = 0, it = 0, nbt_total = 0;
    double alpha_p = 0.0, alpha_d = 0.0, rs_n = 0.0, req_n = 0.0, rin_n = 0.0, rcomp = 0.0, mu = 0.0;
    double f_cur, th_cur, ls_cur;
    s.evaluate(0.0, f_cur, th_cur, ls_cur, req_n, rs_n);
    const int it_cap = min(o.maxit, MIXED_BAIL_IT);
    bool infeasible0;
    {
        double v0 = 0.0;
        if (tid >= 8 && tid < NZ) v0 = fmax(lower_bound<double>(tid) - s.Z[tid], s.Z[tid] - upper_bound<double>(tid));
        const int m0 = min(nr[0], mcap);
        for (int j = tid; j < m0; j += NT) {
            double r[4]; s.load_row(0, j, r);
            v0 = fmax(v0, r[0] * s.Z[8] + r[1] * s.Z[9] + r[2] * s.Z[10] - (r[3] + C::hu));
        }
        double v[1] = {v0};
        const int ops[1] = {1};
        s.template reduce<1>(v, ops);
        infeasible0 = v[0] > o.tol_ineq;
        if (infeasible0) { flag = -7; rin_n = v[0]; }
    }
    for (it = 0; !infeasible0; it++) {
        double csum, cmin;
        s.residuals(rin_n, rcomp, csum, cmin);
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
        s.assemble(mu_t);
        __syncthreads();
        bool ok = s.riccati_backward();
        ok &= s.rollout_and_costates();
        if (tid == 0) s.RED[0] = ok ? 1.0 : 0.0;         // warp 0 holds the pivots' verdict: every thread takes the same branch on it
        __syncthreads();
        ok = s.RED[0] != 0.0;
        if (!ok) { flag = -5; break; }
        const double tau = fmin(fmax(0.995, 1.0 - mu), 0.99999);
        double ap, ad;
        s.step_lengths(mu_t, tau, ap, ad);
        s.update_duals(mu_t, ad);
        __syncthreads();
        const double ph0 = f_cur - mu_t * ls_cur;
        const double th_noise = fmax(10.0 * Eps<double>::v * double(N * NXI) * 20.0, 0.01 * o.tol_eq);
        double a = ap;
        int nbt = 0;
        double ft, tht, lst, reqt, rst;
        for (;;) {
            s.evaluate(a, ft, tht, lst, reqt, rst);
            const double pht = ft - mu_t * lst;
            const bool acc = (tht <= fmax((1.0 - 1e-5) * th_cur, th_noise)) ||
                             (pht <= ph0 - 1e-5 * th_cur + 10.0 * Eps<double>::v * fabs(ph0));
            if (acc || nbt >= o.max_bt) break;
            nbt++;
            a *= 0.5;
        }
        nbt_total += nbt;
        alpha_p = a; alpha_d = ad;
        s.update_primal(a);
        f_cur = ft; th_cur = tht; ls_cur = lst; req_n = reqt; rs_n = rst;
        __syncthreads();
    }

    // ---- results ----
    __syncthreads();
    auto put = [&](void* base, size_t idx, double v) {
        if (io32) static_cast<float*>(base)[idx] = (float)v; else static_cast<double*>(base)[idx] = v;
    };
    store_solution<NT>(prm.peers, prm.z_out, prm.info_int, (size_t)b, s.Z, N * NZ, io32, tid, flag, it, nbt_total, 0);
    if (tid == 0) {
        const double v[8] = {req_n, rin_n, rs_n, rcomp, f_cur, mu, alpha_p, alpha_d};
#pragma unroll
        for (int q = 0; q < 8; q++) put(prm.info_real, (size_t)b * 8 + q, v[q]);
    }
    if (prm.y_out)
        for (int e = tid; e < N * NXI; e += NT) put(prm.y_out, (size_t)b * N * NXI + e, (e < NXI) ? 0.0 : s.Y[e]);
    if (prm.zl_out)
        for (int e = tid; e < N * NZ; e += NT) put(prm.zl_out, (size_t)b * N * NZ + e, s.ZL[e]);
    if (prm.zu_out)
        for (int e = tid; e < N * NZ; e += NT) put(prm.zu_out, (size_t)b * N * NZ + e, s.ZU[e]);
    if (prm.lc_out)
        for (int e = tid; e < N * mcap; e += NT) {
            const int k = e / mcap, j = e - k * mcap;
            put(prm.lc_out, (size_t)b * N * mcap + e, (j < s.live(k)) ? s.LC[k * s.SS + j] : 0.0);
        }
}

}  // namespace nmpc
