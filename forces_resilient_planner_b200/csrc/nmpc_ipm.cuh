// nmpc_ipm.cuh -- fused batched interior-point NMPC solve, one warp (= one CTA) per problem.
//
// This kernel is the B200-native replacement for the whole of
//   FORCESNLPsolver_{normal,final}_solve  (closed ForcesPro v4.4.0 binary,
//   /root/reference/src/resilient_planner/plan_manage/solver/normal/FORCESNLPsolver_normal/include/FORCESNLPsolver_normal.h:321-323)
// including the per-stage model callbacks it drives (nmpc_model.cuh).  It is not a port: the
// reference factorises an interleaved 17/13 block LDL' on one CPU thread with static storage; here
// every problem lives in the shared memory of one warp for its entire solve:
//
//   * inputs (warm start, stage headers, corridor rows) arrive by TMA bulk copies
//     (cp.async.bulk ... mbarrier::complete_tx) and the solution leaves by a bulk store;
//   * "lanes = stages" phases (model evaluation, barrier terms, residual norms, step-to-boundary,
//     line-search merit) run one stage per lane and finish with warp-shuffle reductions;
//   * the KKT system is solved by a Riccati recursion over xi = [x(9); u_prev(4)] whose 13x13 /
//     9x13 / 4x13 products are spread over the 32 lanes ("lanes = matrix entries"), the 4x4
//     pivot block being factorised redundantly in registers;
//   * the iteration loop, convergence test and exit code are per warp, so a slow or diverging
//     instance never stalls another one.
//
// Algorithm (same, step for step, as oracle/nmpc_oracle.c -- which solves the identical KKT
// systems by a ForcesPro-style Schur complement instead, so the two check each other):
// primal-dual interior point, Gauss-Newton Hessian, mu_target = max(sigma*mu, mu_floor),
// fraction-to-boundary tau = min(max(0.995, 1-mu), 0.99999), backtracking on
// (theta, barrier objective), termination on the reference tolerances (1e-4 inf-norms,
// matlab_code/mpc/normal/mpc_generator_normal.m:76-79), iteration cap 200 (:56).
#pragma once
#include <cstdint>
#include "nmpc_model.cuh"

namespace nmpc {

struct Opts {
    double mu0, sigma, mu_floor, tol_stat, tol_eq, tol_ineq, tol_comp, kappa_push, s_floor;
    int maxit, max_bt;
};

template <typename T> struct Params {
    int B, mcap, variant;
    const T* xinit;    // [B][9]
    const T* z0;       // [B][N][17]
    const T* hdr;      // [B][N][10]
    const T* rows;     // [B][N][mcap][4]
    const int* nrows;  // [B][N]
    T* z_out;          // [B][N][17]
    int* info_int;     // [B][4]  exitflag, iterations, backtracks, reserved
    T* info_real;      // [B][8]  res_eq res_ineq rsnorm rcompnorm pobj mu alpha_p alpha_d
    // optional multiplier outputs (nullptr = not wanted); used by the KKT-acceptance tests
    T* y_out;          // [B][N][13]  equality multipliers, c-ordering, y[0] = 0
    T* zl_out;         // [B][N][17]  lower-bound multipliers
    T* zu_out;         // [B][N][17]  upper-bound multipliers
    T* lc_out;         // [B][N][mcap] corridor multipliers
    Opts o;
};

// ------------------------------------------------------------------ shared-memory layout ---
template <typename T, int N> struct Layout {
    static_assert(N % 4 == 0 && N >= 4 && N <= 64, "horizon must be a multiple of 4 (TMA 16-byte granules)");
    static constexpr int HDR_S = 11;    // padded stage-header stride (bank-conflict free)
    static constexpr int PHI_S = 21;    // 17 diagonal + 3 off-diagonal of the position block + u/u_prev coupling
    static constexpr int NR_BYTES = N * 4;
    static constexpr int HEAD_BYTES = 16 + NR_BYTES;   // mbarrier (8, padded to 16) + nrows
    // offsets in units of T from the start of the T region
    static constexpr int Z = 0;
    static constexpr int DZ = Z + N * NZ;
    static constexpr int ZL = DZ + N * NZ;
    static constexpr int ZU = ZL + N * NZ;
    static constexpr int G = ZU + N * NZ;
    static constexpr int Y = G + N * NZ;
    static constexpr int P = Y + N * NXI;
    static constexpr int D = P + N * NXI;
    static constexpr int JC = D + N * NXI;
    static constexpr int PHID = JC + N * NJC;           // last stage's slot unused by the solver
    static constexpr int KG = PHID + N * PHI_S;
    static constexpr int KFF = KG + N * 52;
    static constexpr int HDR = KFF + N * 4;
    // Riccati / rollout scratch
    static constexpr int PN = HDR + N * HDR_S;   // 13x13 cost-to-go
    static constexpr int FD = PN + 169;          // 9x13 dense dynamics Jacobian wrt (u, x)
    static constexpr int PF = FD + 117;          // 13x13 = PN(:, x) * FD
    static constexpr int GG = PF + 169;          // 13x13 = FD' * PF(x, :)
    static constexpr int TV = GG + 169;          // 13    = p+ + PN d
    static constexpr int FT = TV + 13;           // 13    = FD' TV(x)
    static constexpr int QUU = FT + 13;          // 4x4
    static constexpr int QUR = QUU + 16;         // 4x13  [Q_ux | Q_uq]
    static constexpr int QV = QUR + 52;          // 4
    static constexpr int QXI = QV + 4;           // 13
    static constexpr int YS = QXI + 13;          // 4x13  L^-1 QUR
    static constexpr int Y0 = YS + 52;           // 4     L^-1 QV
    static constexpr int DXI = Y0 + 4;           // 13
    static constexpr int FIXED_END = DXI + 13 + 2;
    // mcap-dependent tail: ROWS [N][4*mcap+1], S [N][mcap|1], LC [N][mcap|1]
    __host__ __device__ static constexpr int row_stride(int mcap) { return 4 * mcap + 1; }
    __host__ __device__ static constexpr int s_stride(int mcap) { return mcap | 1; }
    __host__ __device__ static constexpr int rows_off() { return FIXED_END; }
    __host__ __device__ static constexpr int s_off(int mcap) { return rows_off() + N * row_stride(mcap); }
    __host__ __device__ static constexpr int lc_off(int mcap) { return s_off(mcap) + N * s_stride(mcap); }
    __host__ __device__ static constexpr int total_T(int mcap) { return lc_off(mcap) + N * s_stride(mcap); }
    __host__ __device__ static constexpr size_t bytes(int mcap)
    {
        return (size_t)HEAD_BYTES + (size_t)total_T(mcap) * sizeof(T);
    }
    // TMA staging (stage headers + corridor rows as delivered) aliases DZ.. (dead until init)
    static constexpr int STG_HDR = DZ;
    static constexpr int STG_ROWS = DZ + N * 10;
    __host__ __device__ static constexpr bool staging_fits(int mcap) { return N * 10 + N * mcap * 4 <= PHID - DZ; }
};

// ------------------------------------------------------------------ small device helpers ---
template <typename T> __device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T> __device__ __forceinline__ T warp_max(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <typename T> __device__ __forceinline__ T warp_min(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <typename T> struct Eps;
template <> struct Eps<double> { static constexpr double v = 2.220446049250313e-16; static constexpr int loggrp = 8; };
template <> struct Eps<float> { static constexpr float v = 1.1920929e-07f; static constexpr int loggrp = 3; };

__device__ __forceinline__ int e_col(int i) { return i < 9 ? 8 + i : i - 5; }   // xi index -> z index
__device__ __forceinline__ bool is_free(int k, int i) { return k > 0 || i < 8; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// TMA 1-D bulk copy shared -> global
__device__ __forceinline__ void tma_store(void* dst, const void* src, uint32_t bytes)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// in-register Cholesky of a 4x4 SPD matrix given row-major a[16]; L (lower, 10 values) returned
// as l[10] = {l00, l10,l11, l20,l21,l22, l30,l31,l32,l33}; returns false on a non-positive pivot.
template <typename T> __device__ __forceinline__ bool chol4(const T* a, T l[10])
{
    bool ok = true;
    T d = a[0];
    ok &= d > T(0);
    l[0] = sqrt_t(d);
    T i0 = T(1) / l[0];
    l[1] = a[4] * i0; l[3] = a[8] * i0; l[6] = a[12] * i0;
    d = a[5] - l[1] * l[1];
    ok &= d > T(0);
    l[2] = sqrt_t(d);
    T i1 = T(1) / l[2];
    l[4] = (a[9] - l[3] * l[1]) * i1;
    l[7] = (a[13] - l[6] * l[1]) * i1;
    d = a[10] - l[3] * l[3] - l[4] * l[4];
    ok &= d > T(0);
    l[5] = sqrt_t(d);
    T i2 = T(1) / l[5];
    l[8] = (a[14] - l[6] * l[3] - l[7] * l[4]) * i2;
    d = a[15] - l[6] * l[6] - l[7] * l[7] - l[8] * l[8];
    ok &= d > T(0);
    l[9] = sqrt_t(d);
    return ok;
}
template <typename T> __device__ __forceinline__ void fsub4(const T l[10], T x[4])
{
    x[0] = x[0] / l[0];
    x[1] = (x[1] - l[1] * x[0]) / l[2];
    x[2] = (x[2] - l[3] * x[0] - l[4] * x[1]) / l[5];
    x[3] = (x[3] - l[6] * x[0] - l[7] * x[1] - l[8] * x[2]) / l[9];
}
template <typename T> __device__ __forceinline__ void bsub4(const T l[10], T x[4])
{
    x[3] = x[3] / l[9];
    x[2] = (x[2] - l[8] * x[3]) / l[5];
    x[1] = (x[1] - l[4] * x[2] - l[7] * x[3]) / l[2];
    x[0] = (x[0] - l[1] * x[1] - l[3] * x[2] - l[6] * x[3]) / l[0];
}

// =====================================================================================
// per-warp solver state: thin view over the shared-memory block
// =====================================================================================
template <typename T, int N> struct Solver {
    using L = Layout<T, N>;
    using C = Const<T>;
    T* sm;        // T region
    int* nr;      // live rows per stage
    int lane, mcap, RS, SS;
    bool final_variant;
    T *Z, *DZ, *ZL, *ZU, *G, *Y, *P, *D, *JC, *PHID, *KG, *KFF, *HDR, *ROWS, *S, *LC;
    T* fac_out = nullptr;   // when set, riccati_backward streams the factor (P | K | Quu^-1 | J) to HBM

    __device__ __forceinline__ int live(int k) const { return k == 0 ? 0 : min(nr[k], mcap); }

    // ---------------------------------------------------------------- model evaluation ---
    // FULL: at z, storing gradient / compact Jacobian / defects.  !FULL: at z + a dz (values only).
    template <bool FULL> __device__ void evaluate(T a, T& f_out, T& th_out, T& ls_out)
    {
        T f = T(0), th = T(0), ls = T(0), prod = T(1);
        int np = 0;
        auto acc_log = [&](T v) {
            prod *= v;
            if (++np == Eps<T>::loggrp) { ls += log_t(prod); prod = T(1); np = 0; }
        };
        for (int k = lane; k < N; k += 32) {
            T zk[NZ];
#pragma unroll
            for (int i = 0; i < NZ; i++) zk[i] = FULL ? Z[k * NZ + i] : Z[k * NZ + i] + a * DZ[k * NZ + i];
            const T* hdr = HDR + k * L::HDR_S;
            f += objective<T, FULL>(zk, hdr, k == 0, final_variant && k == N - 1, G + k * NZ);
            if (k < N - 1) {
                T c[NXI];
                dynamics<T, FULL>(zk, hdr + 3, c, JC + k * NJC);
#pragma unroll
                for (int i = 0; i < NXI; i++) {
                    const int zi = (k + 1) * NZ + e_col(i);
                    T zn = FULL ? Z[zi] : Z[zi] + a * DZ[zi];
                    T d = c[i] - zn;
                    th += fabs(d);
                    if (FULL) D[k * NXI + i] = d;
                }
            }
#pragma unroll
            for (int i = 0; i < NZ; i++)
                if (is_free(k, i)) {
                    acc_log(zk[i] - lower_bound<T>(i));
                    acc_log(upper_bound<T>(i) - zk[i]);
                }
            const int m = live(k);
            for (int j = 0; j < m; j++) {
                const T* r = ROWS + k * RS + 4 * j;
                T sj = S[k * SS + j];
                T rc = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + sj;
                if (!FULL) {
                    T adz = r[0] * DZ[k * NZ + 8] + r[1] * DZ[k * NZ + 9] + r[2] * DZ[k * NZ + 10];
                    sj += a * (-rc - adz);
                    rc *= (T(1) - a);
                }
                th += fabs(rc);
                acc_log(sj);
            }
        }
        ls += log_t(prod);
        f_out = warp_sum(f);
        th_out = warp_sum(th);
        ls_out = warp_sum(ls);
    }

    // ------------------------------------------------------- residual norms and mu ------
    __device__ void residuals(T& rs_n, T& req_n, T& rin_n, T& rcomp, T& csum, T& cmin)
    {
        T rs = T(0), req = T(0), rin = T(0), cmx = T(0), cs = T(0), cmn = T(1e30);
        for (int k = lane; k < N; k += 32) {
            const int m = live(k);
            T yn[NXI], yk[NXI];
#pragma unroll
            for (int i = 0; i < NXI; i++) {
                yn[i] = (k < N - 1) ? Y[(k + 1) * NXI + i] : T(0);
                yk[i] = (k > 0) ? Y[k * NXI + i] : T(0);
            }
            T al[3] = {T(0), T(0), T(0)};
            for (int j = 0; j < m; j++) {
                const T* r = ROWS + k * RS + 4 * j;
                T sj = S[k * SS + j], lj = LC[k * SS + j];
                al[0] += r[0] * lj; al[1] += r[1] * lj; al[2] += r[2] * lj;
                T rc = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + sj;
                T cc = sj * lj;
                cs += cc; cmx = fmax(cmx, cc); cmn = fmin(cmn, cc);
                rin = fmax(rin, fmax(fabs(rc), rc - sj));
            }
#pragma unroll
            for (int i = 0; i < NZ; i++) {
                if (!is_free(k, i)) continue;
                T zi = Z[k * NZ + i], zl = ZL[k * NZ + i], zu = ZU[k * NZ + i];
                T r = G[k * NZ + i] - zl + zu;
                if (k < N - 1) r += jt_y<T>(JC + k * NJC, yn, i);
                if (i >= 8) r -= yk[i - 8];
                else if (i >= 4) r -= yk[9 + i - 4];
                if (i >= 8 && i < 11) r += al[i - 8];
                rs = fmax(rs, fabs(r));
                T cl = (zi - lower_bound<T>(i)) * zl, cu = (upper_bound<T>(i) - zi) * zu;
                cs += cl + cu;
                cmx = fmax(cmx, fmax(cl, cu));
                cmn = fmin(cmn, fmin(cl, cu));
            }
            if (k < N - 1) {
#pragma unroll
                for (int i = 0; i < NXI; i++) req = fmax(req, fabs(D[k * NXI + i]));
            }
        }
        // NaN-propagating reductions: fmax drops NaNs, so carry a finite flag through the sum
        rs_n = warp_max(rs); req_n = warp_max(req); rin_n = warp_max(rin); rcomp = warp_max(cmx);
        csum = warp_sum(cs); cmin = warp_min(cmn);
    }

    // ------------------------------------------ barrier-augmented stage Hessian and rhs ---
    __device__ void assemble(T mu_t)
    {
        for (int k = lane; k < N; k += 32) {
            const T* hdr = HDR + k * L::HDR_S;
            const bool first = (k == 0), ft = final_variant && (k == N - 1);
            T* phi = PHID + k * L::PHI_S;
#pragma unroll
            for (int i = 0; i < NZ; i++) {
                if (is_free(k, i)) {
                    T zi = Z[k * NZ + i];
                    T isl = T(1) / (zi - lower_bound<T>(i)), isu = T(1) / (upper_bound<T>(i) - zi);
                    phi[i] = cost_hess_diag<T>(i, hdr, first, ft) + ZL[k * NZ + i] * isl + ZU[k * NZ + i] * isu;
                    G[k * NZ + i] += mu_t * (isu - isl);
                } else {
                    phi[i] = T(1);
                    G[k * NZ + i] = T(0);
                }
            }
            T o01 = T(0), o02 = T(0), o12 = T(0), d0 = T(0), d1 = T(0), d2 = T(0), g0 = T(0), g1 = T(0), g2 = T(0);
            const int m = live(k);
            for (int j = 0; j < m; j++) {
                const T* r = ROWS + k * RS + 4 * j;
                T sj = S[k * SS + j], lj = LC[k * SS + j], is = T(1) / sj;
                T rc = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + sj;
                T sg = lj * is, tt = (mu_t + lj * rc) * is;
                d0 += r[0] * r[0] * sg; d1 += r[1] * r[1] * sg; d2 += r[2] * r[2] * sg;
                o01 += r[0] * r[1] * sg; o02 += r[0] * r[2] * sg; o12 += r[1] * r[2] * sg;
                g0 += r[0] * tt; g1 += r[1] * tt; g2 += r[2] * tt;
            }
            phi[8] += d0; phi[9] += d1; phi[10] += d2;
            phi[17] = o01; phi[18] = o02; phi[19] = o12;
            phi[20] = T(-2) * hdr[8];   // H[u_i][uprev_i]
            if (k > 0) { G[k * NZ + 8] += g0; G[k * NZ + 9] += g1; G[k * NZ + 10] += g2; }
        }
    }

    // position-block entry of Phi_k in xi-ordering (i, j < 9)
    __device__ __forceinline__ T phi_xx(const T* phi, int i, int j) const
    {
        if (i == j) return phi[8 + i];
        if (i < 3 && j < 3) {
            int lo = min(i, j), hi = max(i, j);
            return phi[17 + lo + hi - 1];   // (0,1)->17, (0,2)->18, (1,2)->19
        }
        return T(0);
    }

    // ------------------------------------------------------------- Riccati backward -----
    // Returns false on a non-positive pivot.
    __device__ bool riccati_backward()
    {
        T* PN = sm + L::PN; T* FD = sm + L::FD; T* PF = sm + L::PF; T* GG = sm + L::GG;
        T* TV = sm + L::TV; T* FT = sm + L::FT; T* QUU = sm + L::QUU; T* QUR = sm + L::QUR;
        T* QV = sm + L::QV; T* QXI = sm + L::QXI; T* YS = sm + L::YS; T* Y0 = sm + L::Y0;
        bool ok = true;
        // lower-triangle entries (i >= j) of a 13x13 handled by this lane: e = lane, lane+32, lane+64
        int ti[3], tj[3];
#pragma unroll
        for (int t = 0; t < 3; t++) {
            const int e = lane + 32 * t;
            int i = 0;
            while ((i + 1) * (i + 2) / 2 <= e) i++;
            ti[t] = i;
            tj[t] = e - i * (i + 1) / 2;
        }
        for (int k = N - 1; k >= 0; k--) {
            const bool nx = (k < N - 1);
            const T* phi = PHID + k * L::PHI_S;
            const T* gk = G + k * NZ;
            if (nx) {
                // S1: dense F and tv = p+ + P+ d
                const T* jc = JC + k * NJC;
                for (int e = lane; e < 117; e += 32) FD[e] = f_dense<T>(jc, e / 13, e % 13);
                if (lane < NXI) {
                    T acc = P[(k + 1) * NXI + lane];
#pragma unroll
                    for (int j = 0; j < NXI; j++) acc += PN[lane * 13 + j] * D[k * NXI + j];
                    TV[lane] = acc;
                }
                __syncwarp();
                // S2: PF = PN(:, 0:9) * FD  (13x13),  ft = FD' tv(0:9)
                for (int e = lane; e < 169; e += 32) {
                    const int i = e / 13, j = e - 13 * i;
                    T acc = T(0);
#pragma unroll
                    for (int q = 0; q < 9; q++) acc += PN[i * 13 + q] * FD[q * 13 + j];
                    PF[e] = acc;
                }
                if (lane < 13) {
                    T acc = T(0);
#pragma unroll
                    for (int q = 0; q < 9; q++) acc += FD[q * 13 + lane] * TV[q];
                    FT[lane] = acc;
                }
                __syncwarp();
                // S3: GG = FD' * PF(0:9, :)  (symmetric 13x13; lower triangle, mirrored)
#pragma unroll
                for (int t = 0; t < 3; t++) {
                    if (lane + 32 * t >= 91) break;
                    const int i = ti[t], j = tj[t];
                    T acc = T(0);
#pragma unroll
                    for (int q = 0; q < 9; q++) acc += FD[q * 13 + i] * PF[q * 13 + j];
                    GG[i * 13 + j] = acc;
                    GG[j * 13 + i] = acc;
                }
                __syncwarp();
            }
            // S4: Q blocks.  v-ordering of FD columns: (u0..u3, x0..x8); xi-ordering: (x0..x8, q0..q3)
            const T wr2 = phi[20];
            for (int e = lane; e < 85; e += 32) {
                if (e < 16) {
                    const int a = e >> 2, b = e & 3;
                    T v = (a == b) ? phi[a] : T(0);
                    if (nx) v += GG[a * 13 + b] + PN[(9 + a) * 13 + 9 + b] + PF[(9 + a) * 13 + b] + PF[(9 + b) * 13 + a];
                    QUU[e] = v;
                } else if (e < 68) {
                    const int a = (e - 16) / 13, j = (e - 16) - 13 * a;
                    T v;
                    if (j < 9) v = nx ? GG[a * 13 + 4 + j] + PF[(9 + a) * 13 + 4 + j] : T(0);
                    else v = (j - 9 == a) ? wr2 : T(0);
                    QUR[a * 13 + j] = v;
                } else if (e < 72) {
                    const int a = e - 68;
                    QV[a] = gk[a] + (nx ? TV[9 + a] + FT[a] : T(0));
                } else {
                    const int j = e - 72;
                    QXI[j] = (j < 9) ? gk[8 + j] + (nx ? FT[4 + j] : T(0)) : gk[4 + j - 9];
                }
            }
            __syncwarp();
            // S5: factor the 4x4 pivot block (redundantly, in registers), solve the 13+1 columns
            T l[10];
            ok &= chol4<T>(QUU, l);
            if (lane < 14) {
                T x[4];
#pragma unroll
                for (int r = 0; r < 4; r++) x[r] = (lane < 13) ? QUR[r * 13 + lane] : QV[r];
                fsub4<T>(l, x);
                if (lane < 13) {
#pragma unroll
                    for (int r = 0; r < 4; r++) YS[r * 13 + lane] = x[r];
                } else {
#pragma unroll
                    for (int r = 0; r < 4; r++) Y0[r] = x[r];
                }
                bsub4<T>(l, x);
                if (lane < 13) {
#pragma unroll
                    for (int r = 0; r < 4; r++) KG[k * 52 + r * 13 + lane] = -x[r];
                } else {
#pragma unroll
                    for (int r = 0; r < 4; r++) KFF[k * 4 + r] = -x[r];
                }
            }
            __syncwarp();
            // S6: cost-to-go  P_k = blkdiag(Q_xx, Phi_qq) - YS' YS ,  p_k = q_xi - YS' y0
#pragma unroll
            for (int t = 0; t < 3; t++) {
                if (lane + 32 * t >= 91) break;
                const int i = ti[t], j = tj[t];
                T v = T(0);
                if (i < 9) v = phi_xx(phi, i, j) + (nx ? GG[(4 + i) * 13 + 4 + j] : T(0));
                else if (i == j) v = phi[4 + i - 9];
#pragma unroll
                for (int r = 0; r < 4; r++) v -= YS[r * 13 + i] * YS[r * 13 + j];
                PN[i * 13 + j] = v;
                PN[j * 13 + i] = v;
            }
            if (lane < 13) {
                T v = QXI[lane];
#pragma unroll
                for (int r = 0; r < 4; r++) v -= YS[r * 13 + lane] * Y0[r];
                P[k * NXI + lane] = v;
            }
            __syncwarp();
            if (fac_out) {   // factor block of stage k: [P_k packed lower 91 | K_k 52 | Quu^-1 packed lower 10 | J_k 51]
                T* fk = fac_out + (size_t)k * FAC_WORDS;
#pragma unroll
                for (int t = 0; t < 3; t++)
                    if (lane + 32 * t < 91) fk[lane + 32 * t] = PN[ti[t] * 13 + tj[t]];
                for (int e = lane; e < 52; e += 32) fk[91 + e] = KG[k * 52 + e];
                if (lane < 4) {   // column `lane` of Quu^-1 via the Cholesky factor
                    T x[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
                    for (int r = 0; r < 4; r++) x[r] = (r == lane) ? T(1) : T(0);
                    fsub4<T>(l, x);
                    bsub4<T>(l, x);
#pragma unroll
                    for (int r = 0; r < 4; r++)
                        if (r >= lane) fk[143 + r * (r + 1) / 2 + lane] = x[r];
                }
                for (int e = lane; e < NJC; e += 32) fk[153 + e] = nx ? JC[k * NJC + e] : T(0);
            }
        }
        return ok;
    }

    // ------------------------------------------------------------- forward rollout ------
    __device__ bool rollout()
    {
        T* PN = sm + L::PN; T* DXI = sm + L::DXI;
        bool ok = true;
        // stage 0: x fixed (dx = 0), u_prev free: dq = -Pqq^-1 p_q
        {
            T a[16], l[10], x[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
#pragma unroll
                for (int c = 0; c < 4; c++) a[4 * r + c] = PN[(9 + r) * 13 + 9 + c];
                x[r] = -P[9 + r];
            }
            ok = chol4<T>(a, l);
            fsub4<T>(l, x);
            bsub4<T>(l, x);
            if (lane < 13) DXI[lane] = (lane < 9) ? T(0) : x[lane - 9];
        }
        __syncwarp();
        for (int k = 0; k < N; k++) {
            // du_r = kff_r + K[r][:] . dxi   (8 lanes per row, shuffle-reduced)
            const int r = lane >> 3, part = lane & 7;
            T acc = T(0);
            for (int i = part; i < 13; i += 8) acc += KG[k * 52 + r * 13 + i] * DXI[i];
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += KFF[k * 4 + r];
            T v[13];   // (du(4), dx(9))
#pragma unroll
            for (int q = 0; q < 4; q++) v[q] = __shfl_sync(0xffffffffu, acc, 8 * q);
#pragma unroll
            for (int q = 0; q < 9; q++) v[4 + q] = DXI[q];
            const T du_l = __shfl_sync(0xffffffffu, acc, 8 * (lane & 3));        // du[lane & 3]
            const T du_q = __shfl_sync(0xffffffffu, acc, 8 * ((lane - 9) & 3));  // du[lane - 9] for lanes 9..12
            if (lane < NZ) DZ[k * NZ + lane] = (lane < 4) ? du_l : (lane < 8 ? DXI[5 + lane] : DXI[lane - 8]);
            T nxt = T(0);
            if (k < N - 1 && lane < 13) {
                if (lane < 9) {
                    const T* jc = JC + k * NJC;
                    nxt = D[k * NXI + lane];
#pragma unroll
                    for (int c = 0; c < 13; c++) nxt += f_dense<T>(jc, lane, c) * v[c];
                } else {
                    nxt = du_q + D[k * NXI + lane];
                }
            }
            __syncwarp();
            if (k < N - 1 && lane < 13) DXI[lane] = nxt;
            __syncwarp();
        }
        return ok;
    }

    // ------------------------------------ costates of the QP (new equality multipliers) ---
    // y_k = [Phi_k dz_k + g~_k + J_k' y_{k+1}]_xi , stored over p_k (c-ordering [x; q]).
    __device__ void costates()
    {
        for (int k = N - 1; k >= 1; k--) {
            T v = T(0);
            if (lane < 13) {
                const T* phi = PHID + k * L::PHI_S;
                const T* dz = DZ + k * NZ;
                if (lane < 9) {
                    v = G[k * NZ + 8 + lane] + phi[8 + lane] * dz[8 + lane];
                    if (lane < 3) {
#pragma unroll
                        for (int j = 0; j < 3; j++)
                            if (j != lane) v += phi_xx(phi, lane, j) * dz[8 + j];
                    }
                    if (k < N - 1) v += jt_y<T>(JC + k * NJC, P + (k + 1) * NXI, 8 + lane);
                } else {
                    const int c = lane - 9;
                    v = G[k * NZ + 4 + c] + phi[4 + c] * dz[4 + c] + phi[20] * dz[c];
                }
            }
            __syncwarp();
            if (lane < 13) P[k * NXI + lane] = v;
            __syncwarp();
        }
    }

    // ------------------------------------------- multiplier steps, fraction to boundary ---
    __device__ void step_lengths(T mu_t, T tau, T& ap_out, T& ad_out)
    {
        T ap = T(1), ad = T(1);
        for (int k = lane; k < N; k += 32) {
#pragma unroll
            for (int i = 0; i < NZ; i++) {
                if (!is_free(k, i)) continue;
                T zi = Z[k * NZ + i], dzi = DZ[k * NZ + i], zl = ZL[k * NZ + i], zu = ZU[k * NZ + i];
                T sl = zi - lower_bound<T>(i), su = upper_bound<T>(i) - zi;
                T dzl = (mu_t - zl * dzi) / sl - zl;
                T dzu = (mu_t + zu * dzi) / su - zu;
                if (dzi < T(0)) ap = fmin(ap, -tau * sl / dzi);
                if (dzi > T(0)) ap = fmin(ap, tau * su / dzi);
                if (dzl < T(0)) ad = fmin(ad, -tau * zl / dzl);
                if (dzu < T(0)) ad = fmin(ad, -tau * zu / dzu);
            }
            const int m = live(k);
            for (int j = 0; j < m; j++) {
                const T* r = ROWS + k * RS + 4 * j;
                T sj = S[k * SS + j], lj = LC[k * SS + j];
                T rc = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + sj;
                T ds = -rc - (r[0] * DZ[k * NZ + 8] + r[1] * DZ[k * NZ + 9] + r[2] * DZ[k * NZ + 10]);
                T dl = (mu_t - lj * ds) / sj - lj;
                if (ds < T(0)) ap = fmin(ap, -tau * sj / ds);
                if (dl < T(0)) ad = fmin(ad, -tau * lj / dl);
            }
        }
        ap_out = warp_min(ap);
        ad_out = warp_min(ad);
    }

    // --------------------------------------------------------------- accept the step ----
    __device__ void update(T mu_t, T a, T ad)
    {
        for (int k = lane; k < N; k += 32) {
            const int m = live(k);
            for (int j = 0; j < m; j++) {
                const T* r = ROWS + k * RS + 4 * j;
                T sj = S[k * SS + j], lj = LC[k * SS + j];
                T rc = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + sj;
                T ds = -rc - (r[0] * DZ[k * NZ + 8] + r[1] * DZ[k * NZ + 9] + r[2] * DZ[k * NZ + 10]);
                T dl = (mu_t - lj * ds) / sj - lj;
                S[k * SS + j] = sj + a * ds;
                LC[k * SS + j] = lj + ad * dl;
            }
#pragma unroll
            for (int i = 0; i < NZ; i++) {
                T zi = Z[k * NZ + i], dzi = DZ[k * NZ + i];
                if (is_free(k, i)) {
                    T zl = ZL[k * NZ + i], zu = ZU[k * NZ + i];
                    T sl = zi - lower_bound<T>(i), su = upper_bound<T>(i) - zi;
                    ZL[k * NZ + i] = zl + ad * ((mu_t - zl * dzi) / sl - zl);
                    ZU[k * NZ + i] = zu + ad * ((mu_t + zu * dzi) / su - zu);
                }
                Z[k * NZ + i] = zi + a * dzi;
            }
            if (k >= 1) {
#pragma unroll
                for (int i = 0; i < NXI; i++) Y[k * NXI + i] += a * (P[k * NXI + i] - Y[k * NXI + i]);
            }
        }
    }
};

// =====================================================================================
// the kernel: grid = B CTAs of one warp; dynamic smem = Layout::bytes(mcap)
// =====================================================================================
template <typename T, int N>
__global__ void __launch_bounds__(32) nmpc_ipm_kernel(const Params<T> prm)
{
    using L = Layout<T, N>;
    using C = Const<T>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int b = blockIdx.x;
    if (b >= prm.B) return;
    const int mcap = prm.mcap;

    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    int* nr = reinterpret_cast<int*>(smem_raw + 16);
    T* sm = reinterpret_cast<T*>(smem_raw + L::HEAD_BYTES);

    Solver<T, N> s;
    s.sm = sm; s.nr = nr; s.lane = lane; s.mcap = mcap;
    s.RS = L::row_stride(mcap); s.SS = L::s_stride(mcap);
    s.final_variant = (prm.variant == 1);
    s.Z = sm + L::Z; s.DZ = sm + L::DZ; s.ZL = sm + L::ZL; s.ZU = sm + L::ZU; s.G = sm + L::G;
    s.Y = sm + L::Y; s.P = sm + L::P; s.D = sm + L::D; s.JC = sm + L::JC; s.PHID = sm + L::PHID;
    s.KG = sm + L::KG; s.KFF = sm + L::KFF; s.HDR = sm + L::HDR;
    s.ROWS = sm + L::rows_off(); s.S = sm + L::s_off(mcap); s.LC = sm + L::lc_off(mcap);
    const Opts& o = prm.o;

    // ---- stage the problem into shared memory with TMA bulk copies ------------------------
    const uint32_t bytes_z = N * NZ * sizeof(T), bytes_h = N * 10 * sizeof(T);
    const uint32_t bytes_r = (uint32_t)N * mcap * 4 * sizeof(T), bytes_n = N * 4;
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, bytes_z + bytes_h + bytes_r + bytes_n);
        tma_load(s.Z, prm.z0 + (size_t)b * N * NZ, bytes_z, bar);
        tma_load(sm + L::STG_HDR, prm.hdr + (size_t)b * N * 10, bytes_h, bar);
        if (bytes_r) tma_load(sm + L::STG_ROWS, prm.rows + (size_t)b * N * mcap * 4, bytes_r, bar);
        tma_load(nr, prm.nrows + (size_t)b * N, bytes_n, bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    // re-layout headers / rows into bank-conflict-free padded strides
    {
        // headers: N*10 -> N*11 ; the regions do not overlap (HDR lives beyond the staging alias)
        for (int e = lane; e < N * 10; e += 32) s.HDR[(e / 10) * L::HDR_S + (e % 10)] = sm[L::STG_HDR + e];
        const int nrw = N * mcap * 4;
        for (int e = lane; e < nrw; e += 32) {
            const int k = e / (mcap * 4), q = e - k * mcap * 4;
            s.ROWS[k * s.RS + q] = sm[L::STG_ROWS + e];
        }
    }
    __syncwarp();

    // ---- initial point -------------------------------------------------------------------
    int ncomp = 0;
    for (int k = lane; k < N; k += 32) {
#pragma unroll
        for (int i = 0; i < NZ; i++) {
            T v = s.Z[k * NZ + i];
            if (k == 0 && i >= 8) v = prm.xinit[(size_t)b * 9 + i - 8];
            if (is_free(k, i)) {
                const T lb = lower_bound<T>(i), ub = upper_bound<T>(i), kp = (T)o.kappa_push;
                const T pl = fmin(kp * fmax(T(1), fabs(lb)), kp * (ub - lb));
                const T pu = fmin(kp * fmax(T(1), fabs(ub)), kp * (ub - lb));
                v = fmin(fmax(v, lb + pl), ub - pu);
                s.ZL[k * NZ + i] = (T)o.mu0 / (v - lb);
                s.ZU[k * NZ + i] = (T)o.mu0 / (ub - v);
                ncomp += 2;
            } else {
                s.ZL[k * NZ + i] = T(0);
                s.ZU[k * NZ + i] = T(0);
            }
            s.Z[k * NZ + i] = v;
        }
#pragma unroll
        for (int i = 0; i < NXI; i++) s.Y[k * NXI + i] = T(0);
        const int m = s.live(k);
        for (int j = 0; j < m; j++) {
            const T* r = s.ROWS + k * s.RS + 4 * j;
            T sl = (r[3] + C::hu) - (r[0] * s.Z[k * NZ + 8] + r[1] * s.Z[k * NZ + 9] + r[2] * s.Z[k * NZ + 10]);
            sl = fmax(sl, (T)o.s_floor);
            s.S[k * s.SS + j] = sl;
            s.LC[k * s.SS + j] = (T)o.mu0 / sl;
            ncomp++;
        }
    }
    ncomp = warp_sum(ncomp);
    __syncwarp();

    // ---- interior-point iterations ---------------------------------------------------------
    int flag = 0, it = 0, nbt_total = 0;
    T alpha_p = T(0), alpha_d = T(0), rs_n = T(0), req_n = T(0), rin_n = T(0), rcomp = T(0), mu = T(0);
    T f_cur, th_cur, ls_cur;
    s.template evaluate<true>(T(0), f_cur, th_cur, ls_cur);
    __syncwarp();
    for (it = 0;; it++) {
        T csum, cmin;
        s.residuals(rs_n, req_n, rin_n, rcomp, csum, cmin);
        mu = csum / (T)ncomp;
        const bool finite = isfinite(rs_n) && isfinite(req_n) && isfinite(mu) && isfinite(f_cur) && isfinite(th_cur);
        if (!finite) { flag = (it == 0) ? -6 : -7; break; }
        if (rs_n <= (T)o.tol_stat && req_n <= (T)o.tol_eq && rin_n <= (T)o.tol_ineq && rcomp <= (T)o.tol_comp) { flag = 1; break; }
        if (it >= o.maxit) { flag = 0; break; }
        T sigma = (T)o.sigma;
        if (sigma <= T(0)) {   // LOQO centrality rule
            const T xi = cmin / mu;
            const T q = fmin(T(0.05) * (T(1) - xi) / xi, T(2));
            sigma = T(0.1) * q * q * q;
        }
        const T mu_t = fmax(sigma * mu, (T)o.mu_floor);
        s.assemble(mu_t);
        __syncwarp();
        bool ok = s.riccati_backward();
        ok &= s.rollout();
        if (!ok) { flag = -5; break; }
        s.costates();
        const T tau = fmin(fmax(T(0.995), T(1) - mu), T(0.99999));
        T ap, ad;
        s.step_lengths(mu_t, tau, ap, ad);
        // backtracking line search on (theta, barrier objective)
        const T ph0 = f_cur - mu_t * ls_cur;
        // theta below 1% of TolEq counts as feasible (also absorbs the rounding floor of theta)
        const T th_noise = fmax(T(10) * Eps<T>::v * T(N * NXI) * T(20), T(0.01) * (T)o.tol_eq);
        T a = ap;
        int nbt = 0;
        for (;;) {
            T ft, tht, lst;
            s.template evaluate<false>(a, ft, tht, lst);
            const T pht = ft - mu_t * lst;
            const bool acc = (tht <= fmax((T(1) - T(1e-5)) * th_cur, th_noise)) ||
                             (pht <= ph0 - T(1e-5) * th_cur + T(10) * Eps<T>::v * fabs(ph0));
            if (acc || nbt >= o.max_bt) break;
            nbt++;
            a *= T(0.5);
        }
        nbt_total += nbt;
        alpha_p = a; alpha_d = ad;
        s.update(mu_t, a, ad);
        __syncwarp();
        s.template evaluate<true>(T(0), f_cur, th_cur, ls_cur);
        __syncwarp();
    }

    // ---- results -----------------------------------------------------------------------------
    __syncwarp();
    if (lane == 0) {
        tma_store(prm.z_out + (size_t)b * N * NZ, s.Z, bytes_z);
        int* ii = prm.info_int + (size_t)b * 4;
        ii[0] = flag; ii[1] = it; ii[2] = nbt_total; ii[3] = 0;
        T* ir = prm.info_real + (size_t)b * 8;
        ir[0] = req_n; ir[1] = rin_n; ir[2] = rs_n; ir[3] = rcomp;
        ir[4] = f_cur; ir[5] = mu; ir[6] = alpha_p; ir[7] = alpha_d;
    }
    if (prm.y_out)
        for (int e = lane; e < N * NXI; e += 32) prm.y_out[(size_t)b * N * NXI + e] = (e < NXI) ? T(0) : s.Y[e];
    if (prm.zl_out)
        for (int e = lane; e < N * NZ; e += 32) prm.zl_out[(size_t)b * N * NZ + e] = s.ZL[e];
    if (prm.zu_out)
        for (int e = lane; e < N * NZ; e += 32) prm.zu_out[(size_t)b * N * NZ + e] = s.ZU[e];
    if (prm.lc_out)
        for (int e = lane; e < N * mcap; e += 32) {
            const int k = e / mcap, j = e - k * mcap;
            prm.lc_out[(size_t)b * N * mcap + e] = (j < s.live(k)) ? s.LC[k * s.SS + j] : T(0);
        }
}

}  // namespace nmpc
