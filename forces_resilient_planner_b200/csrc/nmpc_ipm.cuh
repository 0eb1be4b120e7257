// nmpc_ipm.cuh -- fused batched interior-point NMPC solve, one warp (= one CTA) per problem.
//
// This kernel is the B200-native replacement for the whole of
//   FORCESNLPsolver_{normal,final}_solve  (closed ForcesPro v4.4.0 binary,
//   /root/reference/src/resilient_planner/plan_manage/solver/normal/FORCESNLPsolver_normal/include/FORCESNLPsolver_normal.h:321-323)
// including the per-stage model callbacks it drives (nmpc_model.cuh).  It is not a port: the
// reference factorises an interleaved 17/13 block LDL' on one CPU thread with static storage; here
// every problem lives in the shared memory of one warp for its entire solve:
//
//   * inputs (warm start, stage headers, row counts) arrive by TMA bulk copies
//     (cp.async.bulk ... mbarrier::complete_tx) and the solution leaves by a bulk store; the
//     read-only corridor rows are read through L1 (LDG.128) where they are used;
//   * "lanes = stages" phases (model evaluation, barrier terms, residual norms, step-to-boundary,
//     line-search merit) run one stage per lane and finish with warp-shuffle reductions;
//   * the KKT system is solved by a Riccati recursion over xi = [x(9); u_prev(4)]: the products with
//     the structured dynamics Jacobian run one row / column per lane with the Jacobian words
//     broadcast, the rank-4 update of the cost-to-go with "lanes = matrix entries", the 4x4
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
// Option (template parameter PC, opts.pc): Mehrotra predictor-corrector through one factorisation --
// affine solve, sigma = (mu_aff / mu)^3, second-order terms, corrector as a vector-only sweep
// (delta_backward / rollout<true>) with the stored gains and Quu^-1.
#pragma once
#include <cstdint>
#include "nmpc_model.cuh"

namespace nmpc {

struct Opts {
    double mu0, sigma, mu_floor, tol_stat, tol_eq, tol_ineq, tol_comp, kappa_push, s_floor;
    int maxit, max_bt, pc, mixed;
};

// Multi-GPU collation by peer stores (nmpc_peers_* of the C ABI): base pointers of THIS rank's slice inside every other
// rank's copy of the collation buffers, mapped into this process by CUDA IPC.  The epilogue of a solve writes the
// solution and the four info integers to its own buffers and to all of these, so the exchange rides along with the
// compute (NVLink stores spread over the whole kernel) and nothing is left to gather afterwards but a barrier.  The
// host-pointer entry points use the same mechanism with ONE extra destination: the caller's pinned result buffer, so the
// solutions cross PCIe while the batch is still being solved instead of in a copy after the last kernel.
constexpr int MAX_PEERS = 15;
struct PeerOut {
    int n = 0;
    void* z[MAX_PEERS];
    int* info[MAX_PEERS];      // may be null: solution only
};

// Problem data and results are arrays of T in HBM, or -- io32, fp64 kernel only: the re-solve of the
// problems the mixed-precision kernel gave up on (nmpc_ipm_mixed.cuh) -- arrays of float.
template <typename T> struct Params {
    int B, mcap, variant, io32;
    const void* xinit;    // [B][9]
    const void* z0;       // [B][N][17]
    const void* hdr;      // [B][N][10]
    const void* rows;     // [B][N][mcap][4]
    const int* nrows;  // [B][N]
    const int* order;  // [B] or nullptr: CTA i solves problem order[i] (launch order = scheduling order)
    const int* count;  // nullptr, or device counter: only the first *count entries of `order` are live
    const void* z_warm;   // re-solve only (count != nullptr): the mixed-precision kernel's last iterate [B][N][17]; a problem it
                          // gave up on for a factorisation breakdown (-5) or at its iteration cap (0) restarts from there
    void* z_out;          // [B][N][17]
    int* info_int;     // [B][4]  exitflag, iterations, backtracks, 1 if this is a re-solve (count != nullptr)
    void* info_real;      // [B][8]  res_eq res_ineq rsnorm rcompnorm pobj mu alpha_p alpha_d
    // optional multiplier outputs (nullptr = not wanted); used by the KKT-acceptance tests
    void* y_out;          // [B][N][13]  equality multipliers, c-ordering, y[0] = 0
    void* zl_out;         // [B][N][17]  lower-bound multipliers
    void* zu_out;         // [B][N][17]  upper-bound multipliers
    void* lc_out;         // [B][N][mcap] corridor multipliers
    PeerOut peers;        // n = 0: single GPU
    Opts o;
};

// ------------------------------------------------------------------ shared-memory layout ---
// Three groups of arrays (offsets in units of T from the start of the T region):
//   R  "stage-phase" state: z, z_l, z_u, y, headers, bounds table, corridor slacks/multipliers.
//      Untouched by the KKT sweeps (Riccati, rollout, costates).  The corridor rows themselves are
//      read-only and identical in every iteration: they stay in global memory and are read through
//      L1 (32 bytes per row, one row per lane and step) -- that buys the sixth resident warp per SM.
//   O  sweep-private arrays: Riccati gains K, feed-forward terms and the Riccati scratch.
//      Dead outside the sweeps.
//   SH shared by both: dz, gradient, p/y_new, defects, compact Jacobians, Phi diagonals.
// O is OVERLAID on R: before the sweeps every lane parks its slice of the first NPARK words of R in
// registers (the sweeps use few registers, the evaluation phases that need many do not run then),
// and restores it afterwards.  That is 15.6 KB less shared memory per problem (fp64, N = 20) and
// one more resident warp per SM -- occupancy is what bounds this kernel.
template <typename T, int N, bool PC = false> struct Layout {
    static_assert(N % 4 == 0 && N >= 4 && N <= 64, "horizon must be a multiple of 4 (TMA 16-byte granules)");
    static constexpr int HDR_S = 11;    // padded stage-header stride (bank-conflict free)
    static constexpr int PHI_S = 21;    // 17 diagonal + 3 off-diagonal of the position block + u/u_prev coupling
    static constexpr int NR_BYTES = N * 4;
    static constexpr int HEAD_BYTES = 16 + NR_BYTES;   // mbarrier (8, padded to 16) + nrows
    // ---- R (fixed part) ----
    static constexpr int Z = 0;
    static constexpr int ZL = Z + N * NZ;
    static constexpr int ZU = ZL + N * NZ;
    static constexpr int Y = ZU + N * NXI + N * (NZ - NXI);   // = ZU + N*NZ
    static constexpr int HDR = Y + N * NXI;
    static constexpr int BND = HDR + N * HDR_S;               // lb(17) | ub(17)
    static constexpr int R_FIXED = BND + 2 * NZ;
    __host__ __device__ static constexpr int s_stride(int mcap) { return mcap | 1; }
    __host__ __device__ static constexpr int s_off(int) { return R_FIXED; }
    __host__ __device__ static constexpr int lc_off(int mcap) { return R_FIXED + N * s_stride(mcap); }
    __host__ __device__ static constexpr int r_end(int mcap) { return R_FIXED + 2 * N * s_stride(mcap); }
    // ---- O (overlay at offset 0): Riccati / rollout scratch first, gains last ----
    static constexpr int PN = 0;                 // 13x13 cost-to-go
    static constexpr int PF = PN + 169;          // 13x13 = PN(:, x) * F    (F = d x+ / d(u, x), 9x13, never formed)
    static constexpr int GG = PC ? PF : PF + 169;   // 13x13 = F' * PF(x, :); with PC it overwrites PF (after a warp barrier)
    static constexpr int TV = GG + 169;          // 13    = p+ + PN d
    static constexpr int QUU = TV + 13;          // 4x4
    static constexpr int QUR = QUU + 16;         // 4x13  [Q_ux | Q_uq]
    static constexpr int QV = QUR + 52;          // 4
    static constexpr int QXI = QV + 4;           // 13
    static constexpr int YS = QXI + 13;          // 4x13  L^-1 QUR
    static constexpr int Y0 = YS + 52;           // 4     L^-1 QV
    static constexpr int DXI = Y0 + 4;           // 13
    static constexpr int KFF = DXI + 14;
    static constexpr int QINV = KFF + N * 4;                 // PC: Quu^-1 packed lower per stage (the corrector re-uses the factorisation)
    static constexpr int PQQ0 = QINV + (PC ? N * 10 : 0);    // PC: the q-block of P_0 (stage-0 solve of the corrector)
    static constexpr int KG = PQQ0 + (PC ? 10 : 0);
    static constexpr int O_END = KG + N * 52;
    // words parked in registers across the sweeps: at most 64 per lane; if the overlay is larger
    // (long horizons) the gains are placed in SH instead of being overlaid
    static constexpr bool KG_OVERLAID = (O_END + 31) / 32 <= 64;
    static constexpr int O_USED = KG_OVERLAID ? O_END : KG;
    static constexpr int NPARK_LANE = (O_USED + 31) / 32;
    static constexpr int NPARK = NPARK_LANE * 32;
    // ---- SH: starts after max(R, parked overlay), 4-word aligned (16 B for fp32 and fp64 TMA) ----
    __host__ __device__ static constexpr int sh_off(int mcap)
    {
        return ((r_end(mcap) > NPARK ? r_end(mcap) : NPARK) + 3) & ~3;
    }
    static constexpr int SH_DZ = 0;
    static constexpr int SH_G = SH_DZ + N * NZ;
    static constexpr int SH_P = SH_G + N * NZ;
    static constexpr int SH_D = SH_P + N * NXI;
    static constexpr int SH_JC = SH_D + N * NXI;
    static constexpr int SH_PHID = SH_JC + N * NJC;         // last stage's Jacobian slot unused by the solver
    static constexpr int SH_KG = SH_PHID + N * PHI_S;       // only when the gains are not overlaid
    static constexpr int SH_DZAP = SH_KG + (KG_OVERLAID ? 0 : N * 52);   // PC: position part of the affine step, [N][3]
    static constexpr int SH_END = SH_DZAP + (PC ? N * 3 : 0);
    static constexpr int NT = (N * NZ + 31) / 32;           // flat (stage, variable) pairs per lane
    __host__ __device__ static constexpr int total_T(int mcap) { return sh_off(mcap) + SH_END; }
    __host__ __device__ static constexpr size_t bytes(int mcap)
    {
        return (size_t)HEAD_BYTES + (size_t)total_T(mcap) * sizeof(T);
    }
    // TMA staging (stage headers as delivered) lands at the start of SH (dead until init)
    static constexpr int STG_HDR = 0;
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
template <> struct Eps<double> { static constexpr double v = 2.220446049250313e-16; static constexpr int logvars = 9, logrows = 8; };
template <> struct Eps<float> { static constexpr float v = 1.1920929e-07f; static constexpr int logvars = 1, logrows = 2; };

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
// The solution [nz] and the info integers of problem b to this GPU's arrays and to every peer's (PeerOut).  NT threads of
// the CTA take part (tid); fp64 results leave shared memory by TMA bulk stores (one per destination, one commit group),
// float results (io32) by coalesced stores of the converted values.
template <int NT, typename T>
__device__ __forceinline__ void store_solution(const PeerOut& po, void* z_out, int* info_int, size_t b, const T* Zs, int nz, bool io32,
                                               int tid, int flag, int it, int nbt, int resolved)
{
    if (io32) {
        for (int e = tid; e < nz; e += NT) {
            const float v = (float)Zs[e];
            static_cast<float*>(z_out)[b * nz + e] = v;
            for (int p = 0; p < po.n; p++) static_cast<float*>(po.z[p])[b * nz + e] = v;
        }
    } else if (tid == 0) {
        const uint32_t bytes = (uint32_t)nz * sizeof(T);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(static_cast<T*>(z_out) + b * nz), "r"(smem_u32(Zs)),
                     "r"(bytes)
                     : "memory");
        for (int p = 0; p < po.n; p++)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(static_cast<T*>(po.z[p]) + b * nz),
                         "r"(smem_u32(Zs)), "r"(bytes)
                         : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    if (tid == 0) {
        int* ii = info_int + b * 4;
        ii[0] = flag; ii[1] = it; ii[2] = nbt; ii[3] = resolved;
        for (int p = 0; p < po.n; p++)
            if (po.info[p]) *reinterpret_cast<int4*>(po.info[p] + b * 4) = make_int4(flag, it, nbt, resolved);
    }
}

// 1/sqrt(x) and 1/x in double precision from the SFU's double-precision seeds (rsqrt.approx.ftz.f64 / rcp.approx.ftz.f64:
// MUFU.RSQ64H / MUFU.RCP64H on the upper word, relative error 2^-22) and two Newton steps (2^-22 -> 2^-44 -> rounding
// level): about ten instructions on the critical path of every pivot / barrier slack instead of the ~30 of the library
// routine or of an IEEE division, and no double <-> float conversions.  Arguments are pivots of the 4x4 blocks and
// barrier slacks: normal, positive numbers (a non-positive pivot is caught before its root is used).
__device__ __forceinline__ double rsqrt_t(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = y * (1.5 - hx * y * y);
    y = y * (1.5 - hx * y * y);
    return y;
}
__device__ __forceinline__ float rsqrt_t(float x) { return rsqrtf(x); }
__device__ __forceinline__ double rcp_t(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = r * (2.0 - x * r);
    r = r * (2.0 - x * r);
    return r;
}
__device__ __forceinline__ float rcp_t(float x) { return __frcp_rn(x); }
// in-register Cholesky of a 4x4 SPD matrix given row-major a[16] (lower triangle used); L (10 values)
// as l[10] = {l00, l10,l11, l20,l21,l22, l30,l31,l32,l33}, li[4] = 1/diag (no divisions anywhere:
// 1/sqrt(d) comes from rsqrt); returns false on a non-positive pivot.
template <typename T> __device__ __forceinline__ bool chol4(const T* a, T l[10], T li[4])
{
    bool ok = true;
    T d = a[0];
    ok &= d > T(0);
    li[0] = rsqrt_t(d); l[0] = d * li[0];
    l[1] = a[4] * li[0]; l[3] = a[8] * li[0]; l[6] = a[12] * li[0];
    d = a[5] - l[1] * l[1];
    ok &= d > T(0);
    li[1] = rsqrt_t(d); l[2] = d * li[1];
    l[4] = (a[9] - l[3] * l[1]) * li[1];
    l[7] = (a[13] - l[6] * l[1]) * li[1];
    d = a[10] - l[3] * l[3] - l[4] * l[4];
    ok &= d > T(0);
    li[2] = rsqrt_t(d); l[5] = d * li[2];
    l[8] = (a[14] - l[6] * l[3] - l[7] * l[4]) * li[2];
    d = a[15] - l[6] * l[6] - l[7] * l[7] - l[8] * l[8];
    ok &= d > T(0);
    li[3] = rsqrt_t(d); l[9] = d * li[3];
    return ok;
}
template <typename T> __device__ __forceinline__ void fsub4(const T l[10], const T li[4], T x[4])
{
    x[0] = x[0] * li[0];
    x[1] = (x[1] - l[1] * x[0]) * li[1];
    x[2] = (x[2] - l[3] * x[0] - l[4] * x[1]) * li[2];
    x[3] = (x[3] - l[6] * x[0] - l[7] * x[1] - l[8] * x[2]) * li[3];
}
template <typename T> __device__ __forceinline__ void bsub4(const T l[10], const T li[4], T x[4])
{
    x[3] = x[3] * li[3];
    x[2] = (x[2] - l[8] * x[3]) * li[2];
    x[1] = (x[1] - l[4] * x[2] - l[7] * x[3]) * li[1];
    x[0] = (x[0] - l[1] * x[1] - l[3] * x[2] - l[6] * x[3]) * li[0];
}
// dot product of n (compile-time) strided shared-memory operands in three independent chains
template <typename T, int n, int sa, int sb> __device__ __forceinline__ T dot3(const T* a, const T* b, T init)
{
    T c0 = init, c1 = T(0), c2 = T(0);
#pragma unroll
    for (int q = 0; q < n; q += 3) {
        c0 += a[q * sa] * b[q * sb];
        if (q + 1 < n) c1 += a[(q + 1) * sa] * b[(q + 1) * sb];
        if (q + 2 < n) c2 += a[(q + 2) * sa] * b[(q + 2) * sb];
    }
    return (c0 + c1) + c2;
}

// =====================================================================================
// per-warp solver state: thin view over the shared-memory block
// =====================================================================================
template <typename T, int N, bool PC = false> struct Solver {
    using L = Layout<T, N, PC>;
    using C = Const<T>;
    T* sm;        // T region
    int* nr;      // live rows per stage
    int lane, mcap, SS;
    const void* rows_g;   // this problem's corridor rows in global memory, [N][mcap][4] = (a0 a1 a2 b)
    bool final_variant, io32 = false;
    T *Z, *DZ, *ZL, *ZU, *G, *Y, *P, *D, *JC, *PHID, *KG, *KFF, *HDR, *S, *LC, *BND;
    T *QINV, *PQQ0, *DZAP;   // predictor-corrector only
    T* fac_out = nullptr;   // when set, riccati_backward streams the factor ([P: N x 91][K | Quu^-1 | J: N x 113]) to HBM

    __device__ __forceinline__ void bind(unsigned char* smem_raw, int lane_, int mcap_)
    {
        sm = reinterpret_cast<T*>(smem_raw + L::HEAD_BYTES);
        nr = reinterpret_cast<int*>(smem_raw + 16);
        lane = lane_; mcap = mcap_;
        SS = L::s_stride(mcap);
        Z = sm + L::Z; ZL = sm + L::ZL; ZU = sm + L::ZU; Y = sm + L::Y; HDR = sm + L::HDR; BND = sm + L::BND;
        S = sm + L::s_off(mcap); LC = sm + L::lc_off(mcap);
        T* sh = sm + L::sh_off(mcap);
        DZ = sh + L::SH_DZ; G = sh + L::SH_G; P = sh + L::SH_P; D = sh + L::SH_D; JC = sh + L::SH_JC; PHID = sh + L::SH_PHID;
        KG = L::KG_OVERLAID ? sm + L::KG : sh + L::SH_KG;
        KFF = sm + L::KFF;
        QINV = sm + L::QINV; PQQ0 = sm + L::PQQ0; DZAP = sh + L::SH_DZAP;
    }
    // registers <-> the slice of R that the sweep-private overlay is about to overwrite
    __device__ __forceinline__ void park(T (&regs)[L::NPARK_LANE]) const
    {
#pragma unroll
        for (int t = 0; t < L::NPARK_LANE; t++) regs[t] = sm[lane + 32 * t];
    }
    __device__ __forceinline__ void unpark(const T (&regs)[L::NPARK_LANE]) const
    {
#pragma unroll
        for (int t = 0; t < L::NPARK_LANE; t++) sm[lane + 32 * t] = regs[t];
    }

    // exchange the parked registers with shared memory in place: after the first call the stage-phase state R is back
    // in shared memory and the sweep-private arrays (gains, Quu^-1, ...) wait in the registers; the second call undoes it
    __device__ __forceinline__ void swap_parked(T (&regs)[L::NPARK_LANE]) const
    {
#pragma unroll
        for (int t = 0; t < L::NPARK_LANE; t++) { const T tmp = sm[lane + 32 * t]; sm[lane + 32 * t] = regs[t]; regs[t] = tmp; }
    }

    __device__ __forceinline__ int live(int k) const { return k == 0 ? 0 : min(nr[k], mcap); }
    // corridor row j of stage k: (a0, a1, a2, b), read-only path (LDG.128, L1-resident across iterations)
    __device__ __forceinline__ void load_row(int k, int j, T (&r)[4]) const
    {
        if constexpr (sizeof(T) == 8) {
            if (io32) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(rows_g) + (size_t)(k * mcap + j) * 4));
                r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w;
                return;
            }
        }
        const T* p = static_cast<const T*>(rows_g) + (size_t)(k * mcap + j) * 4;
        if constexpr (sizeof(T) == 8) {
            const double2 a = __ldg(reinterpret_cast<const double2*>(p)), c = __ldg(reinterpret_cast<const double2*>(p) + 1);
            r[0] = a.x; r[1] = a.y; r[2] = c.x; r[3] = c.y;
        } else {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p));
            r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w;
        }
    }

    // ---------------------------------------------------------------- model evaluation ---
    // At z + a dz ("lanes = stages"): cost, gradient, compact Jacobian, defects, theta and the barrier
    // log-sum.  One code path serves the initial point (a = 0) and every line-search trial; an
    // accepted trial's gradient / Jacobians / defects are exactly what the next iteration needs,
    // so nothing is evaluated twice.  Logs are taken of fixed-size products (5 + ceil(m/8) logs per
    // stage instead of 34 + m).
    __device__ void evaluate(T a, T& f_out, T& th_out, T& ls_out)
    {
        T f = T(0), th = T(0), ls = T(0);
        for (int k = lane; k < N; k += 32) {
            T zk[NZ];
#pragma unroll
            for (int i = 0; i < NZ; i++) zk[i] = Z[k * NZ + i] + a * DZ[k * NZ + i];
            const T* hdr = HDR + k * L::HDR_S;
            f += objective<T, true>(zk, hdr, k == 0, final_variant && k == N - 1, G + k * NZ);
            if (k < N - 1) {
                T c[NXI];
                dynamics<T, true>(zk, hdr + 3, c, JC + k * NJC);
#pragma unroll
                for (int i = 0; i < NXI; i++) {
                    const int zi = (k + 1) * NZ + e_col(i);
                    const T d = c[i] - (Z[zi] + a * DZ[zi]);
                    th += fabs(d);
                    D[k * NXI + i] = d;
                }
            }
            T prod = T(1);
#pragma unroll
            for (int i = 0; i < NZ; i++) {
                T sl = zk[i] - lower_bound<T>(i), su = upper_bound<T>(i) - zk[i];
                if (i >= 8 && k == 0) { sl = T(1); su = T(1); }      // stage-0 states are fixed, not bounded
                prod *= sl * su;
                if (i % Eps<T>::logvars == Eps<T>::logvars - 1 || i == NZ - 1) { ls += log_t(prod); prod = T(1); }
            }
            const int m = live(k);
            for (int j = 0; j < m; j++) {
                T r[4]; load_row(k, j, r);
                T sj = S[k * SS + j];
                T rc = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + sj;
                const T adz = r[0] * DZ[k * NZ + 8] + r[1] * DZ[k * NZ + 9] + r[2] * DZ[k * NZ + 10];
                sj += a * (-rc - adz);                 // trial slack s + a ds
                rc *= (T(1) - a);                      // the constraint is linear: residual shrinks by (1 - a)
                th += fabs(rc);
                prod *= sj;
                if ((j & (Eps<T>::logrows - 1)) == Eps<T>::logrows - 1) { ls += log_t(prod); prod = T(1); }
            }
            ls += log_t(prod);
        }
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
            const T* yn = Y + (k + 1) * NXI;   // only dereferenced for k < N-1
            const T* yk = Y + k * NXI;         // row 0 stays zero
            const T* jc = JC + k * NJC;
            T al0 = T(0), al1 = T(0), al2 = T(0);
            for (int j = 0; j < m; j++) {
                T r[4]; load_row(k, j, r);
                const T sj = S[k * SS + j], lj = LC[k * SS + j];
                al0 += r[0] * lj; al1 += r[1] * lj; al2 += r[2] * lj;
                const T rc = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + sj;
                const T cc = sj * lj;
                cs += cc; cmx = fmax(cmx, cc); cmn = fmin(cmn, cc);
                rin = fmax(rin, fmax(fabs(rc), rc - sj));
            }
            const int nfree = (k == 0) ? 8 : NZ;
#pragma unroll 1
            for (int i = 0; i < nfree; i++) {
                const T zi = Z[k * NZ + i], zl = ZL[k * NZ + i], zu = ZU[k * NZ + i];
                T r = G[k * NZ + i] - zl + zu;
                if (k < N - 1) r += jt_y<T>(jc, yn, i);
                if (i >= 8) r -= yk[i - 8];
                else if (i >= 4) r -= yk[5 + i];
                if (i >= 8 && i < 11) r += (i == 8 ? al0 : (i == 9 ? al1 : al2));
                rs = fmax(rs, fabs(r));
                const T cl = (zi - BND[i]) * zl, cu = (BND[NZ + i] - zi) * zu;
                cs += cl + cu;
                cmx = fmax(cmx, fmax(cl, cu));
                cmn = fmin(cmn, fmin(cl, cu));
            }
            if (k < N - 1) {
#pragma unroll
                for (int i = 0; i < NXI; i++) req = fmax(req, fabs(D[k * NXI + i]));
            }
        }
        rs_n = warp_max(rs); req_n = warp_max(req); rin_n = warp_max(rin); rcomp = warp_max(cmx);
        csum = warp_sum(cs); cmin = warp_min(cmn);
    }

    // ------------------------------------------ barrier-augmented stage Hessian and rhs ---
    // bounds: flat over the N*17 (stage, variable) pairs, all 32 lanes busy; corridor rows: per stage.
    __device__ void assemble(T mu_t)
    {
        for (int e = lane; e < N * NZ; e += 32) {
            const int k = e / NZ, i = e - k * NZ;
            T* phi = PHID + k * L::PHI_S;
            if (e < 8 || e >= NZ) {
                const T zi = Z[e];
                const T isl = rcp_t(zi - BND[i]), isu = rcp_t(BND[NZ + i] - zi);
                phi[i] = cost_hess_diag<T>(i, HDR + k * L::HDR_S, k == 0, final_variant && k == N - 1) + ZL[e] * isl + ZU[e] * isu;
                G[e] += mu_t * (isu - isl);
            } else {
                phi[i] = T(1);
                G[e] = T(0);
            }
        }
        __syncwarp();
        for (int k = lane; k < N; k += 32) {
            T* phi = PHID + k * L::PHI_S;
            T o01 = T(0), o02 = T(0), o12 = T(0), d0 = T(0), d1 = T(0), d2 = T(0), g0 = T(0), g1 = T(0), g2 = T(0);
            const int m = live(k);
            for (int j = 0; j < m; j++) {
                T r[4]; load_row(k, j, r);
                const T sj = S[k * SS + j], lj = LC[k * SS + j], is = rcp_t(sj);
                const T rc = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + sj;
                const T sg = lj * is, tt = (mu_t + lj * rc) * is;
                d0 += r[0] * r[0] * sg; d1 += r[1] * r[1] * sg; d2 += r[2] * r[2] * sg;
                o01 += r[0] * r[1] * sg; o02 += r[0] * r[2] * sg; o12 += r[1] * r[2] * sg;
                g0 += r[0] * tt; g1 += r[1] * tt; g2 += r[2] * tt;
            }
            phi[8] += d0; phi[9] += d1; phi[10] += d2;
            phi[17] = o01; phi[18] = o02; phi[19] = o12;
            phi[20] = T(-2) * HDR[k * L::HDR_S + 8];   // H[u_i][uprev_i]
            if (k > 0) { G[k * NZ + 8] += g0; G[k * NZ + 9] += g1; G[k * NZ + 10] += g2; }
        }
    }

    // ------------- accept the step, measure the new point, assemble the next Newton system (default algorithm) ----
    // What update() + residuals() + assemble() do one after the other, in one sweep over the corridor rows, one over
    // the (stage, variable) pairs and one over the stages (used when PC is off; the predictor-corrector keeps the
    // three separate phases, its affine analysis sits between them):
    //   * the accepted step: s += a ds, lambda += ad dlambda, z_l / z_u += ad d(.), z += a dz, y += a (y_new - y);
    //     mu_p is the barrier target the multiplier steps were computed for;
    //   * the four residual norms of the NEW point and the complementarity sum / max / min;
    //   * the barrier-augmented stage Hessians, and the gradient of the next QP in a form that does not need the next
    //     barrier target yet (it follows from the complementarity sum this pass produces):
    //         g~ = [grad f + A' lambda r_c / s]  +  mu_t [1/s_u - 1/s_l + A'(1/s)]  =  G + mu_t * T,
    //     T parked in the dead step array dz, the rows' partial sums in the dead costate array p; finish_rhs(mu_t)
    //     adds mu_t * T.   a = ad = 0 with dz = 0 is the initial point.
    __device__ void post_step(T mu_p, T a, T ad, T& rs_n, T& req_n, T& rin_n, T& rcomp, T& csum, T& cmin)
    {
        T rs = T(0), req = T(0), rin = T(0), cmx = T(0), cs = T(0), cmn = T(1e30);
        for (int e = NXI + lane; e < N * NXI; e += 32) Y[e] += a * (P[e] - Y[e]);
        __syncwarp();                                          // the costates are dead from here: rows park their sums there
        for (int k = lane; k < N; k += 32) {
            T* phi = PHID + k * L::PHI_S;
            T* tmp = P + k * NXI;                              // d0 d1 d2 | t0 t1 t2 | g0 g1 g2 | al0 al1 al2
            T acc[12];
#pragma unroll
            for (int q = 0; q < 12; q++) acc[q] = T(0);
            T o01 = T(0), o02 = T(0), o12 = T(0);
            const int m = live(k);
            for (int j = 0; j < m; j++) {
                T r[4]; load_row(k, j, r);
                const T so = S[k * SS + j], lo = LC[k * SS + j];
                const T rco = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + so;
                const T ds = -rco - (r[0] * DZ[k * NZ + 8] + r[1] * DZ[k * NZ + 9] + r[2] * DZ[k * NZ + 10]);
                const T sj = so + a * ds, lj = lo + ad * ((mu_p - lo * ds) * rcp_t(so) - lo);
                // residual of the row at the new point, from the very numbers that will be in memory (not (1 - a) * rco: the
                // rounding of z + a dz and s + a ds is part of what the next Newton step has to remove)
                const T rc = r[0] * (Z[k * NZ + 8] + a * DZ[k * NZ + 8]) + r[1] * (Z[k * NZ + 9] + a * DZ[k * NZ + 9]) +
                             r[2] * (Z[k * NZ + 10] + a * DZ[k * NZ + 10]) - (r[3] + C::hu) + sj;
                S[k * SS + j] = sj;
                LC[k * SS + j] = lj;
                const T cc = sj * lj;
                cs += cc; cmx = fmax(cmx, cc); cmn = fmin(cmn, cc);
                rin = fmax(rin, fmax(fabs(rc), rc - sj));
                const T is = rcp_t(sj), sg = lj * is, tt = lj * rc * is;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    acc[c] += r[c] * r[c] * sg; acc[3 + c] += r[c] * is; acc[6 + c] += r[c] * tt; acc[9 + c] += r[c] * lj;
                }
                o01 += r[0] * r[1] * sg; o02 += r[0] * r[2] * sg; o12 += r[1] * r[2] * sg;
            }
#pragma unroll
            for (int q = 0; q < 12; q++) tmp[q] = acc[q];
            phi[17] = o01; phi[18] = o02; phi[19] = o12;
            phi[20] = T(-2) * HDR[k * L::HDR_S + 8];   // H[u_i][uprev_i]
        }
        __syncwarp();
        for (int e = lane; e < N * NZ; e += 32) {
            const int k = e / NZ, i = e - k * NZ;
            T* phi = PHID + k * L::PHI_S;
            if (e < 8 || e >= NZ) {
                const T zo = Z[e], dzi = DZ[e], zlo = ZL[e], zuo = ZU[e];
                const T islo = rcp_t(zo - BND[i]), isuo = rcp_t(BND[NZ + i] - zo);
                const T zl = zlo + ad * ((mu_p - zlo * dzi) * islo - zlo), zu = zuo + ad * ((mu_p + zuo * dzi) * isuo - zuo);
                const T zi = zo + a * dzi;
                Z[e] = zi; ZL[e] = zl; ZU[e] = zu;
                const T sl = zi - BND[i], su = BND[NZ + i] - zi;
                const T cl = sl * zl, cu = su * zu;
                cs += cl + cu;
                cmx = fmax(cmx, fmax(cl, cu));
                cmn = fmin(cmn, fmin(cl, cu));
                const T isl = rcp_t(sl), isu = rcp_t(su);
                T ph = cost_hess_diag<T>(i, HDR + k * L::HDR_S, k == 0, final_variant && k == N - 1) + zl * isl + zu * isu;
                T tc = isu - isl;
                if (i >= 8 && i < 11) { ph += P[k * NXI + i - 8]; tc += P[k * NXI + i - 5]; }
                phi[i] = ph;
                DZ[e] = tc;
            } else {                                           // stage-0 states: fixed by the xinit equality
                phi[i] = T(1);
                DZ[e] = T(0);
            }
        }
        __syncwarp();
        for (int k = lane; k < N; k += 32) {
            const T* tmp = P + k * NXI;
            const T* yn = Y + (k + 1) * NXI;   // only dereferenced for k < N-1
            const T* yk = Y + k * NXI;         // row 0 stays zero
            const T* jc = JC + k * NJC;
            const int nfree = (k == 0) ? 8 : NZ;
#pragma unroll 1
            for (int i = 0; i < nfree; i++) {
                T r = G[k * NZ + i] - ZL[k * NZ + i] + ZU[k * NZ + i];
                if (k < N - 1) r += jt_y<T>(jc, yn, i);
                if (i >= 8) r -= yk[i - 8];
                else if (i >= 4) r -= yk[5 + i];
                if (i >= 8 && i < 11) r += tmp[1 + i];          // A' lambda
                rs = fmax(rs, fabs(r));
            }
            if (k < N - 1) {
#pragma unroll
                for (int i = 0; i < NXI; i++) req = fmax(req, fabs(D[k * NXI + i]));
            }
            if (k > 0) {
                G[k * NZ + 8] += tmp[6]; G[k * NZ + 9] += tmp[7]; G[k * NZ + 10] += tmp[8];
            } else {
#pragma unroll
                for (int i = 8; i < NZ; i++) G[i] = T(0);
            }
        }
        rs_n = warp_max(rs); req_n = warp_max(req); rin_n = warp_max(rin); rcomp = warp_max(cmx);
        csum = warp_sum(cs); cmin = warp_min(cmn);
    }
    // gradient of the QP: G + mu_t * T
    __device__ void finish_rhs(T mu_t)
    {
        for (int e = lane; e < N * NZ; e += 32) G[e] += mu_t * DZ[e];
    }

    // ------------------------------------------------------------- Riccati backward -----
    // Four warp-synchronous phases per stage:
    //   A: PF = P+(:,x) F,  tv = p+ + P+ d           B: Q blocks (F' PF + coupling), q~ vectors
    //   C: 4x4 Cholesky (rsqrt, no divisions), Y = L^-1 [Q_ux Q_uq | q_u], K = -L^-T Y
    //   D: P_k = blkdiag(Q_xx, Phi_qq) - Y'Y, p_k = q_xi - Y' y0
    // F = d x+ / d(u, x) is never formed: its 60 non-zeros are the 51 compact Jacobian words plus the
    // constants 1 and h, in a fixed pattern
    //     pos+ = pos + Jpv vel + Jpr rpy + JpT T,   vel+ = Jvv vel + Jvr rpy + JvT T + Jvw w,   rpy+ = rpy + h w
    // so one row of P+ times F (phase A, lane = row) and F' times one column of PF (phase B, lane = column;
    // the vector tv rides along as a 14th column) are the same 54 multiply-adds on nine register operands,
    // with every Jacobian word a broadcast load: half the flops of the dense products, a quarter of their
    // shared-memory wavefronts, and thirteen independent accumulators per lane instead of three.
    // Phases C and D: every lane owns a fixed set of matrix entries (compile-time trip counts, per-lane
    // offsets hoisted out of the stage loop).  Returns false on a non-positive pivot.
    //
    // out[0..12] (v-ordering w0 w1 w2 T p0 p1 p2 v0 v1 v2 r0 r1 r2) = structured product of the 9-vector
    // (xp, xv, xa) = (pos, vel, rpy parts) with F: out = F' x.
    __device__ __forceinline__ static void ft_times(const T* __restrict__ jc, const T (&xp)[3], const T (&xv)[3], const T (&xa)[3], T (&out)[13])
    {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            out[c] = (jc[JVW + c] * xv[0] + jc[JVW + 3 + c] * xv[1]) + (jc[JVW + 6 + c] * xv[2] + C::h * xa[c]);
            out[4 + c] = xp[c];
            out[7 + c] = ((jc[JPV + c] * xp[0] + jc[JPV + 3 + c] * xp[1]) + jc[JPV + 6 + c] * xp[2]) +
                         ((jc[JVV + c] * xv[0] + jc[JVV + 3 + c] * xv[1]) + jc[JVV + 6 + c] * xv[2]);
            out[10 + c] = (((jc[JPR + c] * xp[0] + jc[JPR + 3 + c] * xp[1]) + jc[JPR + 6 + c] * xp[2]) +
                           ((jc[JVR + c] * xv[0] + jc[JVR + 3 + c] * xv[1]) + jc[JVR + 6 + c] * xv[2])) + xa[c];
        }
        out[3] = ((jc[JPT] * xp[0] + jc[JPT + 1] * xp[1]) + jc[JPT + 2] * xp[2]) +
                 ((jc[JVT] * xv[0] + jc[JVT + 1] * xv[1]) + jc[JVT + 2] * xv[2]);
    }

    __device__ bool riccati_backward()
    {
        T* PN = sm + L::PN; T* PF = sm + L::PF; T* GG = sm + L::GG;
        T* TV = sm + L::TV; T* QUU = sm + L::QUU; T* QUR = sm + L::QUR;
        T* QV = sm + L::QV; T* QXI = sm + L::QXI; T* YS = sm + L::YS; T* Y0 = sm + L::Y0;
        bool ok = true;
        // lower-triangle entries (i >= j) of a 13x13 owned by this lane: e = lane, lane+32, lane+64 (< 91)
        int ti[3], tj[3], goff[3], poff[3];
#pragma unroll
        for (int t = 0; t < 3; t++) {
            const int e = lane + 32 * t;
            int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);      // row of packed-lower entry e
            i += ((i + 1) * (i + 2) / 2 <= e) - (i * (i + 1) / 2 > e);           // guard the rounding of the root
            const int j = e - i * (i + 1) / 2;
            ti[t] = i; tj[t] = j;
            // phase D (xi-ordering): where the additive terms of P_k[i][j] live (-1: none)
            goff[t] = (i < 9) ? (4 + i) * 13 + 4 + j : -1;                                   // Q_xx part in GG
            poff[t] = (i < 9) ? (i == j ? 8 + i : (i < 3 ? 17 + i + j - 1 : -1))             // Phi_xx
                              : (i == j ? 4 + i - 9 : -1);                                   // Phi_qq
        }
        for (int k = N - 1; k >= 0; k--) {
            const bool nx = (k < N - 1);
            const T* phi = PHID + k * L::PHI_S;
            const T* gk = G + k * NZ;
            if (nx) {
                // single precision: the 51 Jacobian words of the stage are read once into registers and serve both
                // products (phases A and B); double precision has no registers to spare (122 hold the parked state)
                T jreg[sizeof(T) == 4 ? NJC : 1];
                if constexpr (sizeof(T) == 4) {
#pragma unroll
                    for (int e = 0; e < NJC; e++) jreg[e] = JC[k * NJC + e];
                }
                const T* jc = sizeof(T) == 4 ? jreg : JC + k * NJC;
                // ---- phase A: lane = row of P+ (xi-ordering) ----------------------------------
                if (lane < NXI) {
                    const T* pr = PN + lane * 13;
                    const T xp[3] = {pr[0], pr[1], pr[2]}, xv[3] = {pr[3], pr[4], pr[5]}, xa[3] = {pr[6], pr[7], pr[8]};
                    T out[13];
                    ft_times(jc, xp, xv, xa, out);
                    T* pf = PF + lane * 13;
#pragma unroll
                    for (int c = 0; c < 13; c++) pf[c] = out[c];
                    const T* dk = D + k * NXI;
                    const T c0 = (((P[(k + 1) * NXI + lane] + xp[0] * dk[0]) + xv[0] * dk[3]) + xa[0] * dk[6]) + (pr[9] * dk[9] + pr[12] * dk[12]);
                    const T c1 = ((xp[1] * dk[1] + xv[1] * dk[4]) + xa[1] * dk[7]) + pr[10] * dk[10];
                    const T c2 = ((xp[2] * dk[2] + xv[2] * dk[5]) + xa[2] * dk[8]) + pr[11] * dk[11];
                    TV[lane] = (c0 + c1) + c2;
                }
                __syncwarp();
                // ---- phase B: lane = column of PF (v-ordering); lane 13 = the vector tv ----------
                // g = F' (column), then ONE instruction stream for all fourteen lanes turns it into what phases C / D read:
                //   lanes 0-3  (u columns)  Q_uu column (+ P+_qq + PF_qu + PF_qu' + Phi_uu) and the Q_ux row
                //   lanes 4-12 (x columns)  column of Q_xx
                //   lane 13    (vector tv)  q~_u = g + g_k,u + tv_q and q~_x = g + g_k,x
                // the three groups differ only in per-lane base pointers / strides and a few predicated loads (no
                // divergent branches: they used to run one after the other)
                if (lane < 14) {
                    const bool isU = lane < 4, isV = lane == 13, isX = !isU && !isV;
                    const int j = lane & 3;
                    const T* col = lane < 13 ? PF + lane : TV;
                    const int cs = lane < 13 ? 13 : 1;
                    const T xp[3] = {col[0], col[cs], col[2 * cs]}, xv[3] = {col[3 * cs], col[4 * cs], col[5 * cs]};
                    const T xa[3] = {col[6 * cs], col[7 * cs], col[8 * cs]};
                    T g[13];
                    ft_times(jc, xp, xv, xa, g);
                    const T* pA = isV ? gk : PN + (9 * 13 + 9 + j);
                    const T* pB = isV ? TV + 9 : PF + (9 * 13 + j);
                    const int sAB = isV ? 1 : 13;
                    const T* pC = PF + (9 + j) * 13;
                    if (!isX) {
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            T acc = g[i] + (pA[i * sAB] + (pB[i * sAB] + (isU ? pC[i] : T(0))));
                            if (isU && i == j) acc += phi[i];
                            g[i] = acc;
                        }
                        const T* pE = isU ? pC : gk + 4;
#pragma unroll
                        for (int i = 4; i < 13; i++) g[i] += pE[i];
                    }
                    if (PC) __syncwarp(0x3fffu);                   // GG overwrites PF: every read of PF is behind us
                    if (!isX) {
                        T* dT = isV ? QV : QUU + j;
                        const int sT = isV ? 1 : 4;
#pragma unroll
                        for (int i = 0; i < 4; i++) dT[i * sT] = g[i];
                    }
                    T* dB = isU ? QUR + (j * 13 - 4) : (isV ? QXI - 4 : GG + lane);
                    const int sD = isX ? 13 : 1;
#pragma unroll
                    for (int i = 4; i < 13; i++) dB[i * sD] = g[i];
                }
                if (lane < 16) QUR[(lane >> 2) * 13 + 9 + (lane & 3)] = ((lane >> 2) == (lane & 3)) ? phi[20] : T(0);
            } else {
                // terminal stage: no dynamics behind it
                if (lane < 16) {
                    QUU[lane] = ((lane >> 2) == (lane & 3)) ? phi[lane >> 2] : T(0);
                    QUR[(lane >> 2) * 13 + 9 + (lane & 3)] = ((lane >> 2) == (lane & 3)) ? phi[20] : T(0);
                }
                for (int e = lane; e < 36; e += 32) QUR[(e / 9) * 13 + e % 9] = T(0);
                if (lane < 4) QV[lane] = gk[lane];
                if (lane < 13) QXI[lane] = (lane < 9) ? gk[8 + lane] : gk[lane - 5];
            }
            __syncwarp();
            // ---- phase C: factor the pivot block redundantly in registers, solve 13 + 1 columns ----
            T l[10], li[4];
            {
                T a[16];
                a[0] = QUU[0]; a[4] = QUU[4]; a[5] = QUU[5]; a[8] = QUU[8]; a[9] = QUU[9]; a[10] = QUU[10];
                a[12] = QUU[12]; a[13] = QUU[13]; a[14] = QUU[14]; a[15] = QUU[15];
                ok &= chol4<T>(a, l, li);
            }
            if (lane < (PC ? 18 : 14)) {
                T x[4];
#pragma unroll
                for (int r = 0; r < 4; r++) x[r] = (lane < 13) ? QUR[r * 13 + lane] : (lane == 13 ? QV[r] : (r == lane - 14 ? T(1) : T(0)));
                fsub4<T>(l, li, x);
                if (lane < 14) {
                    T* ys = (lane < 13) ? YS + lane : Y0;
                    const int ystr = (lane < 13) ? 13 : 1;
#pragma unroll
                    for (int r = 0; r < 4; r++) ys[r * ystr] = x[r];
                }
                bsub4<T>(l, li, x);
                if (lane < 14) {
                    T* kg = (lane < 13) ? KG + k * 52 + lane : KFF + k * 4;
                    const int ystr = (lane < 13) ? 13 : 1;
#pragma unroll
                    for (int r = 0; r < 4; r++) kg[r * ystr] = -x[r];
                } else {                                           // PC: column lane - 14 of Quu^-1, packed lower
                    const int c = lane - 14;
#pragma unroll
                    for (int r = 0; r < 4; r++)
                        if (r >= c) QINV[k * 10 + r * (r + 1) / 2 + c] = x[r];
                }
            }
            __syncwarp();
            // ---- phase D ------------------------------------------------------------------------
#pragma unroll
            for (int t = 0; t < 3; t++) {
                if (lane + 32 * t < 91) {
                    const int i = ti[t], j = tj[t];                // xi-ordering (x0..8, q0..3), i >= j
                    T v = (nx && goff[t] >= 0) ? GG[goff[t]] : T(0);
                    if (poff[t] >= 0) v += phi[poff[t]];
                    v -= (YS[i] * YS[j] + YS[13 + i] * YS[13 + j]) + (YS[26 + i] * YS[26 + j] + YS[39 + i] * YS[39 + j]);
                    PN[i * 13 + j] = v;
                    PN[j * 13 + i] = v;
                }
            }
            if (lane < NXI) {                                      // p_k, one entry per lane
                const int i = lane;
                const T qxi = (i < 9) ? QXI[i] : gk[i - 5];        // the u_prev rows of q~ are the stage gradient itself
                P[k * NXI + i] = qxi - ((YS[i] * Y0[0] + YS[13 + i] * Y0[1]) + (YS[26 + i] * Y0[2] + YS[39 + i] * Y0[3]));
            }
            __syncwarp();
            if (fac_out) {   // factor of stage k: P_k (91 words, PSYM layout) -> P region; [K_k 52 | Quu^-1 packed lower 10 | J_k 51] -> KQJ region
                T* fp = fac_out + (size_t)k * 91;
                T* fk = fac_out + (size_t)N * 91 + (size_t)k * (FAC_WORDS - 91);
#pragma unroll
                for (int t = 0; t < 3; t++)
                    if (lane + 32 * t < 91) fp[PSYM[ti[t]][tj[t]]] = PN[ti[t] * 13 + tj[t]];
                for (int e = lane; e < 52; e += 32) fk[e] = KG[k * 52 + e];
                if (lane < 4) {   // column `lane` of Quu^-1 via the Cholesky factor
                    T x[4];
#pragma unroll
                    for (int r = 0; r < 4; r++) x[r] = (r == lane) ? T(1) : T(0);
                    fsub4<T>(l, li, x);
                    bsub4<T>(l, li, x);
#pragma unroll
                    for (int r = 0; r < 4; r++)
                        if (r >= lane) fk[52 + r * (r + 1) / 2 + lane] = x[r];
                }
                for (int e = lane; e < NJC; e += 32) fk[62 + e] = nx ? JC[k * NJC + e] : T(0);
            }
        }
        return ok;
    }

    // ------------------------------------------------------------- forward rollout ------
    // dz_k = (du, dq, dx): du = K dxi + kff (8 lanes per row + shuffles); dxi+ = [F (du, dx) + d_x ; du + d_q]
    // with the rows of F addressed through per-lane offsets into the compact Jacobian (branch-free).
    // DELTA (predictor-corrector): the same rollout for the correction of the step -- zero defects, feed-forward terms
    // and p_0 from delta_backward(), the q-block of P_0 saved by the first rollout, dz accumulated
    template <bool DELTA = false> __device__ bool rollout()
    {
        T* PN = sm + L::PN; T* DXI = sm + L::DXI; T* TV = sm + L::TV;
        bool ok = true;
        {   // stage 0: x fixed (dx = 0), u_prev free: dq = -Pqq^-1 p_q
            T a[16], l[10], li[4], x[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
#pragma unroll
                for (int c = 0; c <= r; c++) a[4 * r + c] = DELTA ? PQQ0[r * (r + 1) / 2 + c] : PN[(9 + r) * 13 + 9 + c];
                x[r] = DELTA ? -TV[9 + r] : -P[9 + r];
            }
            if (PC && !DELTA && lane == 0) {
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c <= r; c++) PQQ0[r * (r + 1) / 2 + c] = a[4 * r + c];
            }
            ok = chol4<T>(a, l, li);
            fsub4<T>(l, li, x);
            bsub4<T>(l, li, x);
            if (lane < 13) DXI[lane] = (lane < 9) ? T(0) : (lane == 9 ? x[0] : (lane == 10 ? x[1] : (lane == 11 ? x[2] : x[3])));
        }
        // per-lane row description of dxi+ (lane = row in xi-ordering)
        const int rt = lane < 3 ? 0 : (lane < 6 ? 1 : (lane < 9 ? 2 : (lane < 13 ? 3 : 4)));   // pos+, vel+, rpy+, q+, idle
        const int rr = lane < 3 ? lane : (lane < 6 ? lane - 3 : 0);
        const int offT = rt == 0 ? JPT + rr : (rt == 1 ? JVT + rr : 0);
        const int offV = rt == 0 ? JPV + 3 * rr : (rt == 1 ? JVV + 3 * rr : 0);
        const int offR = rt == 0 ? JPR + 3 * rr : (rt == 1 ? JVR + 3 * rr : 0);
        const int offW = rt == 1 ? JVW + 3 * rr : 0;
        const T mMain = rt <= 1 ? T(1) : T(0), mW = rt == 1 ? T(1) : T(0);
        const T mSelf = (rt == 0 || rt == 2) ? T(1) : T(0);
        const T cDu = rt == 2 ? C::h : (rt == 3 ? T(1) : T(0));
        const int duSrc = 8 * ((rt == 2 ? lane - 6 : lane - 9) & 3);
        const int lrow = lane < 13 ? lane : 0;
        const int r4 = lane >> 3, part = lane & 7;
        __syncwarp();
        for (int k = 0; k < N; k++) {
            const T* kg = KG + k * 52 + r4 * 13;
            T acc = kg[part] * DXI[part];
            if (part < 5) acc += kg[part + 8] * DXI[part + 8];
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += KFF[k * 4 + r4];
            const T dw0 = __shfl_sync(0xffffffffu, acc, 0), dw1 = __shfl_sync(0xffffffffu, acc, 8);
            const T dw2 = __shfl_sync(0xffffffffu, acc, 16), dT = __shfl_sync(0xffffffffu, acc, 24);
            const T du_l = __shfl_sync(0xffffffffu, acc, 8 * (lane & 3));   // du[lane & 3]
            const T du_s = __shfl_sync(0xffffffffu, acc, duSrc);            // du feeding row `lane` of dxi+
            const T self = DXI[lrow];
            if (lane < NZ) {
                const T v = (lane < 4) ? du_l : (lane < 8 ? DXI[5 + lane] : DXI[lane - 8]);
                if (DELTA) DZ[k * NZ + lane] += v; else DZ[k * NZ + lane] = v;
            }
            T nxt = T(0);
            if (k < N - 1) {
                const T* jc = JC + k * NJC;
                const T m0 = jc[offT] * dT + jc[offV] * DXI[3] + jc[offR] * DXI[6];
                const T m1 = jc[offV + 1] * DXI[4] + jc[offR + 1] * DXI[7];
                const T m2 = jc[offV + 2] * DXI[5] + jc[offR + 2] * DXI[8];
                const T mw = jc[offW] * dw0 + jc[offW + 1] * dw1 + jc[offW + 2] * dw2;
                nxt = (DELTA ? T(0) : D[k * NXI + lrow]) + mSelf * self + cDu * du_s + mMain * ((m0 + m1) + m2) + mW * mw;
            }
            __syncwarp();
            if (k < N - 1 && lane < 13) DXI[lane] = nxt;
            __syncwarp();
        }
        return ok;
    }

    // ------------------------------------ predictor-corrector: correction of the step -----
    // The corrector changes only the gradient of the QP (dg, kept in the p | d area while the sweeps do not need it);
    // defects and Hessian are those of the affine solve, so the change of the step obeys the homogeneous recursion
    //   dq = dg_k + J_k' dp_{k+1},  dkff_k = -Quu_k^-1 dq_u,  dp_k = dq_xi + K_k' dq_u
    // with the gains and Quu^-1 stored by riccati_backward(): one 17-lane vector sweep, no matrix work.
    __device__ void delta_backward()
    {
        T* TV = sm + L::TV;
        const T* DG = P;                                           // [N][17], contiguous over the p | d area
        const int zi = lane < NZ ? lane : 0;
        const int zt = zi < 3 ? 0 : (zi == 3 ? 1 : (zi < 8 ? 2 : (zi < 11 ? 3 : (zi < 14 ? 4 : 5))));   // rate, T, uprev, pos, vel, rpy
        const int zj = zt == 0 ? zi : (zt == 3 ? zi - 8 : (zt == 4 ? zi - 11 : (zt == 5 ? zi - 14 : 0)));
        const int oA = zt == 1 ? JPT : (zt == 4 ? JPV + zj : (zt == 5 ? JPR + zj : 0));
        const int sA = zt == 1 ? 1 : 3;
        const T mA = (zt == 1 || zt == 4 || zt == 5) ? T(1) : T(0);
        const int oB = zt == 0 ? JVW + zj : (zt == 1 ? JVT : (zt == 4 ? JVV + zj : (zt == 5 ? JVR + zj : 0)));
        const int sB = zt == 1 ? 1 : 3;
        const T mB = (zt == 0 || zt == 1 || zt == 4 || zt == 5) ? T(1) : T(0);
        const T c1 = zt == 0 ? C::h : ((zt == 3 || zt == 5) ? T(1) : T(0));
        const int i1 = zt == 0 ? 6 + zj : (zt == 3 ? zj : (zt == 5 ? 6 + zj : 0));
        const T c2 = (zt == 0 || zt == 1) ? T(1) : T(0);
        const int i2 = zt == 0 ? 9 + zj : 12;
        const int xi = lane < NXI ? lane : 0;
        const int xz = e_col(xi);
        const bool qrow = lane >= 16 && lane < 20;
        int uoff[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int r = lane - 16;
            uoff[c] = qrow ? (r >= c ? r * (r + 1) / 2 + c : c * (c + 1) / 2 + r) : 13 * c + xi;
        }
        T pnext = T(0);
        for (int k = N - 1; k >= 0; k--) {
            T qz = DG[k * NZ + zi];
            if (k < N - 1) {
                const T* jc = JC + k * NJC;
                if (lane < NXI) TV[lane] = pnext;
                __syncwarp();
                const T sa = (jc[oA] * TV[0] + jc[oA + sA] * TV[1]) + jc[oA + 2 * sA] * TV[2];
                const T sb = (jc[oB] * TV[3] + jc[oB + sB] * TV[4]) + jc[oB + 2 * sB] * TV[5];
                qz = (mA * sa + (c1 * TV[i1] + qz)) + (mB * sb + c2 * TV[i2]);
            }
            const T qu0 = __shfl_sync(0xffffffffu, qz, 0), qu1 = __shfl_sync(0xffffffffu, qz, 1);
            const T qu2 = __shfl_sync(0xffffffffu, qz, 2), qu3 = __shfl_sync(0xffffffffu, qz, 3);
            const T qxi = __shfl_sync(0xffffffffu, qz, xz);
            const T* m = qrow ? QINV + k * 10 : KG + k * 52;
            const T u4 = (m[uoff[0]] * qu0 + m[uoff[1]] * qu1) + (m[uoff[2]] * qu2 + m[uoff[3]] * qu3);
            pnext = qxi + u4;
            if (qrow) KFF[k * 4 + lane - 16] = -u4;
            __syncwarp();
        }
        if (lane < NXI) TV[lane] = pnext;                          // dp_0 for the stage-0 solve of rollout<true>()
        __syncwarp();
    }

    // Affine analysis and corrector right-hand side (R in shared memory, dz = affine step).  Returns the barrier
    // target mu_t = max((mu_aff / mu)^3 mu, mu_floor); leaves dg in the p | d area, adds it to the gradient, keeps the
    // affine step for the multiplier steps that follow (dza: this lane's flat slice; DZAP: position parts per stage).
    __device__ T predictor_corrector_rhs(T mu, T csum, int ncomp, T mu_floor, T (&dza)[L::NT])
    {
        T* DG = P;
        T pn = T(0), pd = T(1), dn = T(0), dd = T(1), S1 = T(0), S2 = T(0), S3 = T(0);
        int t = 0;
        for (int e = lane; e < N * NZ; e += 32, t++) {
            const T dzi = DZ[e];
            dza[t] = dzi;
            if (!(e < 8 || e >= NZ)) continue;
            const int i = e % NZ;
            const T zi = Z[e], zl = ZL[e], zu = ZU[e];
            const T sl = zi - BND[i], su = BND[NZ + i] - zi;
            frac_max(pn, pd, -dzi, sl);
            frac_max(pn, pd, dzi, su);
            frac_max(dn, dd, zl * (sl + dzi), sl * zl);
            frac_max(dn, dd, zu * (su - dzi), su * zu);
            const T sdl = -zl * (dzi + sl), sdu = zu * (dzi - su);          // s_l dz_l^aff, s_u dz_u^aff
            S1 += dzi * (zl - zu); S2 += sdl + sdu; S3 += dzi * (sdl / sl - sdu / su);
        }
        for (int k = lane; k < N; k += 32) {
            const int m = live(k);
            DZAP[k * 3] = DZ[k * NZ + 8]; DZAP[k * 3 + 1] = DZ[k * NZ + 9]; DZAP[k * 3 + 2] = DZ[k * NZ + 10];
            for (int j = 0; j < m; j++) {
                T r[4]; load_row(k, j, r);
                const T sj = S[k * SS + j], lj = LC[k * SS + j];
                const T rc = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + sj;
                const T ds = -rc - (r[0] * DZ[k * NZ + 8] + r[1] * DZ[k * NZ + 9] + r[2] * DZ[k * NZ + 10]);
                frac_max(pn, pd, -ds, sj);
                frac_max(dn, dd, lj * (sj + ds), sj * lj);
                const T sdl = -lj * (ds + sj);
                S1 += ds * lj; S2 += sdl; S3 += ds * sdl / sj;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const T n2 = __shfl_xor_sync(0xffffffffu, pn, o), d2 = __shfl_xor_sync(0xffffffffu, pd, o);
            frac_max(pn, pd, n2, d2);
            const T n3 = __shfl_xor_sync(0xffffffffu, dn, o), d3 = __shfl_xor_sync(0xffffffffu, dd, o);
            frac_max(dn, dd, n3, d3);
        }
        S1 = warp_sum(S1); S2 = warp_sum(S2); S3 = warp_sum(S3);
        const T ap = (pn > T(0)) ? fmin(T(1), pd / pn) : T(1), ad = (dn > T(0)) ? fmin(T(1), dd / dn) : T(1);
        const T mu_aff = (csum + ap * S1 + ad * S2 + ap * ad * S3) / (T)ncomp, sg = mu_aff / mu;
        const T mu_t = fmax(sg * sg * sg * mu, mu_floor);
        // corrector: dg = -(mu_t - c_l) / s_l + (mu_t - c_u) / s_u,  c = ds^aff dlambda^aff
        t = 0;
        for (int e = lane; e < N * NZ; e += 32, t++) {
            T dg = T(0);
            if (e < 8 || e >= NZ) {
                const int i = e % NZ;
                const T zi = Z[e], zl = ZL[e], zu = ZU[e], dzi = dza[t];
                const T isl = rcp_t(zi - BND[i]), isu = rcp_t(BND[NZ + i] - zi);
                const T cl = -zl * dzi * (dzi * isl + T(1)), cu = -zu * dzi * (dzi * isu - T(1));
                dg = (mu_t - cu) * isu - (mu_t - cl) * isl;
            }
            DG[e] = dg;
            G[e] += dg;
        }
        __syncwarp();
        for (int k = lane; k < N; k += 32) {
            const int m = live(k);
            T g0 = T(0), g1 = T(0), g2 = T(0);
            for (int j = 0; j < m; j++) {
                T r[4]; load_row(k, j, r);
                const T sj = S[k * SS + j], lj = LC[k * SS + j], is = rcp_t(sj);
                const T rc = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + sj;
                const T ds = -rc - (r[0] * DZAP[k * 3] + r[1] * DZAP[k * 3 + 1] + r[2] * DZAP[k * 3 + 2]);
                const T cr = -lj * ds * (ds * is + T(1));
                const T tt = (mu_t - cr) * is;
                g0 += r[0] * tt; g1 += r[1] * tt; g2 += r[2] * tt;
            }
            if (k > 0) {
                DG[k * NZ + 8] += g0; DG[k * NZ + 9] += g1; DG[k * NZ + 10] += g2;
                G[k * NZ + 8] += g0; G[k * NZ + 9] += g1; G[k * NZ + 10] += g2;
            }
        }
        return mu_t;
    }

    // ------------------------------------ costates of the QP (new equality multipliers) ---
    // y_k = [Phi_k dz_k + g~_k + J_k' y_{k+1}]_xi , stored over p_k (c-ordering [x; q]); the column of
    // J' each lane needs is addressed through per-lane offsets into the compact Jacobian.
    __device__ void costates()
    {
        const int i = lane < 13 ? lane : 0;                       // xi index handled by this lane
        const int zi = e_col(i);                                   // its z index
        const int ct = i < 3 ? 0 : (i < 6 ? 1 : (i < 9 ? 2 : 3));  // pos, vel, rpy, q
        const int jj = ct == 1 ? i - 3 : (ct == 2 ? i - 6 : 0);
        const int offP = ct == 1 ? JPV + jj : (ct == 2 ? JPR + jj : 0);    // column jj of d pos+/d(.)
        const int offV = ct == 1 ? JVV + jj : (ct == 2 ? JVR + jj : 0);    // column jj of d vel+/d(.)
        const T mJ = (ct == 1 || ct == 2) ? T(1) : T(0), mSelf = (ct == 0 || ct == 2) ? T(1) : T(0);
        // off-diagonal position-block terms: (phi slot, dz index) pairs
        const int pa = i == 2 ? 18 : 17, da = i == 0 ? 9 : 8, pb = i == 0 ? 18 : 19, db = i == 2 ? 9 : 10;
        const T mP = ct == 0 ? T(1) : T(0), mQ = ct == 3 ? T(1) : T(0);
        const int cq = ct == 3 ? i - 9 : 0;
        for (int k = N - 1; k >= 1; k--) {
            const T* phi = PHID + k * L::PHI_S;
            const T* dz = DZ + k * NZ;
            T v = G[k * NZ + zi] + phi[zi] * dz[zi] + mP * (phi[pa] * dz[da] + phi[pb] * dz[db]) + mQ * phi[20] * dz[cq];
            if (k < N - 1) {
                const T* jc = JC + k * NJC;
                const T* yn = P + (k + 1) * NXI;
                const T s0 = jc[offP] * yn[0] + jc[offV] * yn[3];
                const T s1 = jc[offP + 3] * yn[1] + jc[offV + 3] * yn[4];
                const T s2 = jc[offP + 6] * yn[2] + jc[offV + 6] * yn[5];
                v += mJ * ((s0 + s1) + s2) + mSelf * yn[i];
            }
            __syncwarp();
            if (lane < 13) P[k * NXI + lane] = v;
            __syncwarp();
        }
    }

    // ------------------------------------------- multiplier steps, fraction to boundary ---
    // The largest admissible steps are tau / max_i(ratio_i) with ratio_i = -d(slack)/slack resp.
    // -d(mult)/mult.  The maxima are tracked as fractions (num, den > 0) and compared by
    // cross-multiplication, so the whole phase needs two divisions instead of four per variable:
    //   -dz_l/z_l = (z_l (s_l + dz) - mu_t) / (s_l z_l),   -dz_u/z_u = (z_u (s_u - dz) - mu_t) / (s_u z_u).
    __device__ __forceinline__ static void frac_max(T& bn, T& bd, T n, T d)
    {
        if (n * bd > bn * d) { bn = n; bd = d; }
    }
    // With the predictor-corrector the multiplier steps carry the second-order terms c = ds^aff dlambda^aff
    // (dz_l = (mu_t - c_l - z_l dz) / s_l - z_l, ...), rebuilt from the affine step (dza, DZAP); the ratios stay
    // division-free after scaling numerator and denominator by the slack.
    __device__ void step_lengths(T mu_t, T tau, T& ap_out, T& ad_out, const T* dza = nullptr)
    {
        T pn = T(0), pd = T(1), dn = T(0), dd = T(1);
        int t = 0;
        for (int e = lane; e < N * NZ; e += 32, t++) {
            if (!(e < 8 || e >= NZ)) continue;
            const int i = e % NZ;
            const T zi = Z[e], dzi = DZ[e], zl = ZL[e], zu = ZU[e];
            const T sl = zi - BND[i], su = BND[NZ + i] - zi;
            frac_max(pn, pd, -dzi, sl);
            frac_max(pn, pd, dzi, su);
            if (PC) {
                const T da = dza[t];
                frac_max(dn, dd, sl * (zl * (sl + dzi) - mu_t) - zl * da * (da + sl), sl * sl * zl);
                frac_max(dn, dd, su * (zu * (su - dzi) - mu_t) - zu * da * (da - su), su * su * zu);
            } else {
                frac_max(dn, dd, zl * (sl + dzi) - mu_t, sl * zl);
                frac_max(dn, dd, zu * (su - dzi) - mu_t, su * zu);
            }
        }
        for (int k = lane; k < N; k += 32) {
            const int m = live(k);
            for (int j = 0; j < m; j++) {
                T r[4]; load_row(k, j, r);
                const T sj = S[k * SS + j], lj = LC[k * SS + j];
                const T rc = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + sj;
                const T ds = -rc - (r[0] * DZ[k * NZ + 8] + r[1] * DZ[k * NZ + 9] + r[2] * DZ[k * NZ + 10]);
                frac_max(pn, pd, -ds, sj);
                if (PC) {
                    const T dsa = -rc - (r[0] * DZAP[k * 3] + r[1] * DZAP[k * 3 + 1] + r[2] * DZAP[k * 3 + 2]);
                    frac_max(dn, dd, sj * (lj * (sj + ds) - mu_t) - lj * dsa * (dsa + sj), sj * sj * lj);
                } else {
                    frac_max(dn, dd, lj * (sj + ds) - mu_t, sj * lj);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const T n2 = __shfl_xor_sync(0xffffffffu, pn, o), d2 = __shfl_xor_sync(0xffffffffu, pd, o);
            frac_max(pn, pd, n2, d2);
            const T n3 = __shfl_xor_sync(0xffffffffu, dn, o), d3 = __shfl_xor_sync(0xffffffffu, dd, o);
            frac_max(dn, dd, n3, d3);
        }
        ap_out = (pn > T(0)) ? fmin(T(1), tau * pd / pn) : T(1);
        ad_out = (dn > T(0)) ? fmin(T(1), tau * dd / dn) : T(1);
    }

    // --------------------------------------------------------------- accept the step ----
    __device__ void update(T mu_t, T a, T ad, const T* dza = nullptr)
    {
        for (int k = lane; k < N; k += 32) {               // corridor rows first: they read the old position
            const int m = live(k);
            for (int j = 0; j < m; j++) {
                T r[4]; load_row(k, j, r);
                const T sj = S[k * SS + j], lj = LC[k * SS + j];
                const T rc = r[0] * Z[k * NZ + 8] + r[1] * Z[k * NZ + 9] + r[2] * Z[k * NZ + 10] - (r[3] + C::hu) + sj;
                const T ds = -rc - (r[0] * DZ[k * NZ + 8] + r[1] * DZ[k * NZ + 9] + r[2] * DZ[k * NZ + 10]);
                T dl;
                if (PC) {
                    const T is = rcp_t(sj);
                    const T dsa = -rc - (r[0] * DZAP[k * 3] + r[1] * DZAP[k * 3 + 1] + r[2] * DZAP[k * 3 + 2]);
                    const T cr = -lj * dsa * (dsa * is + T(1));
                    dl = (mu_t - cr - lj * ds) * is - lj;
                } else {
                    dl = (mu_t - lj * ds) * rcp_t(sj) - lj;
                }
                S[k * SS + j] = sj + a * ds;
                LC[k * SS + j] = lj + ad * dl;
            }
        }
        __syncwarp();
        int t = 0;
        for (int e = lane; e < N * NZ; e += 32, t++) {
            const T zi = Z[e], dzi = DZ[e];
            if (e < 8 || e >= NZ) {
                const int i = e % NZ;
                const T zl = ZL[e], zu = ZU[e];
                const T isl = rcp_t(zi - BND[i]), isu = rcp_t(BND[NZ + i] - zi);
                T cl = T(0), cu = T(0);
                if (PC) {
                    const T da = dza[t];
                    cl = -zl * da * (da * isl + T(1)); cu = -zu * da * (da * isu - T(1));
                }
                ZL[e] = zl + ad * ((mu_t - cl - zl * dzi) * isl - zl);
                ZU[e] = zu + ad * ((mu_t - cu + zu * dzi) * isu - zu);
            }
            Z[e] = zi + a * dzi;
        }
        for (int e = NXI + lane; e < N * NXI; e += 32) Y[e] += a * (P[e] - Y[e]);
    }
};

// =====================================================================================
// the kernel: grid = B CTAs of one warp; dynamic smem = Layout::bytes(mcap)
// =====================================================================================
template <typename T, int N, bool PC = false>
__global__ void __launch_bounds__(32) nmpc_ipm_kernel(const Params<T> prm)
{
    using L = Layout<T, N, PC>;
    using C = Const<T>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= prm.B) return;
    if (prm.count && (int)blockIdx.x >= *prm.count) return;
    const int b = prm.order ? prm.order[blockIdx.x] : (int)blockIdx.x;
    const int mcap = prm.mcap;
    const bool io32 = sizeof(T) == 8 && prm.io32 != 0;
    const size_t esz = io32 ? 4 : sizeof(T);
    // re-solve of a problem the mixed-precision kernel left at -5 / 0: its last iterate is close to the solution
    // (single precision breaks down late, when the barrier terms are large), so start there with a small barrier
    const bool warm = prm.count && prm.z_warm && (prm.info_int[(size_t)b * 4] == -5 || prm.info_int[(size_t)b * 4] == 0);
    const void* z_src = warm ? prm.z_warm : prm.z0;

    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    int* nr = reinterpret_cast<int*>(smem_raw + 16);
    T* sm = reinterpret_cast<T*>(smem_raw + L::HEAD_BYTES);

    Solver<T, N, PC> s;
    s.bind(smem_raw, lane, mcap);
    s.final_variant = (prm.variant == 1);
    s.io32 = io32;
    T* const stg = sm + L::sh_off(mcap);   // TMA staging area = start of SH
    const Opts& o = prm.o;

    // ---- stage the problem into shared memory with TMA bulk copies ------------------------
    const uint32_t bytes_z = (uint32_t)(N * NZ * esz), bytes_h = (uint32_t)(N * 10 * esz), bytes_n = N * 4;
    s.rows_g = static_cast<const unsigned char*>(prm.rows) + (size_t)b * N * mcap * 4 * esz;
    T* const stg_z = stg + L::STG_HDR + N * 10;   // io32 only: the float warm start waits here for its conversion
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, bytes_z + bytes_h + bytes_n);
        tma_load(io32 ? stg_z : s.Z, static_cast<const unsigned char*>(z_src) + (size_t)b * N * NZ * esz, bytes_z, bar);
        tma_load(stg + L::STG_HDR, static_cast<const unsigned char*>(prm.hdr) + (size_t)b * N * 10 * esz, bytes_h, bar);
        tma_load(nr, prm.nrows + (size_t)b * N, bytes_n, bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    // re-layout the headers into a bank-conflict-free padded stride: N*10 -> N*11 (HDR lives beyond the staging alias)
    if (io32) {
        for (int e = lane; e < N * 10; e += 32) s.HDR[(e / 10) * L::HDR_S + (e % 10)] = (T) reinterpret_cast<const float*>(stg + L::STG_HDR)[e];
        for (int e = lane; e < N * NZ; e += 32) s.Z[e] = (T) reinterpret_cast<const float*>(stg_z)[e];
    } else {
        for (int e = lane; e < N * 10; e += 32) s.HDR[(e / 10) * L::HDR_S + (e % 10)] = stg[L::STG_HDR + e];
    }
    __syncwarp();

    // ---- initial point -------------------------------------------------------------------
    const T mu0 = warm ? (T)fmin(o.mu0, 0.1) : (T)o.mu0;
    if (lane < NZ) { s.BND[lane] = lower_bound<T>(lane); s.BND[NZ + lane] = upper_bound<T>(lane); }
    for (int e = lane; e < N * NZ; e += 32) s.DZ[e] = T(0);      // staging area is dead now; evaluate(0) reads 0 * dz
    int ncomp = 0;
    for (int k = lane; k < N; k += 32) {
#pragma unroll
        for (int i = 0; i < NZ; i++) {
            T v = s.Z[k * NZ + i];
            if (k == 0 && i >= 8)
                v = io32 ? (T) static_cast<const float*>(prm.xinit)[(size_t)b * 9 + i - 8] : static_cast<const T*>(prm.xinit)[(size_t)b * 9 + i - 8];
            if (is_free(k, i)) {
                const T lb = lower_bound<T>(i), ub = upper_bound<T>(i), kp = (T)o.kappa_push;
                const T pl = fmin(kp * fmax(T(1), fabs(lb)), kp * (ub - lb));
                const T pu = fmin(kp * fmax(T(1), fabs(ub)), kp * (ub - lb));
                v = fmin(fmax(v, lb + pl), ub - pu);
                s.ZL[k * NZ + i] = mu0 / (v - lb);
                s.ZU[k * NZ + i] = mu0 / (ub - v);
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
            T r[4]; s.load_row(k, j, r);
            T sl = (r[3] + C::hu) - (r[0] * s.Z[k * NZ + 8] + r[1] * s.Z[k * NZ + 9] + r[2] * s.Z[k * NZ + 10]);
            sl = fmax(sl, (T)o.s_floor);
            s.S[k * s.SS + j] = sl;
            s.LC[k * s.SS + j] = mu0 / sl;
            ncomp++;
        }
    }
    ncomp = warp_sum(ncomp);
    __syncwarp();

    // ---- interior-point iterations ---------------------------------------------------------
    int flag = 0, it = 0, nbt_total = 0;
    T alpha_p = T(0), alpha_d = T(0), rs_n = T(0), req_n = T(0), rin_n = T(0), rcomp = T(0), mu = T(0);
    T f_cur, th_cur, ls_cur;
    s.evaluate(T(0), f_cur, th_cur, ls_cur);
    __syncwarp();
    // The stage-0 states are fixed by the xinit equality: their bounds and the stage-0 corridor rows are not part of the
    // barrier problem.  If xinit violates one of them beyond TolIneq, the reference's NLP -- which carries them
    // (mpc_generator_normal.m:29-50) -- has no feasible point: NOPROGRESS (-7), zero iterations, violation in res_ineq.
    bool infeasible0;
    {
        T v0 = T(0);
        if (lane >= 8 && lane < NZ) v0 = fmax(lower_bound<T>(lane) - s.Z[lane], s.Z[lane] - upper_bound<T>(lane));
        const int m0 = min(nr[0], mcap);
        for (int j = lane; j < m0; j += 32) {
            T r[4]; s.load_row(0, j, r);
            v0 = fmax(v0, r[0] * s.Z[8] + r[1] * s.Z[9] + r[2] * s.Z[10] - (r[3] + C::hu));
        }
        v0 = warp_max(v0);
        infeasible0 = v0 > (T)o.tol_ineq;
        if (infeasible0) { flag = -7; rin_n = v0; }
    }
    T a_acc = T(0), ad_acc = T(0), mu_acc = mu0;     // the step accepted by the last line search (none yet)
    if constexpr (!PC) {
        for (int e = lane; e < N * NXI; e += 32) s.P[e] = T(0);
        __syncwarp();
    }
    for (it = 0; !infeasible0; it++) {
        T csum, cmin;
        if constexpr (PC) s.residuals(rs_n, req_n, rin_n, rcomp, csum, cmin);
        else { s.post_step(mu_acc, a_acc, ad_acc, rs_n, req_n, rin_n, rcomp, csum, cmin); __syncwarp(); }   // G is complete for finish_rhs()
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
        T mu_t = fmax(sigma * mu, (T)o.mu_floor);
        T dza[PC ? L::NT : 1];              // predictor-corrector: this lane's slice of the affine step
        bool ok;
        if constexpr (PC) {
            // Mehrotra predictor-corrector: one factorisation, two solves.  The affine solve (mu = 0) runs with the
            // stage-phase state parked; the analysis needs that state AND must keep the gains, so the two trade
            // places (registers <-> shared memory) around it; the corrector is a vector-only sweep.
            s.assemble(T(0));
            __syncwarp();
            T parked[L::NPARK_LANE];
            s.park(parked);
            __syncwarp();
            ok = s.riccati_backward();
            ok &= s.template rollout<false>();
            __syncwarp();
            s.swap_parked(parked);          // state back in shared memory, gains / Quu^-1 / P_0qq into the registers
            __syncwarp();
            mu_t = s.predictor_corrector_rhs(mu, csum, ncomp, (T)o.mu_floor, dza);
            __syncwarp();
            s.swap_parked(parked);
            __syncwarp();
            s.delta_backward();
            s.template rollout<true>();
            s.costates();
            __syncwarp();
            s.unpark(parked);
            __syncwarp();
        } else {
            s.finish_rhs(mu_t);
            __syncwarp();
            T parked[L::NPARK_LANE];
            s.park(parked);                 // z, z_l, z_u, y, ... leave shared memory for the sweeps
            __syncwarp();
            ok = s.riccati_backward();
            ok &= s.rollout();
            s.costates();
            __syncwarp();
            s.unpark(parked);
            __syncwarp();
        }
        if (!ok) { flag = -5; break; }
        const T tau = fmin(fmax(T(0.995), T(1) - mu), T(0.99999));
        T ap, ad;
        s.step_lengths(mu_t, tau, ap, ad, dza);
        // backtracking line search on (theta, barrier objective)
        const T ph0 = f_cur - mu_t * ls_cur;
        // theta below 1% of TolEq counts as feasible (also absorbs the rounding floor of theta)
        const T th_noise = fmax(T(10) * Eps<T>::v * T(N * NXI) * T(20), T(0.01) * (T)o.tol_eq);
        T a = ap;
        int nbt = 0;
        T ft, tht, lst;
        for (;;) {
            s.evaluate(a, ft, tht, lst);
            __syncwarp();
            const T pht = ft - mu_t * lst;
            const bool acc = (tht <= fmax((T(1) - T(1e-5)) * th_cur, th_noise)) ||
                             (pht <= ph0 - T(1e-5) * th_cur + T(10) * Eps<T>::v * fabs(ph0));
            if (acc || nbt >= o.max_bt) break;
            nbt++;
            a *= T(0.5);
        }
        nbt_total += nbt;
        alpha_p = a; alpha_d = ad;
        if constexpr (PC) s.update(mu_t, a, ad, dza);     // gradient / Jacobians / defects of the accepted trial stay in place
        else { a_acc = a; ad_acc = ad; mu_acc = mu_t; }   // taken by post_step() at the top of the next pass
        f_cur = ft; th_cur = tht; ls_cur = lst;
        __syncwarp();
    }

    // ---- results -----------------------------------------------------------------------------
    __syncwarp();
    auto put = [&](void* base, size_t idx, T v) {
        if (io32) static_cast<float*>(base)[idx] = (float)v; else static_cast<T*>(base)[idx] = v;
    };
    store_solution<32>(prm.peers, prm.z_out, prm.info_int, (size_t)b, s.Z, N * NZ, io32, lane, flag, it, nbt_total, prm.count ? 1 : 0);
    if (lane == 0) {
        const T v[8] = {req_n, rin_n, rs_n, rcomp, f_cur, mu, alpha_p, alpha_d};
#pragma unroll
        for (int q = 0; q < 8; q++) put(prm.info_real, (size_t)b * 8 + q, v[q]);
    }
    if (prm.y_out)
        for (int e = lane; e < N * NXI; e += 32) put(prm.y_out, (size_t)b * N * NXI + e, (e < NXI) ? T(0) : s.Y[e]);
    if (prm.zl_out)
        for (int e = lane; e < N * NZ; e += 32) put(prm.zl_out, (size_t)b * N * NZ + e, s.ZL[e]);
    if (prm.zu_out)
        for (int e = lane; e < N * NZ; e += 32) put(prm.zu_out, (size_t)b * N * NZ + e, s.ZU[e]);
    if (prm.lc_out)
        for (int e = lane; e < N * mcap; e += 32) {
            const int k = e / mcap, j = e - k * mcap;
            put(prm.lc_out, (size_t)b * N * mcap + e, (j < s.live(k)) ? s.LC[k * s.SS + j] : T(0));
        }
}

}  // namespace nmpc
