// nmpc_backsolve.cuh -- stand-alone structured KKT factorisation and backsolve (sm_100a).
//
// The reference solves every Newton system with a stored block factor and a separate
// forward/backward substitution (binary symbols f_17_ldl_forward_solve_rm,
// f_17_backward_solve_rm, f_13_ldl_forward_solve_rm, f_13_backward_solve_rm of
// /root/reference/src/resilient_planner/plan_manage/solver/normal/FORCESNLPsolver_normal/lib/
// libFORCESNLPsolver_normal.so).  These two kernels expose the same split for
// the Riccati factor this library uses:
//
//   riccati_factor_kernel   Phi (compact stage Hessians) + J (compact Jacobians)  ->  factor in HBM
//   kkt_backsolve_kernel    factor + right-hand side (g, d)                       ->  (dz, y)
//
// Stored factor per problem, N * FAC_WORDS (= 204) words in two regions so that each can be fetched
// on its own:   P region   [N][91]   P_k packed lower
//               KQJ region [N][113]  K_k 4x13 | Quu^-1 packed lower 10 | J_k compact 51.
// The backsolve is the memory-bound piece: per problem it must read N*204 factor words + N*30 rhs
// words and write N*30 solution words exactly once (N = 20, fp64: 42 240 B) -- and DRAM sees exactly
// that.  Both sweeps run out of shared memory; the solution leaves by bulk stores.  The serial part
// of each sweep is kept to ~12 dependent FMAs per stage by hoisting the 13x13 products (P+ d before
// the backward sweep, P dxi after the forward sweep) out of the recursion, where they run
// lane-parallel over all stages at once.  Because P is needed only by those two stage-parallel phases
// and K | Quu^-1 | J only by the sweeps between them, the two regions time-share the same shared-memory
// slots (BsLayout, OVL), which is what sets the number of resident problems per SM.
#pragma once
#include "nmpc_ipm.cuh"

namespace nmpc {

template <typename T> struct FactorParams {
    int B;
    const T* phi;   // [B][N][21]  diag(17) | pos-block off-diag (01,02,12) | u/u_prev coupling
    const T* jc;    // [B][N][51]  compact dynamics Jacobians (stage N-1 ignored)
    T* fac;         // [B][N][204]
    int* status;    // [B] 0 ok, -5 non-positive pivot
};

template <typename T, int N>
__global__ void __launch_bounds__(32) riccati_factor_kernel(const FactorParams<T> prm)
{
    using L = Layout<T, N>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x, b = blockIdx.x;
    if (b >= prm.B) return;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    Solver<T, N> s;
    s.bind(smem_raw, lane, 0);
    s.final_variant = false;
    s.fac_out = prm.fac + (size_t)b * N * FAC_WORDS;
    const uint32_t bytes_phi = N * L::PHI_S * sizeof(T), bytes_jc = N * NJC * sizeof(T);
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, bytes_phi + bytes_jc);
        tma_load(s.PHID, prm.phi + (size_t)b * N * L::PHI_S, bytes_phi, bar);
        tma_load(s.JC, prm.jc + (size_t)b * N * NJC, bytes_jc, bar);
    }
    for (int e = lane; e < N * NZ; e += 32) s.G[e] = T(0);
    for (int e = lane; e < N * NXI; e += 32) { s.D[e] = T(0); s.P[e] = T(0); }
    __syncwarp();
    mbar_wait(bar, 0);
    const bool ok = s.riccati_backward();
    if (lane == 0) prm.status[b] = ok ? 0 : -5;
}

// ---------------------------------------------------------------------------------------------
template <typename T> struct BacksolveParams {
    int B;
    const T* fac;   // [B][ P: N x 91 | KQJ: N x 113 ]   (see the file header)
    const T* g;     // [B][N][17]   gradient of the QP (barrier-augmented)
    const T* d;     // [B][N][13]   dynamics defects, c-ordering (row N-1 unused)
    T* dz;          // [B][N][17]
    T* y;           // [B][N][13]   costates, y[0] = 0
};

// Shared-memory plan of one problem.  The factor moves in groups of GS = 4 stages (4 * 91 and 4 * 113
// words are multiples of the 16-byte TMA granule in fp32 and fp64).  The P slices are only needed
// by the two stage-parallel phases (before the backward sweep, after the forward sweep), the
// K | Quu^-1 | J slices only by the two sweeps in between, so they time-share the same slots: slot j
// holds P group j, then KQJ group j (fetched as soon as the first phase is done with the slot; that
// phase runs last group first, i.e. in the order the backward sweep wants them), then P group j
// again (re-fetched -- an L2 hit, the in-flight working set of the whole chip is < 20 MB -- as soon
// as the forward sweep has passed it).  That is 14.6 KB less shared memory per problem
// (fp64, N = 20): 8 resident problems per SM instead of 5, and DRAM still sees every factor byte
// exactly once.
template <typename T, int N> struct BsLayout {
    static constexpr int GS = 4;
    static_assert(N % GS == 0, "horizon must be a multiple of the TMA group");
    static constexpr int NG = N / GS;
    static constexpr int PW = 91, KW = FAC_WORDS - 91;          // 113 = K 52 | Quu^-1 10 | J 51
    static constexpr int HEAD_BYTES = (8 * (1 + NG) + 15) & ~15;   // mbarriers: initial load + one per group
    static constexpr int SLOT = GS * KW;
    static constexpr int GZ = NG * SLOT;             // g on entry, dz on exit
    static constexpr int DD = GZ + N * NZ;
    static constexpr int WY = DD + N * NXI;          // P+ d (hoisted) -> p_k (backward) -> y (exit), all in place
    static constexpr int KF = WY + N * NXI;          // feed-forward terms
    static constexpr int TV = KF + N * 4;            // 13 + pad
    static constexpr int DXI = TV + 16;
    static constexpr int TOTAL = DXI + 16;
    static_assert(GZ % 2 == 0 && DD % 2 == 0 && WY % 2 == 0, "16-byte alignment of the fp64 vector loads");
    static constexpr size_t bytes() { return HEAD_BYTES + (size_t)TOTAL * sizeof(T); }
    __host__ __device__ static constexpr int p_off(int k) { return (k / GS) * SLOT + (k % GS) * PW; }
    __host__ __device__ static constexpr int k_off(int k) { return k * KW; }
    // algorithmic words per problem: factor + rhs read once, solution written once
    static constexpr int ALGO_WORDS = N * (FAC_WORDS + 2 * (NZ + NXI));
};

// y_i = sum_j P[i][j] x[j] for the conflict-free symmetric layout, poff[j] = PSYM[i][j] (+ stage offset)
// precomputed per lane; three independent chains.
template <typename T> __device__ __forceinline__ T sym13_row_dot(const T* Pk, const int poff[NXI], const T (&x)[NXI], T init)
{
    T c0 = init, c1 = T(0), c2 = T(0);
#pragma unroll
    for (int j = 0; j < NXI; j++) {
        const T t = Pk[poff[j]] * x[j];
        if (j % 3 == 0) c0 += t; else if (j % 3 == 1) c1 += t; else c2 += t;
    }
    return (c0 + c1) + c2;
}

// 13 consecutive words into registers; fp64: six 16-byte loads + one 8-byte load (`a16` says whether p
// itself is 16-byte aligned -- a compile-time fact once the caller's loop is unrolled)
template <typename T> __device__ __forceinline__ void load13(const T* p, bool a16, T (&x)[NXI])
{
    if constexpr (sizeof(T) == 8) {
        const int o = a16 ? 0 : 1;
        if (a16) x[12] = p[12]; else x[0] = p[0];
#pragma unroll
        for (int q = 0; q < 6; q++) {
            const double2 v = *reinterpret_cast<const double2*>(p + o + 2 * q);
            x[o + 2 * q] = v.x; x[o + 2 * q + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int q = 0; q < NXI; q++) x[q] = p[q];
    }
}

template <typename T, int N>
__global__ void __launch_bounds__(32) kkt_backsolve_kernel(const BacksolveParams<T> prm)
{
    using L = BsLayout<T, N>;
    using C = Const<T>;
    constexpr int GS = L::GS, NG = L::NG, PW = L::PW, KW = L::KW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x, b = blockIdx.x;
    if (b >= prm.B) return;
    uint64_t* bar0 = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* barg = bar0 + 1;                                 // [NG]  phase 0: KQJ group landed, phase 1: P group re-landed
    T* sm = reinterpret_cast<T*>(smem_raw + L::HEAD_BYTES);
    T* SL = sm;
    T* GZ = sm + L::GZ; T* DD = sm + L::DD; T* WY = sm + L::WY;
    T* KF = sm + L::KF; T* TV = sm + L::TV; T* DXI = sm + L::DXI;

    const T* facP = prm.fac + (size_t)b * N * FAC_WORDS;       // P region of this problem
    const T* facK = facP + N * PW;                             // K | Quu^-1 | J region
    constexpr uint32_t bytes_pg = GS * PW * sizeof(T), bytes_kg = GS * KW * sizeof(T);
    constexpr uint32_t bytes_g = N * NZ * sizeof(T), bytes_d = N * NXI * sizeof(T);
    if (lane == 0) {
        mbar_init(bar0, 1);
        for (int j = 0; j < NG; j++) mbar_init(barg + j, 1);
        mbar_expect_tx(bar0, N * PW * sizeof(T) + bytes_g + bytes_d);
        for (int j = NG - 1; j >= 0; j--) tma_load(SL + j * L::SLOT, facP + j * GS * PW, bytes_pg, bar0);
        tma_load(DD, prm.d + (size_t)b * N * NXI, bytes_d, bar0);
        tma_load(GZ, prm.g + (size_t)b * N * NZ, bytes_g, bar0);
    }
    // ---- per-lane tables (overlap with the copy) ---------------------------------------------------
    // backward: lanes 0..16 own one z index each (q~ = g + J' tv); the ten that need Jacobian words sit in the
    // first half-warp (one shared-memory wavefront per load): z 16 and z 4 swap lanes.
    //   J' tv = sum_r jc[oA + sA r] tv[r] + sum_r jc[oB + sB r] tv[3+r] + c1 tv[i1] + c2 tv[i2]
    const int zi = lane == 4 ? 16 : (lane == 16 ? 4 : (lane < NZ ? lane : 0));
    const int zt = zi < 3 ? 0 : (zi == 3 ? 1 : (zi < 8 ? 2 : (zi < 11 ? 3 : (zi < 14 ? 4 : 5))));   // rate, T, uprev, pos, vel, rpy
    const int zj = zt == 0 ? zi : (zt == 3 ? zi - 8 : (zt == 4 ? zi - 11 : (zt == 5 ? zi - 14 : 0)));
    const int oA = zt == 1 ? JPT : (zt == 4 ? JPV + zj : (zt == 5 ? JPR + zj : 0));
    const int sA = zt == 1 ? 1 : 3;
    const bool needA = lane < 16 && (zt == 1 || zt == 4 || zt == 5);
    const int oB = zt == 0 ? JVW + zj : (zt == 1 ? JVT : (zt == 4 ? JVV + zj : (zt == 5 ? JVR + zj : 0)));
    const int sB = zt == 1 ? 1 : 3;
    const bool needB = lane < 16 && (zt == 0 || zt == 1 || zt == 4 || zt == 5);
    const T c1 = zt == 0 ? C::h : ((zt == 3 || zt == 5) ? T(1) : T(0));
    const int i1 = zt == 0 ? 6 + zj : (zt == 3 ? zj : (zt == 5 ? 6 + zj : 0));
    const T c2 = (zt == 0 || zt == 1) ? T(1) : T(0);
    const int i2 = zt == 0 ? 9 + zj : 12;
    const int xi = lane < NXI ? lane : 0;                 // xi index handled by this lane
    const int xzz = e_col(xi);                            // its z index ...
    const int xz = xzz == 16 ? 4 : (xzz == 4 ? 16 : xzz); // ... and the lane that owns it (shuffle source for q_xi)
    // forward: lane = row of dxi+ (as in the fused solver's rollout)
    const int rt = lane < 3 ? 0 : (lane < 6 ? 1 : (lane < 9 ? 2 : (lane < 13 ? 3 : 4)));
    const int rr = lane < 3 ? lane : (lane < 6 ? lane - 3 : 0);
    const int offT = rt == 0 ? JPT + rr : (rt == 1 ? JVT + rr : 0);
    const int offV = rt == 0 ? JPV + 3 * rr : (rt == 1 ? JVV + 3 * rr : 0);
    const int offR = rt == 0 ? JPR + 3 * rr : (rt == 1 ? JVR + 3 * rr : 0);
    const int offW = rt == 1 ? JVW + 3 * rr : 0;
    const bool needMain = rt <= 1, needW = rt == 1;
    const T mSelf = (rt == 0 || rt == 2) ? T(1) : T(0);
    const T cDu = rt == 2 ? C::h : (rt == 3 ? T(1) : T(0));
    const int duSrc = 8 * ((rt == 2 ? lane - 6 : lane - 9) & 3);
    const int r4 = lane >> 3, part = lane & 7;
    const int gsrc = lane < 8 ? 5 + lane : (lane < NZ ? lane - 8 : 0);    // dz_k[lane] = dxi[gsrc] for lanes 4..16
    // stage-parallel symmetric products: one stage per half-warp (lanes 0..12 and 16..28 = rows), the two
    // stages of a pass two apart (same 16-byte phase of their vectors).  PSYM makes every column access
    // of the 13 row-lanes hit 13 different banks.
    const int hrow = (lane & 15) < NXI ? (lane & 15) : 0, hsub = lane >> 4;
    const bool hact = (lane & 15) < NXI;
    int poff[NXI];
    {
        const uint4 w = *reinterpret_cast<const uint4*>(&PSYM[hrow][0]);
        const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < NXI; j++) poff[j] = (int)((ww[j >> 2] >> (8 * (j & 3))) & 0xffu) + hsub * 2 * PW;
    }
    // uniform 4-term product after the q_u broadcast: lanes 0..12 -> column xi of K (p_k), lanes 16..19 -> row of Quu^-1 (kff)
    int uoff[4];
#pragma unroll
    for (int c = 0; c < 4; c++) uoff[c] = (lane >= 16 && lane < 20) ? 52 + pk(lane - 16, c) : 13 * c + xi;
    __syncwarp();
    mbar_wait(bar0, 0);

    // the 4x4 block P_0[q][q] is all the stage-0 solve needs of P_0: keep it in registers (its slot is recycled)
    T pqq[10];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c <= r; c++) pqq[r * (r + 1) / 2 + c] = SL[L::p_off(0) + PSYM[9 + r][9 + c]];

    // ---- stage-parallel: w_{k-1} = P_k d_{k-1}  (k = 1..N-1), last group first so that the K | Quu^-1 | J
    //      slices arrive in the order the backward sweep wants them ------------------------------------------
#pragma unroll
    for (int j = NG - 1; j >= 0; j--) {
#pragma unroll
        for (int sub = 0; sub < 2; sub++) {
            const int ka = j * GS + sub;                           // first half-warp: stage ka, second: ka + 2
            if (hact && ka + 2 * hsub >= 1) {
                T x[NXI];
                load13<T>(DD + (ka - 1) * NXI + hsub * 2 * NXI, ((ka - 1) & 1) == 0, x);
                WY[(ka - 1) * NXI + hsub * 2 * NXI + hrow] = sym13_row_dot<T>(SL + L::p_off(ka), poff, x, T(0));
            }
        }
        __syncwarp();
        if (lane == 0) {                                       // slot j is free: it takes KQJ group j
            mbar_expect_tx(barg + j, bytes_kg);
            tma_load(SL + j * L::SLOT, facK + j * GS * KW, bytes_kg, barg + j);
        }
    }
    __syncwarp();

    // ---- backward sweep: lane i (< 13) carries p_{k+1}[i] in a register; p_k overwrites w_k -------------
    T pnext = T(0);
#pragma unroll
    for (int k = N - 1; k >= 0; k--) {
        if ((k % GS) == GS - 1) mbar_wait(barg + k / GS, 0);
        const T* fk = SL + L::k_off(k);
        const T* jc = fk + 62;
        T qz = GZ[k * NZ + zi];
        if (k < N - 1) {
            T ja0 = T(0), ja1 = T(0), ja2 = T(0), jb0 = T(0), jb1 = T(0), jb2 = T(0);
            if (needA) { ja0 = jc[oA]; ja1 = jc[oA + sA]; ja2 = jc[oA + 2 * sA]; }
            if (needB) { jb0 = jc[oB]; jb1 = jc[oB + sB]; jb2 = jc[oB + 2 * sB]; }
            if (lane < NXI) TV[lane] = pnext + WY[k * NXI + lane];
            __syncwarp();
            const T sa = (ja0 * TV[0] + ja1 * TV[1]) + ja2 * TV[2];
            const T sb = (jb0 * TV[3] + jb1 * TV[4]) + jb2 * TV[5];
            qz = (sa + (c1 * TV[i1] + qz)) + (sb + c2 * TV[i2]);
        }
        const T qu0 = __shfl_sync(0xffffffffu, qz, 0), qu1 = __shfl_sync(0xffffffffu, qz, 1);
        const T qu2 = __shfl_sync(0xffffffffu, qz, 2), qu3 = __shfl_sync(0xffffffffu, qz, 3);
        const T qxi = __shfl_sync(0xffffffffu, qz, xz);
        const T u4 = (fk[uoff[0]] * qu0 + fk[uoff[1]] * qu1) + (fk[uoff[2]] * qu2 + fk[uoff[3]] * qu3);
        pnext = qxi + u4;                                                // lanes 0..12: p_k = q_xi + K' q_u
        if (lane < NXI) WY[k * NXI + lane] = pnext;
        else if (lane >= 16 && lane < 20) KF[k * 4 + lane - 16] = -u4;   // lanes 16..19: kff = -Quu^-1 q_u
        __syncwarp();
    }

    // ---- stage 0: x fixed, u_prev free: dq = -Pqq^-1 p_q --------------------------------------------------
    {
        T a[16], l[10], li[4], x[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int c = 0; c <= r; c++) a[4 * r + c] = pqq[r * (r + 1) / 2 + c];
            x[r] = -WY[9 + r];
        }
        chol4<T>(a, l, li);
        fsub4<T>(l, li, x);
        bsub4<T>(l, li, x);
        if (lane < NXI) DXI[lane] = (lane < 9) ? T(0) : (lane == 9 ? x[0] : (lane == 10 ? x[1] : (lane == 11 ? x[2] : x[3])));
    }
    __syncwarp();

    // ---- forward sweep (dz overwrites g) -----------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < N; k++) {
        const T* fk = SL + L::k_off(k);
        const T* jc = fk + 62;
        const T* kg = fk + r4 * 13;
        T acc = kg[part] * DXI[part];
        if (part < 5) acc += kg[part + 8] * DXI[part + 8];
        // the part of dxi+ that does not wait for du
        T pre = T(0), jT = T(0), jw0 = T(0), jw1 = T(0), jw2 = T(0);
        if (k < N - 1) {
            T jv0 = T(0), jv1 = T(0), jv2 = T(0), jr0 = T(0), jr1 = T(0), jr2 = T(0);
            if (needMain) {
                jv0 = jc[offV]; jv1 = jc[offV + 1]; jv2 = jc[offV + 2];
                jr0 = jc[offR]; jr1 = jc[offR + 1]; jr2 = jc[offR + 2];
                jT = jc[offT];
            }
            if (needW) { jw0 = jc[offW]; jw1 = jc[offW + 1]; jw2 = jc[offW + 2]; }
            const T va = (jv0 * DXI[3] + jv1 * DXI[4]) + jv2 * DXI[5];
            const T ra = (jr0 * DXI[6] + jr1 * DXI[7]) + jr2 * DXI[8];
            pre = (va + ra) + (mSelf * DXI[xi] + DD[k * NXI + xi]);
        }
        const T dsrc = DXI[gsrc];
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += KF[k * 4 + r4];
        const T dw0 = __shfl_sync(0xffffffffu, acc, 0), dw1 = __shfl_sync(0xffffffffu, acc, 8);
        const T dw2 = __shfl_sync(0xffffffffu, acc, 16), dT = __shfl_sync(0xffffffffu, acc, 24);
        const T du_l = __shfl_sync(0xffffffffu, acc, 8 * (lane & 3));
        const T du_s = __shfl_sync(0xffffffffu, acc, duSrc);
        if (lane < NZ) GZ[k * NZ + lane] = (lane < 4) ? du_l : dsrc;
        const T nxt = ((jT * dT + pre) + (jw0 * dw0 + cDu * du_s)) + (jw1 * dw1 + jw2 * dw2);
        __syncwarp();
        if (k < N - 1 && lane < NXI) DXI[lane] = nxt;
        if ((k % GS) == GS - 1 && lane == 0) {    // the sweep has left group k / GS: its slot takes P again (L2 hit)
            mbar_expect_tx(barg + k / GS, bytes_pg);
            tma_load(SL + (k / GS) * L::SLOT, facP + (k / GS) * GS * PW, bytes_pg, barg + k / GS);
        }
        __syncwarp();
    }

    // ---- stage-parallel: y_k = P_k dxi_k + p_k   (k = 1..N-1),  y_0 = 0 (in place over p_k) ------------------
#pragma unroll
    for (int j = 0; j < NG; j++) {
        mbar_wait(barg + j, 1);
#pragma unroll
        for (int sub = 0; sub < 2; sub++) {
            const int ka = j * GS + sub;
            if (hact && ka + 2 * hsub >= 1) {
                T w[NXI];                                          // dz[4..16] = (dq, dx)
                load13<T>(GZ + ka * NZ + 4 + hsub * 2 * NZ, (ka & 1) == 0, w);
                const T x[NXI] = {w[4], w[5], w[6], w[7], w[8], w[9], w[10], w[11], w[12], w[0], w[1], w[2], w[3]};
                T* yk = WY + ka * NXI + hsub * 2 * NXI + hrow;
                *yk = sym13_row_dot<T>(SL + L::p_off(ka), poff, x, *yk);
            }
        }
    }
    if (lane < NXI) WY[lane] = T(0);
    __syncwarp();
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(prm.dz + (size_t)b * N * NZ),
                     "r"(smem_u32(GZ)), "r"(bytes_g) : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(prm.y + (size_t)b * N * NXI),
                     "r"(smem_u32(WY)), "r"(bytes_d) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

}  // namespace nmpc
