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
// Stored factor per stage (FAC_WORDS = 204 words): P_k packed lower 91 | K_k 4x13 | Quu^-1 packed
// lower 10 | J_k compact 51.  The backsolve is the memory-bound piece: per problem it must read
// N*204 factor words + N*30 rhs words and write N*30 solution words exactly once
// (N = 20, fp64: 42 240 B), and it does exactly that -- the whole factor of a problem is pulled
// into shared memory by one TMA bulk copy, both sweeps run out of shared memory, and the
// solution leaves by bulk stores.  The serial part of each sweep is kept to ~12 dependent FMAs per
// stage by hoisting the 13x13 products (P+ d before the backward sweep, P dxi after the forward
// sweep) out of the recursion, where they run lane-parallel over all stages at once.
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
    const T* fac;   // [B][N][204]
    const T* g;     // [B][N][17]   gradient of the QP (barrier-augmented)
    const T* d;     // [B][N][13]   dynamics defects, c-ordering (row N-1 unused)
    T* dz;          // [B][N][17]
    T* y;           // [B][N][13]   costates, y[0] = 0
};

template <typename T, int N> struct BsLayout {
    static constexpr int HEAD_BYTES = 16;
    static constexpr int FAC = 0;
    static constexpr int GZ = FAC + N * FAC_WORDS;   // g on entry, dz on exit
    static constexpr int DD = GZ + N * NZ;
    static constexpr int WY = DD + N * NXI;          // P+ d (hoisted) -> p_k (backward) -> y (exit), all in place
    static constexpr int KF = WY + N * NXI;          // feed-forward terms
    static constexpr int TV = KF + N * 4;            // 13 + pad
    static constexpr int DXI = TV + 16;
    static constexpr int TOTAL = DXI + 16;
    static constexpr size_t bytes() { return HEAD_BYTES + (size_t)TOTAL * sizeof(T); }
    // algorithmic words per problem: factor + rhs read once, solution written once
    static constexpr int ALGO_WORDS = N * (FAC_WORDS + 2 * (NZ + NXI));
};

__device__ __forceinline__ int pk(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

// y_i = sum_j P[i][j] x[j] for a packed-lower symmetric 13x13 with the 13 packed offsets of row i
// precomputed (poff[j] = pk(i, j)); three independent chains.
template <typename T> __device__ __forceinline__ T sym13_row_dot(const T* Pk, const int poff[NXI], const T* x, T init)
{
    T c0 = init, c1 = T(0), c2 = T(0);
#pragma unroll
    for (int j = 0; j < NXI; j++) {
        const T t = Pk[poff[j]] * x[j];
        if (j % 3 == 0) c0 += t; else if (j % 3 == 1) c1 += t; else c2 += t;
    }
    return (c0 + c1) + c2;
}

template <typename T, int N>
__global__ void __launch_bounds__(32) kkt_backsolve_kernel(const BacksolveParams<T> prm)
{
    using L = BsLayout<T, N>;
    using C = Const<T>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x, b = blockIdx.x;
    if (b >= prm.B) return;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    T* sm = reinterpret_cast<T*>(smem_raw + L::HEAD_BYTES);
    T* FAC = sm + L::FAC; T* GZ = sm + L::GZ; T* DD = sm + L::DD; T* WY = sm + L::WY;
    T* KF = sm + L::KF; T* TV = sm + L::TV; T* DXI = sm + L::DXI;

    constexpr uint32_t bytes_f = N * FAC_WORDS * sizeof(T), bytes_g = N * NZ * sizeof(T), bytes_d = N * NXI * sizeof(T);
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, bytes_f + bytes_g + bytes_d);
        tma_load(FAC, prm.fac + (size_t)b * N * FAC_WORDS, bytes_f, bar);
        tma_load(GZ, prm.g + (size_t)b * N * NZ, bytes_g, bar);
        tma_load(DD, prm.d + (size_t)b * N * NXI, bytes_d, bar);
    }
    // ---- per-lane tables (overlap with the copy) ---------------------------------------------------
    // backward: lane = z index zi of q~ = g + J' tv ;  J' tv = mA sum_r jc[oA + sA r] tv[r] + mB sum_r jc[oB + sB r] tv[3+r]
    //           + c1 tv[i1] + c2 tv[i2]
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
    const int xi = lane < NXI ? lane : 0;                 // xi index handled by this lane
    const int xz = e_col(xi);                             // its z index (shuffle source for q_xi)
    // forward: lane = row of dxi+ (as in the fused solver's rollout)
    const int rt = lane < 3 ? 0 : (lane < 6 ? 1 : (lane < 9 ? 2 : (lane < 13 ? 3 : 4)));
    const int rr = lane < 3 ? lane : (lane < 6 ? lane - 3 : 0);
    const int offT = rt == 0 ? JPT + rr : (rt == 1 ? JVT + rr : 0);
    const int offV = rt == 0 ? JPV + 3 * rr : (rt == 1 ? JVV + 3 * rr : 0);
    const int offR = rt == 0 ? JPR + 3 * rr : (rt == 1 ? JVR + 3 * rr : 0);
    const int offW = rt == 1 ? JVW + 3 * rr : 0;
    const T mMain = rt <= 1 ? T(1) : T(0), mW = rt == 1 ? T(1) : T(0);
    const T mSelf = (rt == 0 || rt == 2) ? T(1) : T(0);
    const T cDu = rt == 2 ? C::h : (rt == 3 ? T(1) : T(0));
    const int duSrc = 8 * ((rt == 2 ? lane - 6 : lane - 9) & 3);
    const int r4 = lane >> 3, part = lane & 7;
    // hoisted symmetric products: lanes 0..12 -> row `lane` of an even pass stage, lanes 13..25 -> odd
    const int hrow = lane < 13 ? lane : (lane < 26 ? lane - 13 : 0), hsub = lane < 13 ? 0 : 1;
    const bool hact = lane < 26;
    int poff[NXI];
#pragma unroll
    for (int j = 0; j < NXI; j++) poff[j] = pk(hrow, j);
    // uniform 4-term product after the q_u broadcast: lanes 0..12 -> column xi of K (p_k), lanes 16..19 -> row of Quu^-1 (kff)
    int uoff[4];
#pragma unroll
    for (int c = 0; c < 4; c++) uoff[c] = (lane >= 16 && lane < 20) ? 143 + pk(lane - 16, c) : 91 + 13 * c + xi;
    __syncwarp();
    mbar_wait(bar, 0);

    // ---- hoisted, stage-parallel: w_k = P_{k+1} d_k   (k = 0..N-2), two stages per pass --------------
    for (int k0 = 0; k0 < N - 1; k0 += 2) {
        const int k = k0 + hsub;
        if (hact && k < N - 1) WY[k * NXI + hrow] = sym13_row_dot<T>(FAC + (k + 1) * FAC_WORDS, poff, DD + k * NXI, T(0));
    }
    __syncwarp();

    // ---- backward sweep: lane i (< 13) carries p_{k+1}[i] in a register; p_k overwrites w_k -------------
    T pnext = T(0);
    for (int k = N - 1; k >= 0; k--) {
        const T* fk = FAC + k * FAC_WORDS;
        const T* jc = fk + 153;
        T qz = GZ[k * NZ + zi];
        if (k < N - 1) {
            if (lane < NXI) TV[lane] = pnext + WY[k * NXI + lane];
            __syncwarp();
            const T s0 = mA * jc[oA] * TV[0] + mB * jc[oB] * TV[3];
            const T s1 = mA * jc[oA + sA] * TV[1] + mB * jc[oB + sB] * TV[4];
            const T s2 = mA * jc[oA + 2 * sA] * TV[2] + mB * jc[oB + 2 * sB] * TV[5];
            qz += ((s0 + s1) + s2) + (c1 * TV[i1] + c2 * TV[i2]);
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
            for (int c = 0; c <= r; c++) a[4 * r + c] = FAC[pk(9 + r, 9 + c)];
            x[r] = -WY[9 + r];
        }
        chol4<T>(a, l, li);
        fsub4<T>(l, li, x);
        bsub4<T>(l, li, x);
        if (lane < NXI) DXI[lane] = (lane < 9) ? T(0) : (lane == 9 ? x[0] : (lane == 10 ? x[1] : (lane == 11 ? x[2] : x[3])));
    }
    __syncwarp();

    // ---- forward sweep (dz overwrites g) -----------------------------------------------------------------------
    for (int k = 0; k < N; k++) {
        const T* fk = FAC + k * FAC_WORDS;
        const T* jc = fk + 153;
        const T* kg = fk + 91 + r4 * 13;
        T acc = kg[part] * DXI[part];
        if (part < 5) acc += kg[part + 8] * DXI[part + 8];
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += KF[k * 4 + r4];
        const T dw0 = __shfl_sync(0xffffffffu, acc, 0), dw1 = __shfl_sync(0xffffffffu, acc, 8);
        const T dw2 = __shfl_sync(0xffffffffu, acc, 16), dT = __shfl_sync(0xffffffffu, acc, 24);
        const T du_l = __shfl_sync(0xffffffffu, acc, 8 * (lane & 3));
        const T du_s = __shfl_sync(0xffffffffu, acc, duSrc);
        const T self = DXI[xi];
        if (lane < NZ) GZ[k * NZ + lane] = (lane < 4) ? du_l : (lane < 8 ? DXI[5 + lane] : DXI[lane - 8]);
        T nxt = T(0);
        if (k < N - 1) {
            const T m0 = jc[offT] * dT + jc[offV] * DXI[3] + jc[offR] * DXI[6];
            const T m1 = jc[offV + 1] * DXI[4] + jc[offR + 1] * DXI[7];
            const T m2 = jc[offV + 2] * DXI[5] + jc[offR + 2] * DXI[8];
            const T mw = jc[offW] * dw0 + jc[offW + 1] * dw1 + jc[offW + 2] * dw2;
            nxt = DD[k * NXI + xi] + mSelf * self + cDu * du_s + mMain * ((m0 + m1) + m2) + mW * mw;
        }
        __syncwarp();
        if (k < N - 1 && lane < NXI) DXI[lane] = nxt;
        __syncwarp();
    }

    // ---- hoisted, stage-parallel: y_k = P_k dxi_k + p_k   (k = 1..N-1),  y_0 = 0 (in place over p_k) ---------
    for (int k0 = 1; k0 < N; k0 += 2) {
        const int k = k0 + hsub;
        if (hact && k < N) {
            const T* dz = GZ + k * NZ;
            const T x[NXI] = {dz[8], dz[9], dz[10], dz[11], dz[12], dz[13], dz[14], dz[15], dz[16], dz[4], dz[5], dz[6], dz[7]};
            WY[k * NXI + hrow] = sym13_row_dot<T>(FAC + k * FAC_WORDS, poff, x, WY[k * NXI + hrow]);
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
