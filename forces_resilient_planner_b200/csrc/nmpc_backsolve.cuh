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
    T* sm = reinterpret_cast<T*>(smem_raw + L::HEAD_BYTES);
    Solver<T, N> s;
    s.sm = sm; s.nr = reinterpret_cast<int*>(smem_raw + 16); s.lane = lane; s.mcap = 0; s.RS = 1; s.SS = 1;
    s.final_variant = false;
    s.Z = sm + L::Z; s.DZ = sm + L::DZ; s.ZL = sm + L::ZL; s.ZU = sm + L::ZU; s.G = sm + L::G;
    s.Y = sm + L::Y; s.P = sm + L::P; s.D = sm + L::D; s.JC = sm + L::JC; s.PHID = sm + L::PHID;
    s.KG = sm + L::KG; s.KFF = sm + L::KFF; s.HDR = sm + L::HDR;
    s.ROWS = sm + L::rows_off(); s.S = sm + L::s_off(0); s.LC = sm + L::lc_off(0);
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
    static constexpr int WY = DD + N * NXI;          // P+ d (hoisted) on the way back, y on exit
    static constexpr int PV = WY + N * NXI;          // p_k
    static constexpr int KF = PV + N * NXI;          // feed-forward terms
    static constexpr int TV = KF + N * 4;            // 13 + pad
    static constexpr int DXI = TV + 16;
    static constexpr int TOTAL = DXI + 16;
    static constexpr size_t bytes() { return HEAD_BYTES + (size_t)TOTAL * sizeof(T); }
    // algorithmic words per problem: factor + rhs read once, solution written once
    static constexpr int ALGO_WORDS = N * (FAC_WORDS + 2 * (NZ + NXI));
};

__device__ __forceinline__ int pk(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

template <typename T, int N>
__global__ void __launch_bounds__(32) kkt_backsolve_kernel(const BacksolveParams<T> prm)
{
    using L = BsLayout<T, N>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x, b = blockIdx.x;
    if (b >= prm.B) return;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    T* sm = reinterpret_cast<T*>(smem_raw + L::HEAD_BYTES);
    T* FAC = sm + L::FAC; T* GZ = sm + L::GZ; T* DD = sm + L::DD; T* WY = sm + L::WY;
    T* PV = sm + L::PV; T* KF = sm + L::KF; T* TV = sm + L::TV; T* DXI = sm + L::DXI;

    constexpr uint32_t bytes_f = N * FAC_WORDS * sizeof(T), bytes_g = N * NZ * sizeof(T), bytes_d = N * NXI * sizeof(T);
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, bytes_f + bytes_g + bytes_d);
        tma_load(FAC, prm.fac + (size_t)b * N * FAC_WORDS, bytes_f, bar);
        tma_load(GZ, prm.g + (size_t)b * N * NZ, bytes_g, bar);
        tma_load(DD, prm.d + (size_t)b * N * NXI, bytes_d, bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);

    // hoisted, stage-parallel: w_k = P_{k+1} d_k   (k = 0..N-2)
    for (int e = lane; e < (N - 1) * NXI; e += 32) {
        const int k = e / NXI, i = e - k * NXI;
        const T* Pn = FAC + (k + 1) * FAC_WORDS;
        T acc = T(0);
#pragma unroll
        for (int j = 0; j < NXI; j++) acc += Pn[pk(i, j)] * DD[k * NXI + j];
        WY[k * NXI + i] = acc;
    }
    __syncwarp();

    // backward sweep: lane i (< 13) carries p_{k+1}[i] in a register
    T pnext = T(0);
    for (int k = N - 1; k >= 0; k--) {
        const T* fk = FAC + k * FAC_WORDS;
        const T* jc = fk + 153;
        if (lane < NXI) TV[lane] = (k < N - 1) ? pnext + WY[k * NXI + lane] : T(0);
        __syncwarp();
        T qz = T(0);   // q~ in z-ordering: g + J' tv
        if (lane < NZ) qz = GZ[k * NZ + lane] + ((k < N - 1) ? jt_y<T>(jc, TV, lane) : T(0));
        T qu[4];
#pragma unroll
        for (int r = 0; r < 4; r++) qu[r] = __shfl_sync(0xffffffffu, qz, r);
        if (lane < 4) {   // kff = -Quu^-1 q_u
            T acc = T(0);
#pragma unroll
            for (int c = 0; c < 4; c++) acc += fk[143 + pk(lane, c)] * qu[c];
            KF[k * 4 + lane] = -acc;
        }
        const T qxi = __shfl_sync(0xffffffffu, qz, e_col(lane < NXI ? lane : 0));
        if (lane < NXI) {   // p_k = q_xi + K' q_u
            T acc = qxi;
#pragma unroll
            for (int r = 0; r < 4; r++) acc += fk[91 + r * 13 + lane] * qu[r];
            pnext = acc;
            PV[k * NXI + lane] = acc;
        }
        __syncwarp();
    }

    // stage 0: x fixed, u_prev free: dq = -Pqq^-1 p_q
    {
        T a[16], l[10], li[4], x[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int c = 0; c <= r; c++) a[4 * r + c] = FAC[pk(9 + r, 9 + c)];
            x[r] = -PV[9 + r];
        }
        chol4<T>(a, l, li);
        fsub4<T>(l, li, x);
        bsub4<T>(l, li, x);
        if (lane < NXI) DXI[lane] = (lane < 9) ? T(0) : x[lane - 9];
    }
    __syncwarp();

    // forward sweep (dz overwrites g)
    for (int k = 0; k < N; k++) {
        const T* fk = FAC + k * FAC_WORDS;
        const T* jc = fk + 153;
        const int r = lane >> 3, part = lane & 7;
        T acc = T(0);
        for (int i = part; i < 13; i += 8) acc += fk[91 + r * 13 + i] * DXI[i];
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += KF[k * 4 + r];
        T v[13];
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] = __shfl_sync(0xffffffffu, acc, 8 * q);
#pragma unroll
        for (int q = 0; q < 9; q++) v[4 + q] = DXI[q];
        const T du_l = __shfl_sync(0xffffffffu, acc, 8 * (lane & 3));
        const T du_q = __shfl_sync(0xffffffffu, acc, 8 * ((lane - 9) & 3));
        if (lane < NZ) GZ[k * NZ + lane] = (lane < 4) ? du_l : (lane < 8 ? DXI[5 + lane] : DXI[lane - 8]);
        T nxt = T(0);
        if (k < N - 1 && lane < 13) {
            if (lane < 9) {
                nxt = DD[k * NXI + lane];
#pragma unroll
                for (int c = 0; c < 13; c++) nxt += f_dense<T>(jc, lane, c) * v[c];
            } else {
                nxt = du_q + DD[k * NXI + lane];
            }
        }
        __syncwarp();
        if (k < N - 1 && lane < 13) DXI[lane] = nxt;
        __syncwarp();
    }

    // hoisted, stage-parallel: y_k = P_k dxi_k + p_k   (k = 1..N-1),  y_0 = 0
    for (int e = lane; e < N * NXI; e += 32) {
        const int k = e / NXI, i = e - k * NXI;
        T acc = T(0);
        if (k > 0) {
            const T* Pk = FAC + k * FAC_WORDS;
            acc = PV[e];
#pragma unroll
            for (int j = 0; j < NXI; j++) acc += Pk[pk(i, j)] * GZ[k * NZ + e_col(j)];
        }
        WY[e] = acc;
    }
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
