// nmpc_ellipsoid.cuh -- disturbance-ellipsoid propagation along the previous plan (sm_100a).
//
// Batched, device-resident form of the ellipsoid part of NMPCSolver::setFORCESParams
// (/root/reference/src/resilient_planner/plan_manage/src/nmpc_solver.cpp:484-521) with the functions it
// calls per stage: updateMatrix (:615-699, linearisation A_t, B_t and the closed loop
// Phi = A_t + B_t K_t), eulerToRot (:552-564), getDistrEllipsoid (:567-611) and the 3x3 matrix square
// root (:511-512).  Output: the N shape matrices E_i (ellipsoid_matrices_) that solveNormal uses to
// tighten the corridor rows (forces_normal.cpp:124-125) -- i.e. the `ellipsoid` input of
// pack_params_kernel, so shift -> propagate -> pack -> solve stays on the device.
//
// Not a port.  The reference solves, per stage and disturbance direction i, the Sylvester equation
//     Phi X + X Phi' = N_i - exp(-Phi t) N_i exp(-Phi' t),   N_i = t w_i^2 d_i d_i'
// by complex Schur forms (Bartels-Stewart) and takes two Pade matrix exponentials.  Its solution is
// the finite-horizon Gramian  X = int_0^t exp(-Phi s) N_i exp(-Phi' s) ds = t w_i^2 int_0^t v(s) v(s)' ds,
// v(s) = exp(-Phi s) d_i  (N_i has rank one).  Here: Krylov vectors u_j = (-Phi t)^j d_i / j!  (K = 20
// terms reach fp64 round-off at ||Phi t|| ~ 1.3), v at the 12 Gauss-Legendre nodes of [0, t] as
// V = U [xi_q^j], and X = t^2 w_i^2 sum_q (omega_q / 2) v_q v_q'  (the rule is exact for the series product
// up to degree 23; the remainder is below 1e-14).  Only the first three rows of exp(Phi t) are needed
// (the position block of exp(Phi t) Q exp(Phi t)'): three more row-vector series instead of a matrix
// exponential.  Everything is 9x9 matrix-vector work plus two small products, which one warp per agent
// does out of shared memory.  oracle/ellipsoid_np.py restates the reference literally
// (Schur/Sylvester, Pade); the two agree to ~1e-13 relative.
//
// Two deliberate deviations, in both (each a latent bug of the reference, restated as what the formula means):
//   1. the reference accumulates `temp += sqrt(X.trace())` into an uninitialised double (:573, :597 -- undefined
//      behaviour); here temp starts at 0;
//   2. updateMatrix never ASSIGNS At_(5,8) -- the thrust term of d(acc_z)/d(yaw) is zero, so there is no `At_(5,8) = ...`
//      line next to :643-644 -- it only does `At_(5,8) += ...` (:689) on the class member At_, so in the reference that
//      entry keeps growing over all 20 stages of a replan and over every replan; here the linearisation of a stage is
//      built from zero (the drag term alone), as for every other entry.
//
// The stages are a recurrence (Q_init and Q2 carry over), so a warp walks its agent's horizon with the lanes
// splitting the matrix entries; what does not take part in the recurrence is done for all stages at
// once with one stage per lane: the linearisation scalars before the walk, the 3x3 square roots after it.
#pragma once
#include <cuda_runtime.h>

namespace nmpc {

struct EllipsoidParams {
    int B, N;
    const double* z;        // [B][N][17]  previous plan (mpc_output_): thrust z[3], vel z[11:14], rpy z[14:17]
    double* ellipsoid;      // [B][N][9]   E_i row-major
    double mass, drag, ego_r, ego_h, ext_noise_bound, epsilon, Ts;
};

// feedback gain K_t of the ancillary controller (nmpc_solver.cpp:28-31), row-major 4x9
__device__ const double ELL_KT[36] = {
    -2.0, 5.0, 0.0, -1.0, 4.0, 0.0, -8.0, 0.0, 0.0,
    -5.0, -2.0, 0.0, -4.0, -1.0, 0.0, 0.0, -8.0, 0.0,
    -2.0, -2.0, 0.0, -1.0, -1.0, 0.0, 0.0, 0.0, -8.0,
    0.0, 0.0, -8.0, 0.0, 0.0, -6.0, 0.0, 0.0, 0.0};

constexpr int ELL_K = 20;              // series terms: ||Phi t|| ~ 1.3, 1.3^20 / 20! ~ 1e-16
constexpr int ELL_KP = ELL_K + 1;
constexpr int ELL_Q = 12;              // Gauss-Legendre nodes on [0, t]: exact to degree 23 of the series product
constexpr int ELL_WARPS = 2;           // agents per CTA (14.1 KB + 312 N bytes of shared memory each)
static_assert(ELL_KP % 3 == 0 && ELL_Q % 2 == 0, "unrolled accumulation chains");

// nodes xi_q in (0, 1) and sqrt(weight_q / 2) of the 12-point Gauss-Legendre rule
__device__ const double ELL_XI[ELL_Q] = {
    0.0092196828766403782, 0.047941371814762601, 0.11504866290284765, 0.20634102285669126, 0.31608425050090994,
    0.43738329574426554, 0.5626167042557344, 0.68391574949909006, 0.79365897714330869, 0.88495133709715235,
    0.95205862818523745, 0.99078031712335957};
__device__ const double ELL_SW[ELL_Q] = {
    0.15358277310055221, 0.23123508167589868, 0.28291193730854342, 0.31872200012163088, 0.3416815304771057,
    0.35294974558242898, 0.35294974558242898, 0.3416815304771057, 0.31872200012163088, 0.28291193730854342,
    0.23123508167589868, 0.15358277310055221};

// symmetric 3x3 -> principal square root, cyclic Jacobi in registers (E = V sqrt(L) V')
__device__ __forceinline__ void sqrtm3_sym(const double q[9], double e[9])
{
    double a[3][3] = {{q[0], q[1], q[2]}, {q[3], q[4], q[5]}, {q[6], q[7], q[8]}};
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
#pragma unroll 1
    for (int sweep = 0; sweep < 6; sweep++) {
        // converged (quadratically) once the off-diagonal mass is below round-off of the trace
        if (fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]) <= 1e-17 * (fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]))) break;
#pragma unroll
        for (int pq = 0; pq < 3; pq++) {
            const int p = pq == 2 ? 1 : 0, r = pq == 0 ? 1 : 2;
            const double apq = a[p][r];
            if (fabs(apq) > 1e-300) {
                const double theta = (a[r][r] - a[p][p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
                for (int k = 0; k < 3; k++) {   // A <- A J
                    const double akp = a[k][p], akr = a[k][r];
                    a[k][p] = c * akp - s * akr; a[k][r] = s * akp + c * akr;
                }
#pragma unroll
                for (int k = 0; k < 3; k++) {   // A <- J' A
                    const double apk = a[p][k], ark = a[r][k];
                    a[p][k] = c * apk - s * ark; a[r][k] = s * apk + c * ark;
                }
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double vkp = v[k][p], vkr = v[k][r];
                    v[k][p] = c * vkp - s * vkr; v[k][r] = s * vkp + c * vkr;
                }
            }
        }
    }
    const double l0 = sqrt(fmax(a[0][0], 0.0)), l1 = sqrt(fmax(a[1][1], 0.0)), l2 = sqrt(fmax(a[2][2], 0.0));
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) e[3 * i + j] = v[i][0] * l0 * v[j][0] + v[i][1] * l1 * v[j][1] + v[i][2] * l2 * v[j][2];
}

constexpr int ELL_SCW = 30;            // per-stage scalars: a[9] = At(3:6, 6:9) | RDR'[9] = At(3:6, 3:6) | bt[3] = Bt(3:6, 3) | Q1[9]
constexpr int ELL_FIXED = 243 + 27 * ELL_KP + 27 * ELL_Q + 243 + ELL_KP * ELL_Q + 81 + ELL_KP + 1 + 32;   // doubles per warp besides the per-stage tables
__host__ __device__ constexpr size_t ellipsoid_smem_bytes(int N) { return (size_t)ELL_WARPS * (ELL_FIXED + N * (ELL_SCW + 9)) * sizeof(double); }

__global__ void __launch_bounds__(32 * ELL_WARPS) ellipsoid_propagate_kernel(const EllipsoidParams q)
{
    // per warp: A = Phi t | Q_origin | Qd | U (3 x 9 x KP) | V (3 x 9 x Q) | X (3 x 81) | node powers sqrt(w_q/2) xi_q^j
    // | first three rows of exp(A) (3 x 9) | row-series terms (2 x 27) | 1/k | scratch | per-stage scalars [N][30] | Q_i [N][9]
    constexpr int O_A = 0, O_QO = 81, O_TQ = 162, O_U = 243, O_V = O_U + 27 * ELL_KP, O_X = O_V + 27 * ELL_Q,
                  O_PW = O_X + 243, O_ER = O_PW + ELL_KP * ELL_Q, O_Y0 = O_ER + 27, O_Y1 = O_Y0 + 27, O_INV = O_Y1 + 27,
                  O_SC = O_INV + ELL_KP + 1, O_SCA = O_SC + 32;
    static_assert(O_SCA == ELL_FIXED, "shared-memory plan");
    extern __shared__ double ell_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.x * ELL_WARPS + wid;
    if (b >= q.B) return;
    double* S = ell_smem + (size_t)wid * (ELL_FIXED + q.N * (ELL_SCW + 9));
    double *A = S + O_A, *QO = S + O_QO, *TQ = S + O_TQ, *U = S + O_U, *V = S + O_V, *X = S + O_X, *PW = S + O_PW;
    double *ER = S + O_ER, *INV = S + O_INV, *SC = S + O_SC, *SCA = S + O_SCA, *QS = SCA + q.N * ELL_SCW;
    const double t = q.Ts;

    for (int e = lane; e < ELL_KP; e += 32) INV[e] = 1.0 / (double)(e + 1);
    for (int e = lane; e < ELL_Q; e += 32) {       // PW[j][q] = sqrt(w_q / 2) xi_q^j
        double pw = ELL_SW[e];
        for (int j = 0; j < ELL_KP; j++) { PW[j * ELL_Q + e] = pw; pw *= ELL_XI[e]; }
    }
    for (int e = lane; e < 81; e += 32) QO[e] = (e / 9 == e % 9) ? q.epsilon * q.epsilon : 0.0;   // Q_init = eps^2 I (:487)
    // ---- updateMatrix + eulerToRot for every stage at once (lane = stage): they depend on the plan only ----
    for (int i = lane; i < q.N; i += 32) {
        const double* zi = q.z + ((size_t)b * q.N + i) * 17;
        const double thrust = zi[3], v1 = zi[11], v2 = zi[12], v3 = zi[13], roll = zi[14], pitch = zi[15], yaw = zi[16];
        double sr, cr, sp, cp, sy, cy;
        sincos(roll, &sr, &cr); sincos(pitch, &sp, &cp); sincos(yaw, &sy, &cy);
        // R = Rz Ry Rx (eulerToRot)
        const double R[9] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
                             sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                             -sp, cp * sr, cp * cr};
        double* sc = SCA + i * ELL_SCW;
        const double comb0 = thrust * 1.0 / q.mass, drag = q.drag;
        const double comb5 = cp * sp, comb6 = cp * sr, comb7 = cp * cr, comb8 = sp * cr, comb9 = sp * sr;
        const double comb1 = cr * sy - comb9 * cy, comb2 = sr * cy - comb8 * sy;
        const double comb3 = cr * cy + comb9 * sy, comb4 = sr * sy + comb8 * cy;
        const double cp2 = cp * cp, sp2 = sp * sp, cy2 = cy * cy, sy2 = sy * sy, sr2 = sr * sr;
        const double t10 = comb6 * comb4 - comb7 * comb1, t11 = comb3 * comb4 + comb1 * comb2, t12 = comb6 * comb2 - comb7 * comb3;
        const double t20 = cy * (sp2 - cp2 + cp2 * sr2) + comb9 * comb1;
        const double t21 = 2 * comb5 * cy * sy - comb6 * (cy * comb3 + sy * comb1);
        const double t22 = sy * (cp2 - sp2 - cp2 * sr2) + comb9 * comb3;
        const double t30 = 2 * drag * (comb3 * comb1 - cp2 * cy * sy), t31 = drag * (comb6 * comb3 - comb5 * sy);
        const double t32 = drag * (comb3 * comb3 - comb1 * comb1 - cp2 * cy2 + cp2 * sy2), t33 = drag * (comb6 * comb1 + comb5 * cy);
        // a[r][c]: r = vel row (3..5), c = roll / pitch / yaw column (6..8)
        sc[0] = comb0 * comb1 + drag * (v3 * t10 + v2 * t11 - 2 * v1 * comb4 * comb1);
        sc[3] = -comb0 * comb3 + drag * (v1 * t11 - v3 * t12 - 2 * v2 * comb3 * comb2);
        sc[6] = -comb0 * comb6 + drag * (v1 * t10 - v2 * t12 + 2 * v3 * comb7 * comb6);
        sc[1] = comb0 * comb7 * cy + drag * (v3 * t20 - v2 * t21 - v1 * 2 * (comb5 * cy2 + comb6 * comb1 * cy));
        sc[4] = comb0 * comb7 * sy - drag * (v3 * t22 - v1 * t21 - v2 * 2 * (comb5 * sy2 - comb6 * comb3 * sy));
        sc[7] = -comb0 * comb8 + drag * (v1 * t20 - v2 * t22 + v3 * 2 * (comb5 - comb5 * sr2));
        sc[2] = comb0 * comb2 + (v1 * t30 - v3 * t31 - v2 * t32);
        sc[5] = comb0 * comb4 + (-v1 * t32 - v3 * t33 - v2 * t30);
        sc[8] = -v2 * t33 - v1 * t31;
        const double er = q.ego_r * q.ego_r, eh = q.ego_h * q.ego_h;
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                sc[9 + 3 * r + c] = drag * (R[3 * r] * R[3 * c] + R[3 * r + 1] * R[3 * c + 1]);   // R diag(d,d,0) R'
                sc[21 + 3 * r + c] = er * (R[3 * r] * R[3 * c] + R[3 * r + 1] * R[3 * c + 1]) + eh * R[3 * r + 2] * R[3 * c + 2];   // Q1 = R ego R' (:503)
            }
        sc[18] = comb4 / q.mass; sc[19] = -comb2 / q.mass; sc[20] = comb7 / q.mass;
    }
    double q2[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    __syncwarp();

    for (int i = 0; i < q.N; i++) {
        // ---- Q_i = Q1 (+) Q2 (:503-510); its square root is taken for all stages at once after the loop ----
        {
            const double* q1 = SCA + i * ELL_SCW + 21;
            double qq[9];
            if (i == 0) {
#pragma unroll
                for (int e = 0; e < 9; e++) qq[e] = q1[e];
            } else {
                const double beta = sqrt((q1[0] + q1[4] + q1[8]) / (q2[0] + q2[4] + q2[8]));
#pragma unroll
                for (int e = 0; e < 9; e++) qq[e] = (1.0 + 1.0 / beta) * q1[e] + (1.0 + beta) * q2[e];
            }
            if (lane == 0) {
#pragma unroll
                for (int e = 0; e < 9; e++) QS[i * 9 + e] = qq[e];
            }
        }
        {   // this stage's linearisation scalars
            const double* sc = SCA + i * ELL_SCW;
            if (lane < 21) SC[lane] = sc[lane];
        }
        __syncwarp();
        // ---- A = Phi t = (At + Bt Kt) t ----
        for (int e = lane; e < 81; e += 32) {
            const int r = e / 9, c = e - 9 * r;
            double v = 0.0;
            if (r < 3) v = (c == r + 3) ? 1.0 : 0.0;
            else if (r < 6) v = (c >= 6 ? SC[3 * (r - 3) + c - 6] : (c >= 3 ? SC[9 + 3 * (r - 3) + c - 3] : 0.0)) + SC[18 + r - 3] * ELL_KT[27 + c];
            else v = ELL_KT[9 * (r - 6) + c];
            A[e] = v * t;
        }
        // seeds: u_0 = d_i = e_{3+i} (U[d][row][0]); row r of exp(A): y_0 = e_r' (lanes 0..26 = (d or r, column))
        double erow = 0.0;
        if (lane < 27) {
            const int d = lane / 9, r = lane - 9 * d;
            U[lane * ELL_KP] = (r == 3 + d) ? 1.0 : 0.0;
            erow = (r == d) ? 1.0 : 0.0;
            S[O_Y0 + lane] = erow;
        }
        __syncwarp();
        // ---- u_j = -A u_{j-1} / j  and  y_j = y_{j-1} A / j  (exp(A)[r, :] = sum_j y_j), j = 1..K ----
        for (int j = 1, cur = O_Y0, nxt = O_Y1; j <= ELL_K; j++) {
            if (lane < 27) {
                const int d = lane / 9, r = lane - 9 * d;
                double au = 0.0, ya = 0.0;
#pragma unroll
                for (int m = 0; m < 9; m++) {
                    au += A[9 * r + m] * U[(9 * d + m) * ELL_KP + j - 1];
                    ya += S[cur + 9 * d + m] * A[9 * m + r];
                }
                const double ij = INV[j - 1];
                U[lane * ELL_KP + j] = -au * ij;
                ya *= ij;
                S[nxt + lane] = ya; erow += ya;
            }
            __syncwarp();
            const int tmp = cur; cur = nxt; nxt = tmp;
        }
        if (lane < 27) ER[lane] = erow;
        // ---- V = U PW (27 x Q): v(s_q) sqrt(w_q / 2);  X_d = t^2 w^2 V_d V_d' ----
        for (int e = lane; e < 27 * ELL_Q; e += 32) {
            const int row = e / ELL_Q, qq = e - ELL_Q * row;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
            for (int j = 0; j < ELL_KP; j += 3) {
                a0 += U[row * ELL_KP + j] * PW[j * ELL_Q + qq];
                a1 += U[row * ELL_KP + j + 1] * PW[(j + 1) * ELL_Q + qq];
                a2 += U[row * ELL_KP + j + 2] * PW[(j + 2) * ELL_Q + qq];
            }
            V[e] = (a0 + a1) + a2;
        }
        __syncwarp();
        const double tw = t * t * q.ext_noise_bound * q.ext_noise_bound;
        for (int e = lane; e < 243; e += 32) {
            const int d = e / 81, rc = e - 81 * d, r = rc / 9, c = rc - 9 * r;
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int l = 0; l < ELL_Q; l += 2) {
                a0 += V[(9 * d + r) * ELL_Q + l] * V[(9 * d + c) * ELL_Q + l];
                a1 += V[(9 * d + r) * ELL_Q + l + 1] * V[(9 * d + c) * ELL_Q + l + 1];
            }
            X[e] = tw * (a0 + a1);
        }
        __syncwarp();
        // ---- Qd = temp * sum_d X_d / sqrt(tr X_d), Q_update (:595-603) ----
        double st[3], temp = 0.0;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            double tr = 0.0;
#pragma unroll
            for (int m = 0; m < 9; m++) tr += X[81 * d + 10 * m];
            st[d] = sqrt(tr); temp += st[d];
        }
        for (int e = lane; e < 81; e += 32) TQ[e] = temp * (X[e] / st[0] + X[81 + e] / st[1] + X[162 + e] / st[2]);
        __syncwarp();
        {
            double tro = 0.0, trd = 0.0;
#pragma unroll
            for (int m = 0; m < 9; m++) { tro += QO[10 * m]; trd += TQ[10 * m]; }
            const double beta = sqrt(tro / trd);
            __syncwarp();
            for (int e = lane; e < 81; e += 32) QO[e] = (1.0 + 1.0 / beta) * QO[e] + (1.0 + beta) * TQ[e];
        }
        __syncwarp();
        // ---- Q2 = (exp(A) Q_update exp(A)')[0:3, 0:3] (:604-610) ----
        if (lane < 27) {
            const int r = lane / 9, c = lane - 9 * r;
            double acc = 0.0;
#pragma unroll
            for (int m = 0; m < 9; m++) acc += ER[9 * r + m] * QO[9 * m + c];
            SC[lane] = acc;                       // T1 = exp(A)[0:3, :] Q_update  (3 x 9)
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < 9; m++) acc += SC[9 * r + m] * ER[9 * c + m];
                q2[3 * r + c] = acc;
            }
        __syncwarp();
    }
    // ---- E_i = sqrtm(Q_i) (:511-512), one stage per lane ----
    for (int i = lane; i < q.N; i += 32) {
        double qq[9], ee[9];
#pragma unroll
        for (int e = 0; e < 9; e++) qq[e] = QS[i * 9 + e];
        sqrtm3_sym(qq, ee);
        double* out = q.ellipsoid + ((size_t)b * q.N + i) * 9;
#pragma unroll
        for (int e = 0; e < 9; e++) out[e] = ee[e];
    }
}

}  // namespace nmpc
