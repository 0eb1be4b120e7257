// nmpc_prep.cuh -- the steps immediately before / after the solve, as device kernels (sm_100a).
//
// Batched, device-resident restatement of the body of FORCESNormal::solveNormal /
// FORCESFinal::solveFinal (/root/reference/src/resilient_planner/plan_manage/src/forces_normal.cpp:55-140)
// and of the receding-horizon bookkeeping around it (nmpc_solver.cpp:524-551):
//
//   pack_params_kernel       weights (setParasNormal, forces_normal.cpp:36-52), reference points,
//                            f_ext, yaw reference, polytope selection by poly_indices, corridor
//                            tightening b_j - ||E_i a_j||_2 (:124-125), truncation at the row
//                            capacity (the reference drops rows beyond 30, :114)  -> hdr / rows / nrows
//   shift_warm_start_kernel  x0[i] <- previous[i+1], last stage duplicated (nmpc_solver.cpp:543),
//                            xinit <- previous[1][8:17] (forces_normal.cpp:62-97), yaw wrapped to
//                            (-pi, pi] as updateFORCESResults does (nmpc_solver.cpp:531-541)
//
// Both are pure data movement plus a few flops per row: one thread per (problem, stage), reads
// and writes coalesced along the stage-major innermost dimension.
#pragma once
#include <cuda_runtime.h>

namespace nmpc {

struct PackParams {
    int B, N, P, M, mcap;
    const double* ref_pos;    // [B][N][3]
    const double* ref_yaw;    // [B][N]
    const double* ext_acc;    // [B][3]
    const double* ellipsoid;  // [B][N][9]   E_i, row-major 3x3
    const double* poly_A;     // [B][P][M][3]
    const double* poly_b;     // [B][P][M]
    const int* poly_m;        // [B][P]      live rows of each polytope
    const int* poly_idx;      // [B][N]      polytope used by stage i
    double w_stage_wp, w_stage_input, w_input_rate, w_terminal_wp, w_terminal_input;
    double* hdr;              // [B][N][10]
    double* rows;             // [B][N][mcap][4]
    int* nrows;               // [B][N]
};

__global__ void pack_params_kernel(const PackParams q)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= q.B * q.N) return;
    const int b = t / q.N, i = t - b * q.N;
    double* h = q.hdr + (size_t)t * 10;
    const bool terminal = (i == q.N - 1);
    h[0] = q.ref_pos[(size_t)t * 3 + 0]; h[1] = q.ref_pos[(size_t)t * 3 + 1]; h[2] = q.ref_pos[(size_t)t * 3 + 2];
    h[3] = q.ext_acc[b * 3 + 0]; h[4] = q.ext_acc[b * 3 + 1]; h[5] = q.ext_acc[b * 3 + 2];
    h[6] = terminal ? q.w_terminal_wp : q.w_stage_wp;
    h[7] = terminal ? q.w_terminal_input : q.w_stage_input;
    h[8] = q.w_input_rate;
    h[9] = q.ref_yaw[t];
    const int pi = q.poly_idx[t];
    const int m_poly = q.poly_m[b * q.P + pi];
    const int m = m_poly < q.mcap ? m_poly : q.mcap;
    const double* E = q.ellipsoid + (size_t)t * 9;
    const double* A = q.poly_A + ((size_t)b * q.P + pi) * q.M * 3;
    const double* bb = q.poly_b + ((size_t)b * q.P + pi) * q.M;
    double* r = q.rows + (size_t)t * q.mcap * 4;
    for (int j = 0; j < q.mcap; j++) {
        if (j < m) {
            const double a0 = A[3 * j], a1 = A[3 * j + 1], a2 = A[3 * j + 2];
            const double e0 = E[0] * a0 + E[1] * a1 + E[2] * a2;
            const double e1 = E[3] * a0 + E[4] * a1 + E[5] * a2;
            const double e2 = E[6] * a0 + E[7] * a1 + E[8] * a2;
            r[4 * j] = a0; r[4 * j + 1] = a1; r[4 * j + 2] = a2;
            r[4 * j + 3] = bb[j] - sqrt(e0 * e0 + e1 * e1 + e2 * e2);
        } else {
            r[4 * j] = r[4 * j + 1] = r[4 * j + 2] = r[4 * j + 3] = 0.0;
        }
    }
    q.nrows[t] = m;
}

__global__ void shift_warm_start_kernel(int B, int N, const double* z_prev, double* xinit, double* z0, int wrap_yaw)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * N) return;
    const int b = t / N, i = t - b * N;
    const int src = (i + 1 < N) ? i + 1 : N - 1;
    const double* zs = z_prev + ((size_t)b * N + src) * 17;
    double* zd = z0 + (size_t)t * 17;
    const double pi = 3.14159265358979323846;
    for (int j = 0; j < 17; j++) {
        double v = zs[j];
        if (j == 16 && wrap_yaw) v = v < -pi ? v + 2 * pi : (v > pi ? v - 2 * pi : v);
        zd[j] = v;
    }
    if (i == 0)
        for (int j = 0; j < 9; j++) xinit[(size_t)b * 9 + j] = zd[8 + j];
}

}  // namespace nmpc
