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
//   sample_reference_kernel  reference point and yaw reference of every stage from the front-end's
//                            polyline (NMPCSolver::getCurTraj, nmpc_solver.cpp:109-142, and
//                            calculate_yaw, :834-862): linear interpolation at Ts, look-ahead point
//                            five samples on, heading with unwrap against the previous value and the
//                            0.2 / 0.8 low-pass -- a recurrence over the stages, so one thread per
//                            agent walks its horizon
//
// The first two are pure data movement plus a few flops per row: one thread per (problem, stage),
// reads and writes coalesced along the stage-major innermost dimension.
#pragma once
#include <cuda_runtime.h>

namespace nmpc {

// the reference's own constant (nmpc_solver.cpp:3, `constexpr double PI = 3.1415926;`): its yaw wrap and
// unwrap use this truncated value, so results are only identical to the reference's with it
constexpr double REF_PI = 3.1415926;

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
    // a polytope index outside [0, P) or a row count beyond the allocated M would read past poly_A / poly_b and
    // hand the solver garbage as safety constraints: clamp both (the reference's own truncation is at 30 rows, :114)
    int pi = q.poly_idx[t];
    pi = pi < 0 ? 0 : (pi >= q.P ? q.P - 1 : pi);
    int m = q.poly_m[b * q.P + pi];
    m = m < 0 ? 0 : m;
    m = m < q.M ? m : q.M;
    m = m < q.mcap ? m : q.mcap;
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
    const double pi = REF_PI;
    for (int j = 0; j < 17; j++) {
        double v = zs[j];
        if (j == 16 && wrap_yaw) v = v < -pi ? v + 2 * pi : (v > pi ? v - 2 * pi : v);
        zd[j] = v;
    }
    if (i == 0)
        for (int j = 0; j < 9; j++) xinit[(size_t)b * 9 + j] = zd[8 + j];
}

// updateFORCESResults' yaw wrap (nmpc_solver.cpp:531-541) on an adopted plan, in place
__global__ void wrap_yaw_kernel(int n_stages_total, double* z)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_stages_total) return;
    const double v = z[(size_t)t * 17 + 16];
    z[(size_t)t * 17 + 16] = v < -REF_PI ? v + 2 * REF_PI : (v > REF_PI ? v - 2 * REF_PI : v);
}

// Result handling of NMPCSolver::solveNMPC (nmpc_solver.cpp:398-427, 363-364): the new plan replaces the previous
// one only when it is accepted (exit flag 1, or the caller's acceptance mask -- the "tolerate MAXIT after more than 3
// replans" rule is host policy, SolveAcceptance); a rejected agent keeps nothing of the failed solve (its output may be
// NaN) and cold-starts on the next cycle exactly as initMPCOutput does (:265-286): every stage = [0,0,0,7.3, 0,0,0,7.3,
// x] with x its current state -- `odom` when the caller has one, otherwise where the previous plan puts the vehicle one
// period later (its stage 2, the plan having been adopted one period ago).  Adopted plans are stored yaw-wrapped
// (updateFORCESResults, :531-541).  One thread per (agent, stage).
// Launch with blockDim.x = N * (agents per block): an agent never straddles two blocks, so one barrier separates every
// read of its old plan from the writes that replace it.
__global__ void adopt_plans_kernel(int B, int N, const double* z_new, const int* info_int, const int* accept, const double* odom,
                                   double* z_prev, int* cold, int wrap_yaw)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < B * N;
    const int b = live ? t / N : 0, i = live ? t - b * N : 0;
    const bool ok = accept ? accept[b] != 0 : info_int[(size_t)b * 4] == 1;
    double x[9];
    if (live && !ok) {
        const double* src = odom ? odom + (size_t)b * 9 : z_prev + ((size_t)b * N + (N > 2 ? 2 : N - 1)) * 17 + 8;
        for (int j = 0; j < 9; j++) x[j] = src[j];
    }
    __syncthreads();
    if (!live) return;
    double* zp = z_prev + (size_t)t * 17;
    if (ok) {
        const double* zn = z_new + (size_t)t * 17;
        for (int j = 0; j < 16; j++) zp[j] = zn[j];
        const double v = zn[16];
        zp[16] = !wrap_yaw ? v : (v < -REF_PI ? v + 2 * REF_PI : (v > REF_PI ? v - 2 * REF_PI : v));
    } else {
        zp[0] = zp[1] = zp[2] = 0.0; zp[3] = 7.3; zp[4] = zp[5] = zp[6] = 0.0; zp[7] = 7.3;
        for (int j = 0; j < 9; j++) zp[8 + j] = x[j];
    }
    if (i == 0 && cold) cold[b] = ok ? 0 : 1;
}

// order[r] = index of the agent with the r-th largest key, key = iterations of the last solve (+1000 for a failed one),
// ties by agent index: the launch order of the next warm solve (longest first, see nmpc_solve_batch_ordered_f64).
// Rank by counting: every CTA stages all B keys in shared memory and ranks its own 128 agents against them
// (B compares per thread, B / 128 CTAs side by side: 1024 agents in ~5 us); no library sort on the replan path.
constexpr int RANK_THREADS = 128;
__global__ void rank_longest_first_kernel(int B, const int* info_int, int* order)
{
    extern __shared__ int keys[];
    for (int b = threadIdx.x; b < B; b += blockDim.x)
        keys[b] = info_int[(size_t)b * 4 + 1] + (info_int[(size_t)b * 4] != 1 ? 1000 : 0);
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int kb = keys[b];
    int rank = 0;
#pragma unroll 4
    for (int c = 0; c < B; c++) {
        const int kc = keys[c];
        rank += (kc > kb) || (kc == kb && c < b);
    }
    order[rank] = b;
}

struct SampleParams {
    int B, N, P;
    double Ts;
    const double* kino_path;   // [B][P][3]  front-end path sampled at Ts (kino_path_, nmpc_solver.cpp:215)
    const int* kino_size;      // [B]        live points of each path (>= 1)
    const double* t_off;       // [B]        mpc_start_time_ - kino_start_time_ in seconds (:111)
    const double* last_yaw;    // [B]        yaw of the previous plan's stage 1 (setFORCESParams, :486)
    const double* pos1;        // [B][3] or nullptr: position of the previous plan's stage 1 (:136)
    double* ref_pos;           // [B][N][3]
    double* ref_yaw;           // [B][N]
    int* hard_to_follow;       // [B] or nullptr: 1 when stage 0's reference is > 1 m from pos1 (kino_replan_, :136-140)
};

__global__ void sample_reference_kernel(const SampleParams q)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= q.B) return;
    const double* path = q.kino_path + (size_t)b * q.P * 3;
    const int size = q.kino_size[b];
    const double toff = q.t_off[b];
    const double pi = REF_PI;
    double last = q.last_yaw[b];
    for (int i = 0; i < q.N; i++) {
        const double index_time = i * q.Ts + toff;
        const unsigned ki = (unsigned)(int)(index_time / q.Ts);
        double r[3], f[3];
        if (ki + 1 < (unsigned)size) {
            const double w = fmod(index_time, q.Ts) / q.Ts;
            for (int c = 0; c < 3; c++) r[c] = path[ki * 3 + c] + w * (path[(ki + 1) * 3 + c] - path[ki * 3 + c]);
        } else {
            for (int c = 0; c < 3; c++) r[c] = path[(size - 1) * 3 + c];
        }
        const unsigned kf = (ki + 5 < (unsigned)size) ? ki + 5 : (unsigned)(size - 1);
        for (int c = 0; c < 3; c++) f[c] = path[kf * 3 + c];
        // calculate_yaw
        const double dx = f[0] - r[0], dy = f[1] - r[1], dz = f[2] - r[2];
        const double yaw_temp = sqrt(dx * dx + dy * dy + dz * dz) > 0.1 ? atan2(dy, dx) : last;
        double yaw = yaw_temp;
        if (fabs(yaw_temp - last) > pi) yaw = yaw_temp > 0 ? yaw_temp - 2 * pi : yaw_temp + 2 * pi;
        yaw = 0.2 * last + 0.8 * yaw;
        last = yaw;
        double* o = q.ref_pos + ((size_t)b * q.N + i) * 3;
        o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
        q.ref_yaw[(size_t)b * q.N + i] = yaw;
        if (i == 0 && q.hard_to_follow) {
            int far = 0;
            if (q.pos1) {
                const double ex = r[0] - q.pos1[b * 3], ey = r[1] - q.pos1[b * 3 + 1], ez = r[2] - q.pos1[b * 3 + 2];
                far = sqrt(ex * ex + ey * ey + ez * ez) > 1.0;
            }
            q.hard_to_follow[b] = far;
        }
    }
}

}  // namespace nmpc
