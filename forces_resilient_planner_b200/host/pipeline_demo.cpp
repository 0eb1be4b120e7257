// pipeline_demo.cpp -- one fleet replan cycle, device-resident, from C++ through the C ABI only
// (include/nmpc_b200.h): what NMPCSolver::setFORCESParams + FORCESNormal::solveNormal do per cycle
// (nmpc_solver.cpp:484-521, forces_normal.cpp:55-140), for B vehicles at once:
//   shift -> ellipsoids -> references -> corridors -> parameters -> solve, three cycles in closed loop.
// The front end (polyline, obstacle cloud) is synthetic.  Exit status 0 iff every solve returned 1
// and no corridor overflowed.  Build: see host/Makefile (g++ + libcudart for the device buffers).
#include <cmath>
#include <cstdio>
#include <vector>

#include <cuda_runtime.h>

#include "forces_wrappers.hpp"

namespace {
template <class T> T* dev(size_t n) { void* p = nullptr; cudaMalloc(&p, n * sizeof(T)); cudaMemset(p, 0, n * sizeof(T)); return (T*)p; }
template <class T> T* dev(const std::vector<T>& h) { T* p = dev<T>(h.size()); cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice); return p; }
#define CHECK(call) do { int rc_ = (call); if (rc_ != 0) { std::printf("%s -> %d: %s\n", #call, rc_, nmpc_last_error()); return 100; } } while (0)
}  // namespace

int main()
{
    const int B = 64, N = 20, P = 60, M = 256, NP = 20, R = 30, mcap = 30;
    const double Ts = 0.05;
    // front end: a circle arc of radius 3 m flown at 1 m/s per vehicle (all inside the +-20 m position bounds of the
    // NLP, mpc_generator_normal.m:33-46), scattered obstacle points outside a 1 m tube around it
    std::vector<double> path((size_t)B * P * 3), cloud((size_t)B * M * 3), xinit((size_t)B * 9, 0.0), z0((size_t)B * N * 17, 0.0);
    std::vector<int> size(B, P), cloud_n(B, 0);
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) * (1.0 / 16777216.0); };
    for (int b = 0; b < B; b++) {
        const double cx = -15.0 + 4.0 * (b % 8), cy = -15.0 + 4.0 * (b / 8), r0 = 3.0, phase = 0.09 * b;
        for (int k = 0; k < P; k++) {
            const double a = phase + 0.05 * k / r0;
            path[((size_t)b * P + k) * 3 + 0] = cx + r0 * std::sin(a); path[((size_t)b * P + k) * 3 + 1] = cy + r0 * (1 - std::cos(a));
            path[((size_t)b * P + k) * 3 + 2] = 1.0;
        }
        int n = 0;
        for (int t = 0; t < 4 * M && n < M; t++) {
            const double x = cx - 6 + 12 * rnd(), y = cy - 3 + 12 * rnd(), z = 2.2 * rnd();
            double dmin = 1e9;
            for (int k = 0; k < P; k += 3) dmin = std::fmin(dmin, std::hypot(x - path[((size_t)b * P + k) * 3], y - path[((size_t)b * P + k) * 3 + 1]));
            if (dmin > 1.0) { cloud[((size_t)b * M + n) * 3] = x; cloud[((size_t)b * M + n) * 3 + 1] = y; cloud[((size_t)b * M + n) * 3 + 2] = z; n++; }
        }
        cloud_n[b] = n;
        double* x0 = &xinit[(size_t)b * 9];
        x0[0] = path[(size_t)b * P * 3]; x0[1] = path[(size_t)b * P * 3 + 1]; x0[2] = 1.0; x0[8] = phase;
        for (int k = 0; k < N; k++) {                     // initMPCOutput (nmpc_solver.cpp:265-286)
            double* z = &z0[((size_t)b * N + k) * 17];
            z[3] = z[7] = 7.3;
            for (int i = 0; i < 9; i++) z[8 + i] = x0[i];
        }
    }
    double *d_path = dev(path), *d_cloud = dev(cloud), *d_xinit = dev(xinit), *d_z0 = dev(z0), *d_zprev = dev(z0);
    int *d_size = dev(size), *d_cn = dev(cloud_n);
    double *d_toff = dev<double>(B), *d_lastyaw = dev<double>(B), *d_pos1 = dev<double>((size_t)B * 3), *d_ext = dev<double>((size_t)B * 3);
    double *d_E = dev<double>((size_t)B * N * 9), *d_ref = dev<double>((size_t)B * N * 3), *d_yaw = dev<double>((size_t)B * N);
    double *d_pA = dev<double>((size_t)B * NP * R * 3), *d_pb = dev<double>((size_t)B * NP * R);
    int *d_pm = dev<int>((size_t)B * NP), *d_pidx = dev<int>((size_t)B * N), *d_np = dev<int>(B), *d_ovf = dev<int>(B), *d_far = dev<int>(B);
    double *d_hdr = dev<double>((size_t)B * N * 10), *d_rows = dev<double>((size_t)B * N * mcap * 4), *d_z = dev<double>((size_t)B * N * 17);
    int *d_nrows = dev<int>((size_t)B * N), *d_ii = dev<int>((size_t)B * 4);
    double* d_ir = dev<double>((size_t)B * 8);
    const double weights[5] = {7.0, 1.0, 80.0, 12.0, 0.5};              // launch/rotors_sim.launch:56-66
    nmpc_opts cold, warm;
    nmpc_default_opts(&cold); nmpc_default_opts(&warm); warm.mu0 = 0.1;
    std::vector<resilient_planner::SolveAcceptance> policy(B);
    std::vector<int> ii((size_t)B * 4), ovf(B), accept(B);
    int* d_accept = dev<int>((size_t)B);
    std::vector<double> zh((size_t)B * N * 17), toff(B);
    int bad = 0;
    for (int cycle = 0; cycle < 3; cycle++) {
        for (int b = 0; b < B; b++) toff[b] = cycle * Ts;
        cudaMemcpy(d_toff, toff.data(), B * sizeof(double), cudaMemcpyHostToDevice);
        // last_yaw_ / pos of stage 1 of the previous plan (setFORCESParams :486, getCurTraj :136)
        cudaMemcpy2D(d_lastyaw, sizeof(double), d_zprev + 17 + 16, (size_t)N * 17 * sizeof(double), sizeof(double), B, cudaMemcpyDeviceToDevice);
        cudaMemcpy2D(d_pos1, 3 * sizeof(double), d_zprev + 17 + 8, (size_t)N * 17 * sizeof(double), 3 * sizeof(double), B, cudaMemcpyDeviceToDevice);
        if (cycle > 0) CHECK(nmpc_shift_warm_start_f64(B, N, d_zprev, d_xinit, d_z0, 0, nullptr));
        CHECK(nmpc_propagate_ellipsoids_f64(B, N, d_zprev, nullptr, d_E, nullptr));
        CHECK(nmpc_sample_reference_f64(B, N, P, Ts, d_path, d_size, d_toff, d_lastyaw, d_pos1, d_ref, d_yaw, d_far, nullptr));
        CHECK(nmpc_select_corridors_f64(B, N, M, NP, R, d_cloud, 3LL * M, d_cn, d_ref, d_yaw, d_E, nullptr, d_pA, d_pb, d_pm, d_pidx, d_np, d_ovf, nullptr));
        CHECK(nmpc_pack_params_f64(B, N, NP, R, mcap, d_ref, d_yaw, d_ext, d_E, d_pA, d_pb, d_pm, d_pidx, weights, d_hdr, d_rows, d_nrows, nullptr));
        CHECK(nmpc_solve_batch_f64(B, N, mcap, d_xinit, d_z0, d_hdr, d_rows, d_nrows, 0, cycle ? &warm : &cold, d_z, d_ii, d_ir, nullptr));
        cudaMemcpy(ii.data(), d_ii, ii.size() * sizeof(int), cudaMemcpyDeviceToHost);
        cudaMemcpy(ovf.data(), d_ovf, B * sizeof(int), cudaMemcpyDeviceToHost);
        cudaMemcpy(zh.data(), d_z, zh.size() * sizeof(double), cudaMemcpyDeviceToHost);
        int ok = 0, itsum = 0, over = 0;
        double track = 0;
        for (int b = 0; b < B; b++) {
            accept[b] = policy[b].consume(ii[(size_t)b * 4]) ? 1 : 0;   // solveNMPC acceptance (:398-421)
            ok += accept[b];
            itsum += ii[(size_t)b * 4 + 1]; over += ovf[b] != 0;
            const double* z1 = &zh[((size_t)b * N + 1) * 17];
            const double* p1 = &path[((size_t)b * P + cycle + 1) * 3];
            track = std::fmax(track, std::hypot(z1[8] - p1[0], z1[9] - p1[1]));
        }
        std::printf("cycle %d: %d/%d accepted, mean it %.2f, corridor overflow %d, max |pos1 - path| %.3f m\n", cycle, ok, B,
                    (double)itsum / B, over, track);
        bad += (B - ok) + over;
        // updateNormal / updateFORCESResults for the accepted agents only (kept yaw-wrapped); a rejected agent keeps
        // nothing of the failed solve and restarts from the cold guess next cycle (:363-364)
        cudaMemcpy(d_accept, accept.data(), B * sizeof(int), cudaMemcpyHostToDevice);
        CHECK(nmpc_adopt_plans_f64(B, N, d_z, d_ii, d_accept, nullptr, d_zprev, nullptr, 1, nullptr));
    }
    return bad;
}
