// sharded_demo.cpp -- the multi-GPU host side in C++ over the C ABI only (include/nmpc_b200.h): one PROCESS per rank, contiguous
// shards, and the end-of-batch collation fused into the solve kernel (nmpc_peers_* / nmpc_solve_batch_sharded_p2p_f64: the
// kernel's epilogue stores every result into all ranks' buffers through CUDA-IPC mappings, barrier kernels around it).
//
//   sharded_demo [world = 2] [problems per rank = 64]
//
// The parent forks `world` ranks and plays the role an MPI_Allgather would: it collects the 64-byte IPC handles and hands the
// whole table to every rank (socketpairs).  Rank r uses GPU r % (number of GPUs) -- with one GPU all ranks share it, which
// still exercises the cross-process mappings, the peer stores and the barrier kernels.  Every rank then re-solves EVERY
// rank's shard locally (the kernels are deterministic) and checks that the collated buffers are bit-identical.
// Exit status 0 iff all ranks agree.  Build: host/Makefile.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <csignal>
#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cuda_runtime.h>

#include "../../include/nmpc_b200.h"

namespace {
constexpr int N = 20, MCAP = 4;

template <class T> T* dev(const std::vector<T>& h)
{
    void* p = nullptr;
    cudaMalloc(&p, (h.empty() ? 1 : h.size()) * sizeof(T));
    cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return static_cast<T*>(p);
}

// rank q's shard: vehicles hovering at scattered start points, asked to follow a straight line at 1 m/s with a yaw change,
// a constant disturbance, and a four-plane corridor (a 3 m tube around the line) from stage 1 on
struct Shard { std::vector<double> xinit, z0, hdr, rows; std::vector<int> nrows; };
Shard make_shard(int q, int B)
{
    Shard s;
    s.xinit.assign((size_t)B * 9, 0.0); s.z0.assign((size_t)B * N * 17, 0.0); s.hdr.assign((size_t)B * N * 10, 0.0);
    s.rows.assign((size_t)B * N * MCAP * 4, 0.0); s.nrows.assign((size_t)B * N, 0);
    unsigned seed = 977u * (unsigned)(q + 1);
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return (seed >> 8) * (1.0 / 16777216.0); };
    for (int b = 0; b < B; b++) {
        double* x0 = &s.xinit[(size_t)b * 9];
        x0[0] = -8 + 16 * rnd(); x0[1] = -8 + 16 * rnd(); x0[2] = 1 + rnd(); x0[8] = -0.5 + rnd();
        const double dir = 6.283185307179586 * rnd(), cx = std::cos(dir), cy = std::sin(dir), yaw = x0[8] + (-1 + 2 * rnd());
        const double fe[3] = {-1 + 2 * rnd(), -1 + 2 * rnd(), -0.5 + rnd()};
        for (int k = 0; k < N; k++) {
            double* z = &s.z0[((size_t)b * N + k) * 17];
            z[3] = z[7] = 7.3;                                   // initMPCOutput: hover thrust, states = xinit
            for (int i = 0; i < 9; i++) z[8 + i] = x0[i];
            double* h = &s.hdr[((size_t)b * N + k) * 10];
            h[0] = x0[0] + 0.05 * k * cx; h[1] = x0[1] + 0.05 * k * cy; h[2] = x0[2];
            h[3] = fe[0]; h[4] = fe[1]; h[5] = fe[2]; h[6] = 7.0; h[7] = 1.0; h[8] = 80.0; h[9] = yaw;
            if (k == 0) continue;
            double* r = &s.rows[((size_t)b * N + k) * MCAP * 4];
            const double nx = -cy, ny = cx;                      // +-normal of the line in the plane, +-z
            const double a[4][3] = {{nx, ny, 0}, {-nx, -ny, 0}, {0, 0, 1}, {0, 0, -1}};
            for (int j = 0; j < 4; j++) {
                r[4 * j] = a[j][0]; r[4 * j + 1] = a[j][1]; r[4 * j + 2] = a[j][2];
                r[4 * j + 3] = a[j][0] * x0[0] + a[j][1] * x0[1] + a[j][2] * x0[2] + 3.0;
            }
            s.nrows[(size_t)b * N + k] = 4;
        }
    }
    return s;
}

bool xfer(int fd, void* buf, size_t n, bool wr)
{
    char* p = static_cast<char*>(buf);
    while (n) {
        const ssize_t k = wr ? write(fd, p, n) : read(fd, p, n);
        if (k <= 0) return false;
        p += k; n -= (size_t)k;
    }
    return true;
}

#define CHECK(call) do { int rc_ = (call); if (rc_ != 0) { std::printf("rank %d: %s -> %d: %s\n", rank, #call, rc_, nmpc_last_error()); return 1; } } while (0)

int run_rank(int rank, int world, int B, int fd)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { std::printf("rank %d: no CUDA device\n", rank); return 1; }
    cudaSetDevice(rank % ndev);
    const size_t zw = (size_t)B * N * 17;
    nmpc_peers* peers = nullptr;
    CHECK(nmpc_peers_create(world, rank, (size_t)world * zw * sizeof(double), (size_t)world * B * 4, &peers));
    std::vector<unsigned char> table((size_t)world * 64);
    unsigned char mine[64];
    CHECK(nmpc_peers_export(peers, mine));
    if (!xfer(fd, mine, 64, true) || !xfer(fd, table.data(), table.size(), false)) { std::printf("rank %d: handle exchange failed\n", rank); return 1; }
    CHECK(nmpc_peers_connect(peers, table.data()));

    Shard s = make_shard(rank, B);
    double *d_x = dev(s.xinit), *d_z0 = dev(s.z0), *d_h = dev(s.hdr), *d_r = dev(s.rows);
    int* d_n = dev(s.nrows);
    double* d_ir = dev(std::vector<double>((size_t)B * 8));
    nmpc_opts o;
    nmpc_default_opts(&o);
    for (int rep = 0; rep < 3; rep++)                             // three batches in a row: the barrier epochs advance
        CHECK(nmpc_solve_batch_sharded_p2p_f64(peers, B, N, MCAP, d_x, d_z0, d_h, d_r, d_n, 0, &o, d_ir, /*mode: fp64 kernel*/ 0, nullptr));
    cudaDeviceSynchronize();
    CHECK(nmpc_peers_status(peers));
    std::vector<double> z_all((size_t)world * zw), z_ref(zw);
    std::vector<int> info_all((size_t)world * B * 4);
    cudaMemcpy(z_all.data(), nmpc_peers_z(peers), z_all.size() * sizeof(double), cudaMemcpyDeviceToHost);
    cudaMemcpy(info_all.data(), nmpc_peers_info(peers), info_all.size() * sizeof(int), cudaMemcpyDeviceToHost);

    // the checker: every rank's shard solved here, on this rank's GPU, by the plain entry point
    int bad = 0, solved = 0;
    double* d_zr = dev(std::vector<double>(zw));
    int* d_ii = dev(std::vector<int>((size_t)B * 4));
    for (int q = 0; q < world; q++) {
        Shard t = make_shard(q, B);
        double *qx = dev(t.xinit), *qz = dev(t.z0), *qh = dev(t.hdr), *qr = dev(t.rows);
        int* qn = dev(t.nrows);
        CHECK(nmpc_solve_batch_f64(B, N, MCAP, qx, qz, qh, qr, qn, 0, &o, d_zr, d_ii, d_ir, nullptr));
        cudaMemcpy(z_ref.data(), d_zr, zw * sizeof(double), cudaMemcpyDeviceToHost);
        bad += std::memcmp(z_ref.data(), &z_all[(size_t)q * zw], zw * sizeof(double)) != 0;
        for (int b = 0; b < B; b++) solved += info_all[((size_t)q * B + b) * 4] == 1;
        cudaFree(qx); cudaFree(qz); cudaFree(qh); cudaFree(qr); cudaFree(qn);
    }
    std::printf("rank %d of %d on GPU %d: %d x %d problems collated, %d / %d with exit flag 1, %d of %d slices differ from a local solve\n",
                rank, world, rank % ndev, world, B, solved, world * B, bad, world);
    // nobody may still be writing into this rank's buffers when they are freed: meet at the parent first
    char tok = bad ? 'x' : 'k';
    if (!xfer(fd, &tok, 1, true) || !xfer(fd, &tok, 1, false)) return 1;
    nmpc_peers_destroy(peers);
    return (bad || solved != world * B) ? 1 : 0;
}
}  // namespace

int main(int argc, char** argv)
{
    const int world = argc > 1 ? std::atoi(argv[1]) : 2, B = argc > 2 ? std::atoi(argv[2]) : 64;
    if (world < 1 || world > 16 || B < 2 || (B & 1)) { std::printf("usage: sharded_demo [world 1..16] [even problems per rank]\n"); return 2; }
    std::signal(SIGPIPE, SIG_IGN);                                // a rank that died shows up as a failed write, not as a signal
    std::vector<int> fds(world);
    std::vector<pid_t> pids(world);
    for (int r = 0; r < world; r++) {                             // fork BEFORE any CUDA call: a CUDA context does not survive fork()
        int sv[2];
        if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) return 3;
        pids[r] = fork();
        if (pids[r] == 0) {
            close(sv[0]);
            for (int q = 0; q < r; q++) close(fds[q]);
            const int rc = run_rank(r, world, B, sv[1]);
            std::fflush(stdout);
            std::_Exit(rc);
        }
        close(sv[1]);
        fds[r] = sv[0];
    }
    std::vector<unsigned char> table((size_t)world * 64);
    bool ok = true;
    for (int r = 0; r < world; r++) ok &= xfer(fds[r], &table[(size_t)r * 64], 64, false);          // "all-gather" of the handles
    for (int r = 0; r < world; r++) ok &= xfer(fds[r], table.data(), table.size(), true);
    std::vector<char> tok(world, 0);
    for (int r = 0; r < world; r++) ok &= xfer(fds[r], &tok[r], 1, false);                          // barrier before the buffers go away
    for (int r = 0; r < world; r++) ok &= xfer(fds[r], &tok[r], 1, true);
    int fail = ok ? 0 : 1;
    for (int r = 0; r < world; r++) {
        int st = 0;
        waitpid(pids[r], &st, 0);
        fail += !(WIFEXITED(st) && WEXITSTATUS(st) == 0);
    }
    std::printf("%s\n", fail ? "sharded_demo: FAILED" : "sharded_demo: all ranks hold identical, locally reproducible results");
    return fail;
}
