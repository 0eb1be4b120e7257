// nmpc_capi.cu -- C ABI of libnmpc_b200.so (see include/nmpc_b200.h).
//
// Host side of the drop-in: the batched entry points, and the two reference-named symbols
//   FORCESNLPsolver_normal_solve / FORCESNLPsolver_final_solve
// that plan_manage/src/forces_normal.cpp:139 and forces_final.cpp:138 call.  No CPU fallback
// exists anywhere in this file: every path ends in the sm_100a kernel or in an error code.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/FORCESNLPsolver_final.h"
#include "../../include/FORCESNLPsolver_normal.h"
#include "../../include/nmpc_b200.h"
#include "nmpc_backsolve.cuh"
#include "nmpc_corridor.cuh"
#include "nmpc_ellipsoid.cuh"
#include "nmpc_ipm.cuh"
#include "nmpc_ipm_mixed.cuh"
#include "nmpc_ipm_group.cuh"
#include "nmpc_prep.cuh"

// ---- the ABI contract of the reference headers --------------------------------
static_assert(sizeof(FORCESNLPsolver_normal_params) == 23600, "params size");
static_assert(offsetof(FORCESNLPsolver_normal_params, x0) == 72, "x0 offset");
static_assert(offsetof(FORCESNLPsolver_normal_params, all_parameters) == 2792, "all_parameters offset");
static_assert(offsetof(FORCESNLPsolver_normal_params, num_of_threads) == 23592, "num_of_threads offset");
static_assert(sizeof(FORCESNLPsolver_normal_output) == 2720, "output size");
static_assert(sizeof(FORCESNLPsolver_normal_info) == 136, "info size");
static_assert(offsetof(FORCESNLPsolver_normal_info, res_eq) == 8, "res_eq offset");
static_assert(offsetof(FORCESNLPsolver_normal_info, lsit_aff) == 96, "lsit_aff offset");
static_assert(offsetof(FORCESNLPsolver_normal_info, step_aff) == 104, "step_aff offset");
static_assert(offsetof(FORCESNLPsolver_normal_info, solvetime) == 120, "solvetime offset");
static_assert(sizeof(FORCESNLPsolver_final_params) == 23600 && sizeof(FORCESNLPsolver_final_output) == 2720 &&
                  sizeof(FORCESNLPsolver_final_info) == 136, "final ABI");
static_assert(sizeof(nmpc_opts) == sizeof(nmpc::Opts), "opts mirror");
static_assert(offsetof(nmpc_opts, mixed) == offsetof(nmpc::Opts, mixed), "opts mirror");

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return fail(NMPC_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));     \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: remember the largest
// value configured per (kernel, device), so that a thread which solves on cuda:0 and then on cuda:1 configures both.
constexpr int kMaxDevices = 64;
template <typename K>
int ensure_smem(K kernel, size_t smem, size_t (&configured)[kMaxDevices])
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return fail(NMPC_ERR_CUDA, "device ordinal %d out of range", dev);
    if (smem > configured[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = smem;
    }
    return 0;
}

template <typename T, int N, bool PC>
int launch_pc(const nmpc::Params<T>& prm, cudaStream_t st)
{
    using L = nmpc::Layout<T, N, PC>;
    const size_t smem = L::bytes(prm.mcap);
    static thread_local size_t configured[kMaxDevices] = {};
    if (int rc = ensure_smem(nmpc::nmpc_ipm_kernel<T, N, PC>, smem, configured)) return rc;
    nmpc::nmpc_ipm_kernel<T, N, PC><<<prm.B, 32, smem, st>>>(prm);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
template <typename T, int N>
int launch(const nmpc::Params<T>& prm, cudaStream_t st)
{
    if (prm.o.pc) return launch_pc<T, N, true>(prm, st);
    return launch_pc<T, N, false>(prm, st);
}
template <int N>
int launch_mixed(const nmpc::MixedParams& prm, cudaStream_t st)
{
    const size_t smem = nmpc::MLayout<N>::bytes(prm.mcap);
    static thread_local size_t configured[kMaxDevices] = {};
    if (int rc = ensure_smem(nmpc::nmpc_ipm_mixed_kernel<N>, smem, configured)) return rc;
    nmpc::nmpc_ipm_mixed_kernel<N><<<prm.B, 32, smem, st>>>(prm);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

template <int N>
int launch_group(const nmpc::MixedParams& prm, cudaStream_t st)
{
    const size_t smem = nmpc::GLayout<N>::bytes(prm.mcap);
    static thread_local size_t configured[kMaxDevices] = {};
    // more problems than SMs: the variant compiled for two resident CTAs per SM (128 registers per thread) keeps them all
    // in one wave up to twice the SM count
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    static thread_local int sm_count[kMaxDevices] = {};
    if (dev >= 0 && dev < kMaxDevices) {
        if (!sm_count[dev]) CUDA_TRY(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
        sms = sm_count[dev];
    }
    if (N == 20 && prm.B > sms && 2 * smem + 2048 <= 233472) {
        static thread_local size_t configured2[kMaxDevices] = {};
        if (int rc = ensure_smem(nmpc::nmpc_ipm_group_kernel<N, 2>, smem, configured2)) return rc;
        nmpc::nmpc_ipm_group_kernel<N, 2><<<prm.B, nmpc::GROUP_THREADS, smem, st>>>(prm);
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    if (int rc = ensure_smem(nmpc::nmpc_ipm_group_kernel<N>, smem, configured)) return rc;
    nmpc::nmpc_ipm_group_kernel<N><<<prm.B, nmpc::GROUP_THREADS, smem, st>>>(prm);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Stream-ordered scratch memory comes from the library's OWN pool (one per device) that keeps what it has been given: the
// default pool returns freed memory to the driver at the next synchronisation, which turns every small solve into a
// cudaMalloc / cudaFree pair (milliseconds).  The process-wide default pool's settings are not touched.
std::mutex g_pool_mutex;
cudaMemPool_t g_pool[kMaxDevices] = {};
int scratch_pool(cudaMemPool_t* out)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return fail(NMPC_ERR_CUDA, "device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!g_pool[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        CUDA_TRY(cudaMemPoolCreate(&g_pool[dev], &props));
        uint64_t keep = UINT64_MAX;
        CUDA_TRY(cudaMemPoolSetAttribute(g_pool[dev], cudaMemPoolAttrReleaseThreshold, &keep));
    }
    *out = g_pool[dev];
    return 0;
}

// order[0 .. *count) <- the problems whose exit flag is not 1 (optimal); one pass, order of arrival
__global__ void collect_unsolved_kernel(int B, const int* info_int, int* count, int* order)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B && info_int[(size_t)b * 4] != 1) order[atomicAdd(count, 1)] = b;
}

// argument checks shared by the fp64 and the mixed-precision device entry points
int check_solve_args(int B, int N, int mcap, const void* xinit, const void* z0, const void* hdr, const void* rows,
                     const int* nrows, int variant, const nmpc_opts* opts, const void* z_out, const int* info_int,
                     const void* info_real, nmpc_opts* o)
{
    if (B < 0 || mcap < 0 || mcap > 32 || (variant != 0 && variant != 1))
        return fail(NMPC_ERR_ARG, "bad argument: B=%d mcap=%d variant=%d", B, mcap, variant);
    if (!nmpc_supported_horizon(N)) return fail(NMPC_ERR_ARG, "unsupported horizon N=%d (20 or 40)", N);
    if (B == 0) return 0;
    if (!xinit || !z0 || !hdr || !nrows || !z_out || !info_int || !info_real || (mcap > 0 && !rows))
        return fail(NMPC_ERR_ARG, "null pointer argument");
    // the per-problem blocks are moved by TMA bulk copies / 16-byte vector loads
    for (const void* p : {z0, hdr, rows, (const void*)nrows, z_out})
        if (reinterpret_cast<uintptr_t>(p) & 15) return fail(NMPC_ERR_ARG, "device pointers must be 16-byte aligned");
    if (opts) *o = *opts; else nmpc_default_opts(o);
    if (!(o->mu0 > 0) || !(o->mu_floor > 0) || o->maxit < 0 || o->max_bt < 0)
        return fail(NMPC_ERR_ARG, "bad solver options");
    return 0;
}

// fp64 kernel.  io32: the arrays are float (re-solve of mixed-precision failures); count: device counter of live `order` entries
int solve_device(int B, int N, int mcap, const void* xinit, const void* z0, const void* hdr, const void* rows,
                 const int* nrows, int variant, const nmpc_opts* opts, void* z_out, int* info_int,
                 void* info_real, void* stream, void* y_out = nullptr, void* zl_out = nullptr, void* zu_out = nullptr,
                 void* lc_out = nullptr, const int* order = nullptr, const int* count = nullptr, int io32 = 0,
                 const void* z_warm = nullptr, const nmpc::PeerOut* peers = nullptr)
{
    nmpc_opts o;
    if (int rc = check_solve_args(B, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_out, info_int, info_real, &o)) return rc;
    if (B == 0) return 0;
    nmpc::Params<double> prm;
    prm.B = B; prm.mcap = mcap; prm.variant = variant; prm.io32 = io32;
    prm.xinit = xinit; prm.z0 = z0; prm.hdr = hdr; prm.rows = rows; prm.nrows = nrows; prm.order = order; prm.count = count;
    prm.z_warm = z_warm;
    prm.z_out = z_out; prm.info_int = info_int; prm.info_real = info_real;
    prm.y_out = y_out; prm.zl_out = zl_out; prm.zu_out = zu_out; prm.lc_out = lc_out;
    if (peers) prm.peers = *peers;
    std::memcpy(&prm.o, &o, sizeof(o));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return N == 20 ? launch<double, 20>(prm, st) : launch<double, 40>(prm, st);
}

// Mixed-precision solve (nmpc_ipm_mixed.cuh) followed, on the same stream and without a host round trip, by the
// fp64 re-solve of whatever it did not bring to exit flag 1: single precision loses the cost-to-go's positive
// definiteness on rare, badly scaled instances (flag -5) or stalls (flag 0 after MIXED_BAIL_IT iterations).
// The list of those problems is built on the device; the fp64 grid is launched at full width and the CTAs beyond
// the list exit at once.  Re-solved problems carry info_int[b][3] = 1.
int solve_mixed(int B, int N, int mcap, const void* xinit, const void* z0, const void* hdr, const void* rows,
                const int* nrows, int variant, const nmpc_opts* opts, void* z_out, int* info_int, void* info_real,
                void* stream, int io32, void* y_out = nullptr, void* zl_out = nullptr, void* zu_out = nullptr,
                void* lc_out = nullptr, const int* order = nullptr, bool group = false, const nmpc::PeerOut* peers = nullptr)
{
    nmpc_opts o;
    if (int rc = check_solve_args(B, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_out, info_int, info_real, &o)) return rc;
    if (B == 0) return 0;
    nmpc::MixedParams prm;
    prm.B = B; prm.mcap = mcap; prm.variant = variant; prm.io32 = io32;
    prm.xinit = xinit; prm.z0 = z0; prm.hdr = hdr; prm.rows = rows; prm.nrows = nrows; prm.order = order;
    prm.z_out = z_out; prm.info_int = info_int; prm.info_real = info_real;
    prm.y_out = y_out; prm.zl_out = zl_out; prm.zu_out = zu_out; prm.lc_out = lc_out;
    if (peers) prm.peers = *peers;
    std::memcpy(&prm.o, &o, sizeof(o));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // group: the low-latency variant, one warp-group (256 threads) per problem (nmpc_ipm_group.cuh)
    if (int rc = group ? (N == 20 ? launch_group<20>(prm, st) : launch_group<40>(prm, st))
                       : (N == 20 ? launch_mixed<20>(prm, st) : launch_mixed<40>(prm, st))) return rc;
    if (o.mixed < 0) return 0;                      // opts.mixed = -1: no fp64 safety net (tests, profiling)
    int* ws = nullptr;
    cudaMemPool_t pool;
    if (int rc = scratch_pool(&pool)) return rc;
    CUDA_TRY(cudaMallocFromPoolAsync(reinterpret_cast<void**>(&ws), ((size_t)B + 4) * sizeof(int), pool, st));
    CUDA_TRY(cudaMemsetAsync(ws, 0, 4 * sizeof(int), st));
    collect_unsolved_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, info_int, ws, ws + 4);
    CUDA_TRY(cudaGetLastError());
    nmpc_opts o64 = o;
    o64.pc = 0;
    int rc = solve_device(B, N, mcap, xinit, z0, hdr, rows, nrows, variant, &o64, z_out, info_int, info_real, stream,
                          y_out, zl_out, zu_out, lc_out, ws + 4, ws, io32, /*z_warm = the mixed kernel's last iterates*/ z_out, peers);
    CUDA_TRY(cudaFreeAsync(ws, st));
    return rc;
}

template <typename T, int N>
int launch_factor(const nmpc::FactorParams<T>& q, cudaStream_t st)
{
    const size_t smem = nmpc::Layout<T, N>::bytes(0);
    static thread_local size_t configured[kMaxDevices] = {};
    if (int rc = ensure_smem(nmpc::riccati_factor_kernel<T, N>, smem, configured)) return rc;
    nmpc::riccati_factor_kernel<T, N><<<q.B, 32, smem, st>>>(q);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
template <typename T, int N>
int launch_backsolve(const nmpc::BacksolveParams<T>& q, cudaStream_t st)
{
    const size_t smem = nmpc::BsLayout<T, N>::bytes();
    static thread_local size_t configured[kMaxDevices] = {};
    if (int rc = ensure_smem(nmpc::kkt_backsolve_kernel<T, N>, smem, configured)) return rc;
    nmpc::kkt_backsolve_kernel<T, N><<<q.B, 32, smem, st>>>(q);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
template <typename T>
int factor_device(int B, int N, const T* phi, const T* jc, T* fac, int* status, void* stream)
{
    if (B < 0 || !nmpc_supported_horizon(N)) return fail(NMPC_ERR_ARG, "bad argument: B=%d N=%d", B, N);
    if (B == 0) return 0;
    if (!phi || !jc || !fac || !status) return fail(NMPC_ERR_ARG, "null pointer argument");
    nmpc::FactorParams<T> q{B, phi, jc, fac, status};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return N == 20 ? launch_factor<T, 20>(q, st) : launch_factor<T, 40>(q, st);
}
template <typename T>
int backsolve_device(int B, int N, const T* fac, const T* g, const T* d, T* dz, T* y, void* stream)
{
    if (B < 0 || !nmpc_supported_horizon(N)) return fail(NMPC_ERR_ARG, "bad argument: B=%d N=%d", B, N);
    if (B == 0) return 0;
    if (!fac || !g || !d || !dz || !y) return fail(NMPC_ERR_ARG, "null pointer argument");
    for (const void* p : {(const void*)fac, (const void*)g, (const void*)d, (const void*)dz, (const void*)y})
        if (reinterpret_cast<uintptr_t>(p) & 15) return fail(NMPC_ERR_ARG, "device pointers must be 16-byte aligned");
    nmpc::BacksolveParams<T> q{B, fac, g, d, dz, y};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return N == 20 ? launch_backsolve<T, 20>(q, st) : launch_backsolve<T, 40>(q, st);
}

// grow-only device arena for the host-pointer API and the FORCES shim; one arena and one set of streams per device
struct Arena {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t n)
    {
        if (n <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        CUDA_TRY(cudaMalloc(&p, n));
        cap = n;
        return 0;
    }
};
constexpr int kHostStreams = 8;
struct HostCtx {
    Arena arena;
    cudaStream_t streams[kHostStreams] = {};
};
std::mutex g_host_mutex;   // the reference solver is non-reentrant (static workspace); mirror that
HostCtx g_host[kMaxDevices];

inline size_t align256(size_t n) { return (n + 255) & ~size_t(255); }

// Host buffers in, host buffers out.  esz = 8: double arrays, fp64 kernel (mixed = false) or mixed-precision kernel;
// esz = 4: float arrays, mixed-precision kernel.
int solve_host(int B, int N, int mcap, const void* xinit, const void* z0, const void* hdr, const void* rows,
               const int* nrows, int variant, const nmpc_opts* opts, void* z_out, int* info_int, void* info_real,
               size_t esz, bool mixed, bool group = false)
{
    if (B < 0) return fail(NMPC_ERR_ARG, "B < 0");
    if (B == 0) return 0;
    if (!nmpc_supported_horizon(N) || mcap < 0 || mcap > 32) return fail(NMPC_ERR_ARG, "bad N=%d or mcap=%d", N, mcap);
    std::lock_guard<std::mutex> lock(g_host_mutex);
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return fail(NMPC_ERR_CUDA, "device ordinal %d out of range", dev);
    HostCtx& ctx = g_host[dev];
    const int n_chunks = B >= 4096 ? 8 : (B >= 2048 ? 4 : 1);     // one stream per chunk (kHostStreams >= 8)
    for (int i = 0; i < n_chunks; i++)
        if (!ctx.streams[i]) CUDA_TRY(cudaStreamCreateWithFlags(&ctx.streams[i], cudaStreamNonBlocking));
    const size_t n_x = (size_t)B * 9 * esz, n_z = (size_t)B * N * 17 * esz;
    const size_t n_h = (size_t)B * N * 10 * esz, n_r = (size_t)B * N * mcap * 4 * esz;
    const size_t n_n = (size_t)B * N * sizeof(int), n_ii = (size_t)B * 4 * sizeof(int), n_ir = (size_t)B * 8 * esz;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += align256(n ? n : 1); return o; };
    const size_t o_x = take(n_x), o_z = take(n_z), o_h = take(n_h), o_r = take(n_r), o_n = take(n_n);
    const size_t o_zo = take(n_z), o_ii = take(n_ii), o_ir = take(n_ir);
    if (int rc = ctx.arena.reserve(off)) return rc;
    char* base = static_cast<char*>(ctx.arena.p);
    const char *hx = static_cast<const char*>(xinit), *hz = static_cast<const char*>(z0), *hh = static_cast<const char*>(hdr),
               *hr = static_cast<const char*>(rows);
    char *hzo = static_cast<char*>(z_out), *hir = static_cast<char*>(info_real);
    // Large batches are cut into chunks that round-robin over a few streams: the H2D copy of chunk
    // i+1 and the D2H copy of chunk i-1 overlap the solve of chunk i, and the next chunk's CTAs fill
    // the tail wave of the previous kernel.  Chunk boundaries are multiples of 4 problems, so every
    // per-problem block keeps the 16-byte alignment the TMA copies need (fp32 and fp64).
    const int per = ((B + n_chunks - 1) / n_chunks + 3) & ~3;
    // Pinned (device-accessible) result buffers are written by the kernel itself: the epilogue of every solve stores its
    // solution to the device copy AND straight into the caller's buffer (one more destination of store_solution, as for
    // the multi-GPU peers), so the 2.7 KB per problem cross PCIe while the rest of the batch is still being solved.
    // Otherwise the kernels of all chunks end together and their device-to-host copies queue up behind the last one.
    char* hz_direct = nullptr;
    int* hii_direct = nullptr;
    {
        const char* off = std::getenv("NMPC_B200_DIRECT_HOST");
        cudaPointerAttributes pa;
        if (!(off && off[0] == '0')) {
            if (cudaPointerGetAttributes(&pa, z_out) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer &&
                (reinterpret_cast<uintptr_t>(pa.devicePointer) & 15) == 0)
                hz_direct = static_cast<char*>(pa.devicePointer);
            if (hz_direct && cudaPointerGetAttributes(&pa, info_int) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer &&
                (reinterpret_cast<uintptr_t>(pa.devicePointer) & 15) == 0)
                hii_direct = static_cast<int*>(pa.devicePointer);
        }
        cudaGetLastError();                                   // a pageable pointer makes the query fail on old drivers: not an error here
    }
    auto enqueue = [&]() -> int {
        for (int c = 0, lo = 0; lo < B; c++, lo += per) {
            const int nb = (B - lo < per) ? B - lo : per;
            cudaStream_t st = ctx.streams[c % kHostStreams];
            const size_t px = 9 * esz, pz = (size_t)N * 17 * esz, ph = (size_t)N * 10 * esz, pr = (size_t)N * mcap * 4 * esz;
            const size_t pn = (size_t)N * sizeof(int), pii = 4 * sizeof(int), pir = 8 * esz;
            CUDA_TRY(cudaMemcpyAsync(base + o_x + lo * px, hx + lo * px, nb * px, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(base + o_z + lo * pz, hz + lo * pz, nb * pz, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(base + o_h + lo * ph, hh + lo * ph, nb * ph, cudaMemcpyHostToDevice, st));
            if (pr) CUDA_TRY(cudaMemcpyAsync(base + o_r + lo * pr, hr + lo * pr, nb * pr, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(base + o_n + lo * pn, nrows + (size_t)lo * N, nb * pn, cudaMemcpyHostToDevice, st));
            int* d_ii = reinterpret_cast<int*>(base + o_ii + lo * pii);
            nmpc::PeerOut po;
            if (hz_direct) {
                po.n = 1;
                po.z[0] = hz_direct + lo * pz;
                po.info[0] = hii_direct ? hii_direct + (size_t)lo * 4 : nullptr;
            }
            int rc;
            if (mixed)
                rc = solve_mixed(nb, N, mcap, base + o_x + lo * px, base + o_z + lo * pz, base + o_h + lo * ph, base + o_r + lo * pr,
                                 reinterpret_cast<const int*>(base + o_n + lo * pn), variant, opts, base + o_zo + lo * pz, d_ii,
                                 base + o_ir + lo * pir, st, esz == 4, nullptr, nullptr, nullptr, nullptr, nullptr, group, &po);
            else
                rc = solve_device(nb, N, mcap, base + o_x + lo * px, base + o_z + lo * pz, base + o_h + lo * ph, base + o_r + lo * pr,
                                  reinterpret_cast<const int*>(base + o_n + lo * pn), variant, opts, base + o_zo + lo * pz, d_ii,
                                  base + o_ir + lo * pir, st, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, &po);
            if (rc) return rc;
            if (!hz_direct) CUDA_TRY(cudaMemcpyAsync(hzo + lo * pz, base + o_zo + lo * pz, nb * pz, cudaMemcpyDeviceToHost, st));
            if (!hii_direct) CUDA_TRY(cudaMemcpyAsync(info_int + (size_t)lo * 4, d_ii, nb * pii, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(hir + lo * pir, base + o_ir + lo * pir, nb * pir, cudaMemcpyDeviceToHost, st));
        }
        return 0;
    };
    const int rc = enqueue();
    // also on failure: nothing may still be in flight into the shared arena (or the caller's buffers) when we return
    cudaError_t serr = cudaSuccess;
    for (int i = 0; i < n_chunks; i++) {
        const cudaError_t e = cudaStreamSynchronize(ctx.streams[i]);
        if (e != cudaSuccess && serr == cudaSuccess) serr = e;
    }
    if (rc) return rc;
    if (serr != cudaSuccess) return fail(NMPC_ERR_CUDA, "cudaStreamSynchronize failed: %s", cudaGetErrorString(serr));
    return 0;
}

// ---- end-of-batch collation over NCCL (north star: "NCCL all-gather only for the end-of-batch result
// collation") -------------------------------------------------------------------------------------------------------
// libnccl is resolved at run time (dlopen "libnccl.so.2": the copy already in the process -- e.g. the one torch brought
// -- or the system's), so single-GPU users carry no NCCL dependency and the library never mixes two NCCL builds' handles.
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*MemAlloc)(void**, size_t) = nullptr;            // optional (NCCL >= 2.19)
    ncclResult_t (*MemFree)(void*) = nullptr;
    ncclResult_t (*CommRegister)(const ncclComm_t, void*, size_t, void**) = nullptr;
    ncclResult_t (*CommDeregister)(const ncclComm_t, void*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};
std::mutex g_nccl_mutex;
NcclApi g_nccl;

int nccl_api(NcclApi** out)
{
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (!g_nccl.h) {
        const char* env = getenv("NMPC_B200_NCCL");
        void* h = dlopen(env && env[0] ? env : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return fail(NMPC_ERR_CUDA, "cannot load libnccl.so.2: %s (set NMPC_B200_NCCL)", dlerror());
        NcclApi a;
        a.h = h;
#define NMPC_NCCL_SYM(field, name, required)                                                       \
        a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name));                             \
        if (required && !a.field) return fail(NMPC_ERR_CUDA, "libnccl lacks %s", name);
        NMPC_NCCL_SYM(GetUniqueId, "ncclGetUniqueId", true)
        NMPC_NCCL_SYM(CommInitRank, "ncclCommInitRank", true)
        NMPC_NCCL_SYM(CommDestroy, "ncclCommDestroy", true)
        NMPC_NCCL_SYM(AllGather, "ncclAllGather", true)
        NMPC_NCCL_SYM(GroupStart, "ncclGroupStart", true)
        NMPC_NCCL_SYM(GroupEnd, "ncclGroupEnd", true)
        NMPC_NCCL_SYM(GetErrorString, "ncclGetErrorString", true)
        NMPC_NCCL_SYM(GetVersion, "ncclGetVersion", true)
        NMPC_NCCL_SYM(MemAlloc, "ncclMemAlloc", false)
        NMPC_NCCL_SYM(MemFree, "ncclMemFree", false)
        NMPC_NCCL_SYM(CommRegister, "ncclCommRegister", false)
        NMPC_NCCL_SYM(CommDeregister, "ncclCommDeregister", false)
#undef NMPC_NCCL_SYM
        g_nccl = a;
    }
    *out = &g_nccl;
    return 0;
}

#define NCCL_TRY(api, expr)                                                                        \
    do {                                                                                           \
        ncclResult_t _r = (expr);                                                                  \
        if (_r != ncclSuccess) return fail(NMPC_ERR_CUDA, "%s failed: %s", #expr, (api)->GetErrorString(_r)); \
    } while (0)

}  // namespace

struct nmpc_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    bool owned = true;
    struct Reg { void* ptr; void* handle; bool nccl_mem; };
    std::vector<Reg> regs;
};

namespace {

// ---- reference ABI shim ------------------------------------------------------------------------
// Unpacks the 130-slot per-stage parameter layout (matlab_code/setup.m:60-66) into the native one.
// All-zero padding rows (forces_normal.cpp:127-135: A = 0, b = 0, i.e. the constant 0 <= 1e-5)
// carry no information and are dropped.
int forces_solve(const double* xinit, const double* x0, const double* allp, double* out340,
                 int variant, int* it, int* nbt, double ir[8], double* seconds, int* n_ineq)
{
    constexpr int N = 20;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<double> hdr(N * 10);
    std::vector<int> nrows(N, 0);
    int mcap = 1;
    for (int k = 0; k < N; k++) {
        const double* p = allp + k * 130;
        int m = 0;
        for (int j = 0; j < 30; j++)
            if (p[10 + 3 * j] != 0.0 || p[11 + 3 * j] != 0.0 || p[12 + 3 * j] != 0.0 || p[100 + j] != 0.0) m++;
        nrows[k] = m;
        if (m > mcap) mcap = m;
    }
    std::vector<double> rows((size_t)N * mcap * 4, 0.0);
    for (int k = 0; k < N; k++) {
        const double* p = allp + k * 130;
        std::memcpy(&hdr[k * 10], p, 10 * sizeof(double));
        int m = 0;
        for (int j = 0; j < 30; j++) {
            if (p[10 + 3 * j] == 0.0 && p[11 + 3 * j] == 0.0 && p[12 + 3 * j] == 0.0 && p[100 + j] == 0.0) continue;
            double* r = &rows[((size_t)k * mcap + m) * 4];
            r[0] = p[10 + 3 * j]; r[1] = p[11 + 3 * j]; r[2] = p[12 + 3 * j]; r[3] = p[100 + j];
            m++;
        }
    }
    int ii[4] = {0, 0, 0, 0};
    // One vehicle, one solve: what counts is the latency of that solve, so the default is the warp-group kernel (256
    // threads on the one problem, mixed precision at the reference tolerances, fp64 safety net on the same stream).
    // NMPC_B200_SHIM=fp64 selects the one-warp fp64 kernel, =mixed the one-warp mixed-precision kernel.
    const char* sel = std::getenv("NMPC_B200_SHIM");
    const bool fp64 = sel && std::strcmp(sel, "fp64") == 0, warp = sel && std::strcmp(sel, "mixed") == 0;
    int rc = solve_host(1, N, mcap, xinit, x0, hdr.data(), rows.data(), nrows.data(), variant, nullptr,
                        out340, ii, ir, sizeof(double), !fp64, !fp64 && !warp);
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    // inequalities of the problem as solved: two bound sides per free variable + the live corridor rows of stages 1..N-1
    *n_ineq = 2 * (8 + (N - 1) * 17);
    for (int k = 1; k < N; k++) *n_ineq += nrows[k];
    // library failures in the reference's own vocabulary (header :110-139): a bad argument is PARAM_VALUE_ERROR (-11);
    // "no CUDA device / CUDA runtime failure" has no counterpart but LICENSE_ERROR (-100, "solver not valid on this
    // machine"), which the planner treats like any other failure (nmpc_solver.cpp:398-421).  nmpc_last_error() has the text.
    if (rc) return rc == NMPC_ERR_ARG ? PARAM_VALUE_ERROR_FORCESNLPsolver_normal : LICENSE_ERROR_FORCESNLPsolver_normal;
    *it = ii[1]; *nbt = ii[2];
    return ii[0];
}

template <typename Info>
void fill_info(Info* info, int it, int nbt, const double ir[8], double seconds, int n_ineq)
{
    if (!info) return;
    nmpc_opts o;
    nmpc_default_opts(&o);
    info->it = it; info->it2opt = it;
    info->res_eq = ir[0]; info->res_ineq = ir[1]; info->rsnorm = ir[2]; info->rcompnorm = ir[3];
    info->pobj = ir[4];
    info->dgap = ir[5] * n_ineq;
    info->dobj = ir[4] - info->dgap;
    info->rdgap = ir[4] != 0.0 ? std::fabs(info->dgap / ir[4]) : 0.0;
    info->mu = ir[5]; info->mu_aff = ir[5]; info->sigma = o.sigma;
    info->lsit_aff = 0; info->lsit_cc = nbt;
    info->step_aff = ir[7]; info->step_cc = ir[6];
    info->solvetime = seconds; info->fevalstime = 0.0;
}

template <typename Params, typename Output, typename Info>
int forces_entry(Params* params, Output* output, Info* info, FILE* fs, int variant)
{
    if (!params || !output) return PARAM_VALUE_ERROR_FORCESNLPsolver_normal;
    int it = 0, nbt = 0, n_ineq = 0; double ir[8] = {0}, sec = 0;
    const int flag = forces_solve(params->xinit, params->x0, params->all_parameters, output->x01, variant, &it, &nbt, ir, &sec, &n_ineq);
    fill_info(info, it, nbt, ir, sec, n_ineq);
    if (fs) fprintf(fs, "FORCESNLPsolver_%s (nmpc_b200): exitflag %d, it %d, pobj %.6e, res_eq %.2e, rsnorm %.2e, time %.3e s%s%s\n",
                    variant ? "final" : "normal", flag, it, ir[4], ir[0], ir[2], sec, flag <= -100 ? " -- " : "", flag <= -100 ? g_err : "");
    return flag;
}

// ---- device model probe (tests): evaluates the device model exactly as the solver does -----------
__global__ void model_eval_kernel(int n, const double* z, const double* p, const int* stage, int n_stages,
                                  int variant, double* f, double* grad, double* c, double* jc, double* h, double* jh)
{
    using namespace nmpc;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double* zt = z + (size_t)t * 17;
    const double* pt = p + (size_t)t * 130;
    const int st = stage[t];
    double zk[NZ], g[NZ], cc[NXI], jcomp[NJC];
    for (int i = 0; i < NZ; i++) zk[i] = zt[i];
    f[t] = objective<double, true>(zk, pt, st == 0, variant == 1 && st == n_stages - 1, g);
    for (int i = 0; i < NZ; i++) grad[(size_t)t * 17 + i] = g[i];
    dynamics<double, true>(zk, pt + 3, cc, jcomp);
    for (int i = 0; i < NXI; i++) c[(size_t)t * 13 + i] = cc[i];
    // expand the compact Jacobian through the very accessor the solver uses (jt_y on unit vectors)
    for (int r = 0; r < NXI; r++) {
        double e[NXI];
        for (int i = 0; i < NXI; i++) e[i] = (i == r) ? 1.0 : 0.0;
        for (int col = 0; col < NZ; col++) jc[(size_t)t * 221 + col * 13 + r] = jt_y<double>(jcomp, e, col);
    }
    for (int j = 0; j < 30; j++) {
        const double a0 = pt[10 + 3 * j], a1 = pt[11 + 3 * j], a2 = pt[12 + 3 * j];
        h[(size_t)t * 30 + j] = a0 * zk[8] + a1 * zk[9] + a2 * zk[10] - pt[100 + j];
        for (int col = 0; col < NZ; col++)
            jh[(size_t)t * 510 + col * 30 + j] = col == 8 ? a0 : (col == 9 ? a1 : (col == 10 ? a2 : 0.0));
    }
}

// ---- FMA-issue probe: the roofline denominator MEASURED_PEAKS.json does not carry (fp64 / fp32 CUDA cores)
template <typename T> __global__ void fma_probe_kernel(int iters, T* sink)
{
    T a0 = T(threadIdx.x) * T(1e-3), a1 = a0 + T(1), a2 = a0 + T(2), a3 = a0 + T(3);
    T a4 = a0 + T(4), a5 = a0 + T(5), a6 = a0 + T(6), a7 = a0 + T(7);
    const T m = T(0.999999), c = T(1e-6);
    for (int i = 0; i < iters; i++) {
        a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
        a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == T(-1)) sink[0] = a0;
}
template <typename T> int fma_probe(double* tflops)
{
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    T* sink;
    CUDA_TRY(cudaMalloc(&sink, sizeof(T)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    const int iters = 1 << 16, blocks = sms * 8, threads = 256;
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        CUDA_TRY(cudaEventRecord(e0));
        fma_probe_kernel<T><<<blocks, threads>>>(iters, sink);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) * 1e-12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    *tflops = best;
    return 0;
}

}  // namespace

// ---- peer-store collation: barrier kernel and the per-rank object (entry points below) ----
namespace {
constexpr int PEER_HDR_BYTES = 256;            // flags[16] | epoch | status, then the collation buffers
constexpr int PEER_EPOCH = 16, PEER_STATUS = 17;
struct PeerSync { unsigned* hdr[nmpc::MAX_PEERS + 1]; int rank, world; };

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// All ranks meet here: thread p tells rank p "rank `rank` has reached epoch e" (a release store into p's header, after
// a system-wide fence: everything this GPU wrote to p before -- the solutions -- is visible there first) and waits for
// p's word in its own header.  The epoch lives on the device, so a CUDA graph can replay the kernel.  A rank that does
// not show up within two seconds sets the status word instead of hanging the GPU (nmpc_peers_status).
__global__ void peer_barrier_kernel(PeerSync ps)
{
    __shared__ unsigned e_sh;
    unsigned* mine = ps.hdr[ps.rank];
    if (threadIdx.x == 0) e_sh = mine[PEER_EPOCH] + 1;
    __syncthreads();
    const unsigned e = e_sh;
    const int p = threadIdx.x;
    if (p < ps.world && p != ps.rank) {
        __threadfence_system();
        st_release_sys(ps.hdr[p] + ps.rank, e);
        const unsigned long long t0 = globaltimer_ns();
        while ((int)(ld_acquire_sys(mine + p) - e) < 0)
            if (globaltimer_ns() - t0 > 2000000000ull) { mine[PEER_STATUS] = 1; break; }
    }
    __syncthreads();
    if (threadIdx.x == 0) mine[PEER_EPOCH] = e;
}
}  // namespace

struct nmpc_peers {
    int rank = 0, world = 1, device = 0;
    size_t z_bytes = 0, info_ints = 0, z_off = 0, info_off = 0, total = 0;
    char* base[nmpc::MAX_PEERS + 1] = {};      // base[rank] = this rank's allocation, the others are IPC mappings
    bool connected = false;
};

extern "C" {

int nmpc_fma_peak_probe(int elem_size, double* tflops)
{
    if (!tflops || (elem_size != 8 && elem_size != 4)) return fail(NMPC_ERR_ARG, "bad argument");
    return elem_size == 8 ? fma_probe<double>(tflops) : fma_probe<float>(tflops);
}

void nmpc_default_opts(nmpc_opts* o)
{
    o->mu0 = 1.0; o->sigma = 0.1; o->mu_floor = 1e-5;
    o->tol_stat = o->tol_eq = o->tol_ineq = o->tol_comp = 1e-4;
    o->kappa_push = 1e-2; o->s_floor = 1e-2;
    o->maxit = 200; o->max_bt = 6;
    o->pc = 0; o->mixed = 0;
}

const char* nmpc_last_error(void) { return g_err; }
const char* nmpc_version(void) { return "nmpc_b200 0.2 (sm_100a; fused warp-per-problem IPM, fp64 and mixed precision)"; }
int nmpc_supported_horizon(int N) { return N == 20 || N == 40; }

long nmpc_smem_bytes(int N, int mcap, int elem_size)
{
    if (!nmpc_supported_horizon(N) || mcap < 0 || mcap > 32) return -1;
    if (elem_size == 8) return N == 20 ? (long)nmpc::Layout<double, 20>::bytes(mcap) : (long)nmpc::Layout<double, 40>::bytes(mcap);
    if (elem_size == 4) return N == 20 ? (long)nmpc::MLayout<20>::bytes(mcap) : (long)nmpc::MLayout<40>::bytes(mcap);
    return -1;
}

long nmpc_smem_bytes_pc(int N, int mcap, int elem_size)
{
    if (!nmpc_supported_horizon(N) || mcap < 0 || mcap > 32 || elem_size != 8) return -1;
    return N == 20 ? (long)nmpc::Layout<double, 20, true>::bytes(mcap) : (long)nmpc::Layout<double, 40, true>::bytes(mcap);
}

int nmpc_solve_batch_f64(int B, int N, int mcap, const double* xinit, const double* z0, const double* hdr,
                         const double* rows, const int* nrows, int variant, const nmpc_opts* opts, double* z_out,
                         int* info_int, double* info_real, void* cuda_stream)
{
    return solve_device(B, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_out, info_int, info_real, cuda_stream);
}
int nmpc_solve_batch_f32(int B, int N, int mcap, const float* xinit, const float* z0, const float* hdr,
                         const float* rows, const int* nrows, int variant, const nmpc_opts* opts, float* z_out,
                         int* info_int, float* info_real, void* cuda_stream)
{
    return solve_mixed(B, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_out, info_int, info_real, cuda_stream, 1);
}
int nmpc_solve_batch_mixed_f64(int B, int N, int mcap, const double* xinit, const double* z0, const double* hdr,
                               const double* rows, const int* nrows, int variant, const nmpc_opts* opts, double* z_out,
                               int* info_int, double* info_real, double* y_out, double* zl_out, double* zu_out,
                               double* lc_out, const int* order, void* cuda_stream)
{
    return solve_mixed(B, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_out, info_int, info_real, cuda_stream, 0,
                       y_out, zl_out, zu_out, lc_out, order);
}
int nmpc_solve_batch_lowlatency_f64(int B, int N, int mcap, const double* xinit, const double* z0, const double* hdr,
                                    const double* rows, const int* nrows, int variant, const nmpc_opts* opts, double* z_out,
                                    int* info_int, double* info_real, double* y_out, double* zl_out, double* zu_out,
                                    double* lc_out, const int* order, void* cuda_stream)
{
    return solve_mixed(B, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_out, info_int, info_real, cuda_stream, 0,
                       y_out, zl_out, zu_out, lc_out, order, true);
}
int nmpc_solve_batch_ex_f64(int B, int N, int mcap, const double* xinit, const double* z0, const double* hdr,
                            const double* rows, const int* nrows, int variant, const nmpc_opts* opts, double* z_out,
                            int* info_int, double* info_real, double* y_out, double* zl_out, double* zu_out,
                            double* lc_out, void* cuda_stream)
{
    return solve_device(B, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_out, info_int, info_real,
                        cuda_stream, y_out, zl_out, zu_out, lc_out);
}
int nmpc_solve_batch_ordered_f64(int B, int N, int mcap, const double* xinit, const double* z0, const double* hdr,
                                 const double* rows, const int* nrows, int variant, const nmpc_opts* opts, double* z_out,
                                 int* info_int, double* info_real, const int* order, void* cuda_stream)
{
    return solve_device(B, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_out, info_int, info_real,
                        cuda_stream, nullptr, nullptr, nullptr, nullptr, order);
}
int nmpc_solve_batch_host_f64(int B, int N, int mcap, const double* xinit, const double* z0, const double* hdr,
                              const double* rows, const int* nrows, int variant, const nmpc_opts* opts,
                              double* z_out, int* info_int, double* info_real)
{
    return solve_host(B, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_out, info_int, info_real, 8, false);
}
int nmpc_solve_batch_host_f32(int B, int N, int mcap, const float* xinit, const float* z0, const float* hdr,
                              const float* rows, const int* nrows, int variant, const nmpc_opts* opts,
                              float* z_out, int* info_int, float* info_real)
{
    return solve_host(B, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_out, info_int, info_real, 4, true);
}
int nmpc_solve_batch_host_mixed_f64(int B, int N, int mcap, const double* xinit, const double* z0, const double* hdr,
                                    const double* rows, const int* nrows, int variant, const nmpc_opts* opts,
                                    double* z_out, int* info_int, double* info_real)
{
    return solve_host(B, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_out, info_int, info_real, 8, true);
}

/* ---- multi-GPU: contiguous sharding + end-of-batch NCCL all-gather ---------------------------------------------- */
int nmpc_comm_unique_id(char id[128])
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    NcclApi* api;
    if (int rc = nccl_api(&api)) return rc;
    if (!id) return fail(NMPC_ERR_ARG, "null pointer argument");
    ncclUniqueId u;
    NCCL_TRY(api, api->GetUniqueId(&u));
    std::memcpy(id, &u, 128);
    return 0;
}
int nmpc_comm_create(int world, int rank, const char id[128], nmpc_comm** out)
{
    NcclApi* api;
    if (int rc = nccl_api(&api)) return rc;
    if (!id || !out || world < 1 || rank < 0 || rank >= world) return fail(NMPC_ERR_ARG, "bad argument: world=%d rank=%d", world, rank);
    ncclUniqueId u;
    std::memcpy(&u, id, 128);
    nmpc_comm* c = new nmpc_comm;
    c->rank = rank; c->world = world; c->owned = true;
    ncclResult_t r = api->CommInitRank(&c->comm, world, u, rank);
    if (r != ncclSuccess) { delete c; return fail(NMPC_ERR_CUDA, "ncclCommInitRank failed: %s", api->GetErrorString(r)); }
    *out = c;
    return 0;
}
int nmpc_comm_wrap(void* nccl_comm, int world, int rank, nmpc_comm** out)
{
    NcclApi* api;
    if (int rc = nccl_api(&api)) return rc;
    if (!nccl_comm || !out || world < 1 || rank < 0 || rank >= world) return fail(NMPC_ERR_ARG, "bad argument");
    nmpc_comm* c = new nmpc_comm;
    c->comm = static_cast<ncclComm_t>(nccl_comm); c->rank = rank; c->world = world; c->owned = false;
    *out = c;
    return 0;
}
int nmpc_comm_rank(const nmpc_comm* c) { return c ? c->rank : -1; }
int nmpc_comm_world(const nmpc_comm* c) { return c ? c->world : -1; }
int nmpc_comm_nccl_version(void)
{
    NcclApi* api;
    if (nccl_api(&api)) return -1;
    int v = 0;
    return api->GetVersion(&v) == ncclSuccess ? v : -1;
}
int nmpc_comm_alloc(nmpc_comm* c, size_t bytes, void** ptr)
{
    NcclApi* api;
    if (int rc = nccl_api(&api)) return rc;
    if (!c || !ptr || bytes == 0) return fail(NMPC_ERR_ARG, "bad argument");
    nmpc_comm::Reg reg{nullptr, nullptr, false};
    if (api->MemAlloc && api->MemFree && api->MemAlloc(&reg.ptr, bytes) == ncclSuccess) reg.nccl_mem = true;
    else CUDA_TRY(cudaMalloc(&reg.ptr, bytes));
    // user-buffer registration: NCCL then reads / writes this buffer directly (no staging copy; NVLS-eligible)
    if (api->CommRegister && api->CommRegister(c->comm, reg.ptr, bytes, &reg.handle) != ncclSuccess) reg.handle = nullptr;
    c->regs.push_back(reg);
    *ptr = reg.ptr;
    return 0;
}
int nmpc_comm_free(nmpc_comm* c, void* ptr)
{
    NcclApi* api;
    if (int rc = nccl_api(&api)) return rc;
    if (!c || !ptr) return fail(NMPC_ERR_ARG, "bad argument");
    for (size_t i = 0; i < c->regs.size(); i++)
        if (c->regs[i].ptr == ptr) {
            if (c->regs[i].handle && api->CommDeregister) api->CommDeregister(c->comm, c->regs[i].handle);
            if (c->regs[i].nccl_mem) api->MemFree(ptr); else cudaFree(ptr);
            c->regs.erase(c->regs.begin() + i);
            return 0;
        }
    return fail(NMPC_ERR_ARG, "pointer was not allocated by nmpc_comm_alloc");
}
int nmpc_comm_destroy(nmpc_comm* c)
{
    if (!c) return 0;
    NcclApi* api;
    if (int rc = nccl_api(&api)) return rc;
    while (!c->regs.empty()) nmpc_comm_free(c, c->regs.back().ptr);
    if (c->owned && c->comm) api->CommDestroy(c->comm);
    delete c;
    return 0;
}

// in-place all-gather of every rank's slice (sendbuff = recvbuff + rank * count): results land where the kernel wrote them
static int collate_inplace(nmpc_comm* c, void* z_all, size_t z_bytes_per_rank, int* info_all, size_t info_ints_per_rank, cudaStream_t st)
{
    NcclApi* api;
    if (int rc = nccl_api(&api)) return rc;
    if (c->world == 1) return 0;
    NCCL_TRY(api, api->GroupStart());
    ncclResult_t r1 = api->AllGather(static_cast<char*>(z_all) + (size_t)c->rank * z_bytes_per_rank, z_all, z_bytes_per_rank, ncclChar, c->comm, st);
    ncclResult_t r2 = info_all ? api->AllGather(info_all + (size_t)c->rank * info_ints_per_rank, info_all, info_ints_per_rank, ncclInt, c->comm, st)
                               : ncclSuccess;
    ncclResult_t r3 = api->GroupEnd();
    if (r1 != ncclSuccess || r2 != ncclSuccess || r3 != ncclSuccess)
        return fail(NMPC_ERR_CUDA, "ncclAllGather failed: %s", api->GetErrorString(r1 != ncclSuccess ? r1 : (r2 != ncclSuccess ? r2 : r3)));
    return 0;
}

static int solve_sharded(nmpc_comm* c, int B_local, int N, int mcap, const void* xinit, const void* z0, const void* hdr, const void* rows,
                         const int* nrows, int variant, const nmpc_opts* opts, void* z_all, int* info_int_all, void* info_real_local,
                         void* stream, size_t esz, bool mixed)
{
    if (!c || !z_all || !info_int_all) return fail(NMPC_ERR_ARG, "null pointer argument");
    if (B_local < 0) return fail(NMPC_ERR_ARG, "B_local < 0");
    const size_t zb = (size_t)B_local * N * 17 * esz, ib = (size_t)B_local * 4;
    if ((zb & 15) != 0) return fail(NMPC_ERR_ARG, "B_local must keep every rank's slice 16-byte aligned (even B_local)");
    char* z_mine = static_cast<char*>(z_all) + (size_t)c->rank * zb;
    int* ii_mine = info_int_all + (size_t)c->rank * ib;
    int rc = mixed ? solve_mixed(B_local, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_mine, ii_mine, info_real_local, stream, esz == 4)
                   : solve_device(B_local, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_mine, ii_mine, info_real_local, stream);
    if (rc) return rc;
    return collate_inplace(c, z_all, zb, info_int_all, ib, reinterpret_cast<cudaStream_t>(stream));
}

int nmpc_solve_batch_sharded_f64(nmpc_comm* comm, int B_local, int N, int mcap, const double* xinit, const double* z0,
                                 const double* hdr, const double* rows, const int* nrows, int variant, const nmpc_opts* opts,
                                 double* z_all, int* info_int_all, double* info_real_local, int mixed, void* cuda_stream)
{
    return solve_sharded(comm, B_local, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_all, info_int_all, info_real_local,
                         cuda_stream, 8, mixed != 0);
}
int nmpc_solve_batch_sharded_f32(nmpc_comm* comm, int B_local, int N, int mcap, const float* xinit, const float* z0,
                                 const float* hdr, const float* rows, const int* nrows, int variant, const nmpc_opts* opts,
                                 float* z_all, int* info_int_all, float* info_real_local, void* cuda_stream)
{
    return solve_sharded(comm, B_local, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_all, info_int_all, info_real_local,
                         cuda_stream, 4, true);
}
int nmpc_collate_inplace(nmpc_comm* comm, void* buf_all, size_t bytes_per_rank, void* cuda_stream)
{
    if (!comm || !buf_all) return fail(NMPC_ERR_ARG, "null pointer argument");
    return collate_inplace(comm, buf_all, bytes_per_rank, nullptr, 0, reinterpret_cast<cudaStream_t>(cuda_stream));
}

/* ---- multi-GPU: collation by peer stores from inside the solve kernel (CUDA IPC over NVLink) ------------------- */
int nmpc_peers_create(int world, int rank, size_t z_bytes_all, size_t info_ints_all, nmpc_peers** out)
{
    if (!out || world < 1 || world > nmpc::MAX_PEERS + 1 || rank < 0 || rank >= world || z_bytes_all == 0)
        return fail(NMPC_ERR_ARG, "bad argument: world=%d (at most %d) rank=%d", world, nmpc::MAX_PEERS + 1, rank);
    nmpc_peers* p = new nmpc_peers;
    p->rank = rank; p->world = world;
    p->z_bytes = z_bytes_all; p->info_ints = info_ints_all;
    p->z_off = PEER_HDR_BYTES;
    p->info_off = p->z_off + ((z_bytes_all + 255) & ~(size_t)255);
    p->total = p->info_off + ((info_ints_all * sizeof(int) + 255) & ~(size_t)255);
    void* mem = nullptr;
    // plain cudaMalloc: the allocation has to be exportable by cudaIpcGetMemHandle (no pool / VMM memory)
    if (cudaGetDevice(&p->device) != cudaSuccess || cudaMalloc(&mem, p->total) != cudaSuccess || cudaMemset(mem, 0, PEER_HDR_BYTES) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess) {
        cudaError_t e = cudaGetLastError();
        if (mem) cudaFree(mem);
        delete p;
        return fail(NMPC_ERR_CUDA, "peer buffer allocation failed: %s", cudaGetErrorString(e));
    }
    p->base[rank] = static_cast<char*>(mem);
    p->connected = (world == 1);
    *out = p;
    return 0;
}
int nmpc_peers_export(nmpc_peers* p, unsigned char handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    if (!p || !handle) return fail(NMPC_ERR_ARG, "null pointer argument");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, p->base[p->rank]));
    std::memcpy(handle, &h, 64);
    return 0;
}
int nmpc_peers_connect(nmpc_peers* p, const unsigned char* handles)
{
    if (!p || !handles) return fail(NMPC_ERR_ARG, "null pointer argument");
    if (p->world == 1) return 0;                                  // nothing to map
    if (p->connected) return fail(NMPC_ERR_ARG, "already connected");
    for (int r = 0; r < p->world; r++) {
        if (r == p->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * 64, 64);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < r; q++)
                if (q != p->rank && p->base[q]) { cudaIpcCloseMemHandle(p->base[q]); p->base[q] = nullptr; }
            return fail(NMPC_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s (peers must be GPUs of one node with P2P access)", r,
                        cudaGetErrorString(e));
        }
        p->base[r] = static_cast<char*>(ptr);
    }
    p->connected = true;
    return 0;
}
void* nmpc_peers_z(nmpc_peers* p) { return p ? p->base[p->rank] + p->z_off : nullptr; }
int* nmpc_peers_info(nmpc_peers* p) { return p ? reinterpret_cast<int*>(p->base[p->rank] + p->info_off) : nullptr; }
int nmpc_peers_rank(const nmpc_peers* p) { return p ? p->rank : -1; }
int nmpc_peers_world(const nmpc_peers* p) { return p ? p->world : -1; }
int nmpc_peers_barrier(nmpc_peers* p, void* cuda_stream)
{
    if (!p) return fail(NMPC_ERR_ARG, "null pointer argument");
    if (!p->connected) return fail(NMPC_ERR_ARG, "nmpc_peers_connect has not been called");
    if (p->world == 1) return 0;
    PeerSync ps;
    for (int r = 0; r < p->world; r++) ps.hdr[r] = reinterpret_cast<unsigned*>(p->base[r]);
    ps.rank = p->rank; ps.world = p->world;
    peer_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(ps);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
int nmpc_peers_status(nmpc_peers* p)
{
    if (!p) return fail(NMPC_ERR_ARG, "null pointer argument");
    unsigned st = 0;
    CUDA_TRY(cudaMemcpy(&st, p->base[p->rank] + PEER_STATUS * sizeof(unsigned), sizeof(unsigned), cudaMemcpyDeviceToHost));
    if (st) return fail(NMPC_ERR_CUDA, "a peer did not reach the barrier within 2 s: the collated results are incomplete");
    return 0;
}
int nmpc_peers_destroy(nmpc_peers* p)
{
    if (!p) return 0;
    cudaDeviceSynchronize();
    for (int r = 0; r < p->world; r++)
        if (r != p->rank && p->base[r]) cudaIpcCloseMemHandle(p->base[r]);
    if (p->base[p->rank]) cudaFree(p->base[p->rank]);
    delete p;
    return 0;
}

static int solve_sharded_p2p(nmpc_peers* p, int B_local, int N, int mcap, const void* xinit, const void* z0, const void* hdr, const void* rows,
                             const int* nrows, int variant, const nmpc_opts* opts, void* info_real_local, void* stream, size_t esz, int mode)
{
    if (!p) return fail(NMPC_ERR_ARG, "null pointer argument");
    if (!p->connected) return fail(NMPC_ERR_ARG, "nmpc_peers_connect has not been called");
    if (B_local < 0 || mode < 0 || mode > 2 || (esz == 4 && mode == 0)) return fail(NMPC_ERR_ARG, "bad argument: B_local=%d mode=%d", B_local, mode);
    const size_t zb = (size_t)B_local * N * 17 * esz, ib = (size_t)B_local * 4;
    if ((zb & 15) != 0) return fail(NMPC_ERR_ARG, "B_local must keep every rank's slice 16-byte aligned (even B_local)");
    if (zb * p->world > p->z_bytes || ib * p->world > p->info_ints)
        return fail(NMPC_ERR_ARG, "collation buffers too small: %zu bytes of z for %d ranks x %d problems", p->z_bytes, p->world, B_local);
    nmpc::PeerOut po;
    for (int r = 0; r < p->world; r++) {
        if (r == p->rank) continue;
        po.z[po.n] = p->base[r] + p->z_off + (size_t)p->rank * zb;
        po.info[po.n] = reinterpret_cast<int*>(p->base[r] + p->info_off) + (size_t)p->rank * ib;
        po.n++;
    }
    char* z_mine = p->base[p->rank] + p->z_off + (size_t)p->rank * zb;
    int* ii_mine = reinterpret_cast<int*>(p->base[p->rank] + p->info_off) + (size_t)p->rank * ib;
    // barrier 1: every rank has finished reading the previous batch's results (its readers precede this call on its
    // stream), so the peers may be overwritten; barrier 2: every rank's results have landed everywhere
    if (int rc = nmpc_peers_barrier(p, stream)) return rc;
    int rc = mode == 0 ? solve_device(B_local, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_mine, ii_mine, info_real_local, stream,
                                      nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, &po)
                       : solve_mixed(B_local, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, z_mine, ii_mine, info_real_local, stream,
                                     esz == 4, nullptr, nullptr, nullptr, nullptr, nullptr, mode == 2, &po);
    if (rc) return rc;
    return nmpc_peers_barrier(p, stream);
}
int nmpc_solve_batch_sharded_p2p_f64(nmpc_peers* peers, int B_local, int N, int mcap, const double* xinit, const double* z0,
                                     const double* hdr, const double* rows, const int* nrows, int variant, const nmpc_opts* opts,
                                     double* info_real_local, int mode, void* cuda_stream)
{
    return solve_sharded_p2p(peers, B_local, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, info_real_local, cuda_stream, 8, mode);
}
int nmpc_solve_batch_sharded_p2p_f32(nmpc_peers* peers, int B_local, int N, int mcap, const float* xinit, const float* z0,
                                     const float* hdr, const float* rows, const int* nrows, int variant, const nmpc_opts* opts,
                                     float* info_real_local, void* cuda_stream)
{
    return solve_sharded_p2p(peers, B_local, N, mcap, xinit, z0, hdr, rows, nrows, variant, opts, info_real_local, cuda_stream, 4, 1);
}

int nmpc_backsolve_factor_words(void) { return nmpc::FAC_WORDS; }
long nmpc_backsolve_algorithmic_bytes(int N, int elem_size) { return (long)N * (nmpc::FAC_WORDS + 2 * (17 + 13)) * elem_size; }

int nmpc_riccati_factor_f64(int B, int N, const double* phi, const double* jc, double* fac, int* status, void* stream)
{
    return factor_device<double>(B, N, phi, jc, fac, status, stream);
}
int nmpc_riccati_factor_f32(int B, int N, const float* phi, const float* jc, float* fac, int* status, void* stream)
{
    return factor_device<float>(B, N, phi, jc, fac, status, stream);
}
int nmpc_kkt_backsolve_f64(int B, int N, const double* fac, const double* g, const double* d, double* dz, double* y, void* stream)
{
    return backsolve_device<double>(B, N, fac, g, d, dz, y, stream);
}
int nmpc_kkt_backsolve_f32(int B, int N, const float* fac, const float* g, const float* d, float* dz, float* y, void* stream)
{
    return backsolve_device<float>(B, N, fac, g, d, dz, y, stream);
}

int nmpc_pack_params_f64(int B, int N, int P, int M, int mcap, const double* ref_pos, const double* ref_yaw,
                         const double* ext_acc, const double* ellipsoid, const double* poly_A, const double* poly_b,
                         const int* poly_m, const int* poly_idx, const double* weights5, double* hdr, double* rows,
                         int* nrows, void* stream)
{
    if (B < 0 || N <= 0 || P <= 0 || M <= 0 || mcap <= 0 || mcap > 32) return fail(NMPC_ERR_ARG, "bad argument");
    if (B == 0) return 0;
    if (!ref_pos || !ref_yaw || !ext_acc || !ellipsoid || !poly_A || !poly_b || !poly_m || !poly_idx || !weights5 || !hdr || !rows || !nrows)
        return fail(NMPC_ERR_ARG, "null pointer argument");
    nmpc::PackParams q{B, N, P, M, mcap, ref_pos, ref_yaw, ext_acc, ellipsoid, poly_A, poly_b, poly_m, poly_idx,
                       weights5[0], weights5[1], weights5[2], weights5[3], weights5[4], hdr, rows, nrows};
    const int n = B * N;
    nmpc::pack_params_kernel<<<(n + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(q);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int nmpc_select_corridors_f64(int B, int N, int M, int P, int R, const double* cloud, long long cloud_stride,
                              const int* cloud_n, const double* ref_pos, const double* ref_yaw, const double* ellipsoid,
                              const double* bbox3, double* poly_A, double* poly_b, int* poly_m, int* poly_idx, int* n_poly,
                              int* overflow, void* stream)
{
    if (B < 0 || N <= 0 || M < 0 || P <= 0 || R < 7 || (cloud_stride != 0 && cloud_stride < 3LL * M))
        return fail(NMPC_ERR_ARG, "bad argument: B=%d N=%d M=%d P=%d R=%d stride=%lld", B, N, M, P, R, cloud_stride);
    if (B == 0) return 0;
    if ((M > 0 && !cloud) || !cloud_n || !ref_pos || !ref_yaw || !ellipsoid || !poly_A || !poly_b || !poly_m || !poly_idx ||
        !n_poly || !overflow)
        return fail(NMPC_ERR_ARG, "null pointer argument");
    const size_t smem = ((size_t)(M + 15) & ~(size_t)15) + (size_t)R * 4 * sizeof(double);
    if (smem > 200 * 1024) return fail(NMPC_ERR_ARG, "cloud too large for the per-agent flag array: M=%d", M);
    static thread_local size_t configured[kMaxDevices] = {};
    if (smem > 48 * 1024)
        if (int rc = ensure_smem(nmpc::corridor_select_kernel, smem, configured)) return rc;
    nmpc::CorridorParams q{B, N, M, P, R, cloud, cloud_stride, cloud_n, ref_pos, ref_yaw, ellipsoid,
                           bbox3 ? bbox3[0] : 2.0, bbox3 ? bbox3[1] : 2.0, bbox3 ? bbox3[2] : 1.0,
                           poly_A, poly_b, poly_m, poly_idx, n_poly, overflow};
    nmpc::corridor_select_kernel<<<B, nmpc::COR_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(q);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

void nmpc_default_ellipsoid_consts(nmpc_ellipsoid_consts* c)
{
    c->mass = 0.745319; c->drag = 0.33; c->ego_r = 0.27; c->ego_h = 0.0425;   // rotors_sim.launch:53-70
    c->ext_noise_bound = 0.5; c->epsilon = 0.06; c->Ts = 0.05;                 // nmpc_solver.cpp:80, nmpc_utils.h:188
}

int nmpc_propagate_ellipsoids_f64(int B, int N, const double* z, const nmpc_ellipsoid_consts* consts, double* ellipsoid,
                                  void* stream)
{
    if (B < 0 || N <= 0) return fail(NMPC_ERR_ARG, "bad argument: B=%d N=%d", B, N);
    if (B == 0) return 0;
    if (!z || !ellipsoid) return fail(NMPC_ERR_ARG, "null pointer argument");
    nmpc_ellipsoid_consts c;
    if (consts) c = *consts; else nmpc_default_ellipsoid_consts(&c);
    if (!(c.mass > 0) || !(c.Ts > 0) || !(c.ext_noise_bound > 0) || !(c.epsilon > 0))
        return fail(NMPC_ERR_ARG, "bad ellipsoid constants");
    nmpc::EllipsoidParams q{B, N, z, ellipsoid, c.mass, c.drag, c.ego_r, c.ego_h, c.ext_noise_bound, c.epsilon, c.Ts};
    const size_t smem = nmpc::ellipsoid_smem_bytes(N);
    if (smem > 200 * 1024) return fail(NMPC_ERR_ARG, "horizon too long for the ellipsoid kernel: N=%d", N);
    static thread_local size_t configured[kMaxDevices] = {};
    if (smem > 48 * 1024)
        if (int rc = ensure_smem(nmpc::ellipsoid_propagate_kernel, smem, configured)) return rc;
    nmpc::ellipsoid_propagate_kernel<<<(B + nmpc::ELL_WARPS - 1) / nmpc::ELL_WARPS, 32 * nmpc::ELL_WARPS, smem,
                                       reinterpret_cast<cudaStream_t>(stream)>>>(q);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int nmpc_wrap_yaw_f64(int B, int N, double* z, void* stream)
{
    if (B < 0 || N <= 0) return fail(NMPC_ERR_ARG, "bad argument: B=%d N=%d", B, N);
    if (B == 0) return 0;
    if (!z) return fail(NMPC_ERR_ARG, "null pointer argument");
    const int n = B * N;
    nmpc::wrap_yaw_kernel<<<(n + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(n, z);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int nmpc_sample_reference_f64(int B, int N, int P, double Ts, const double* kino_path, const int* kino_size,
                              const double* t_off, const double* last_yaw, const double* pos1, double* ref_pos,
                              double* ref_yaw, int* hard_to_follow, void* stream)
{
    if (B < 0 || N <= 0 || P <= 0 || !(Ts > 0)) return fail(NMPC_ERR_ARG, "bad argument: B=%d N=%d P=%d Ts=%g", B, N, P, Ts);
    if (B == 0) return 0;
    if (!kino_path || !kino_size || !t_off || !last_yaw || !ref_pos || !ref_yaw) return fail(NMPC_ERR_ARG, "null pointer argument");
    nmpc::SampleParams q{B, N, P, Ts, kino_path, kino_size, t_off, last_yaw, pos1, ref_pos, ref_yaw, hard_to_follow};
    nmpc::sample_reference_kernel<<<(B + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(q);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int nmpc_shift_warm_start_f64(int B, int N, const double* z_prev, double* xinit, double* z0, int wrap_yaw, void* stream)
{
    if (B < 0 || N <= 1) return fail(NMPC_ERR_ARG, "bad argument");
    if (B == 0) return 0;
    if (!z_prev || !xinit || !z0) return fail(NMPC_ERR_ARG, "null pointer argument");
    const int n = B * N;
    nmpc::shift_warm_start_kernel<<<(n + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(B, N, z_prev, xinit, z0, wrap_yaw);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int nmpc_adopt_plans_f64(int B, int N, const double* z_new, const int* info_int, const int* accept, const double* odom,
                         double* z_prev, int* cold, int wrap_yaw, void* stream)
{
    if (B < 0 || N <= 1 || N > 256) return fail(NMPC_ERR_ARG, "bad argument: B=%d N=%d", B, N);
    if (B == 0) return 0;
    if (!z_new || !z_prev || (!info_int && !accept)) return fail(NMPC_ERR_ARG, "null pointer argument");
    const int apb = 256 / N > 0 ? 256 / N : 1;                       // whole agents per block
    nmpc::adopt_plans_kernel<<<(B + apb - 1) / apb, apb * N, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        B, N, z_new, info_int, accept, odom, z_prev, cold, wrap_yaw);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int nmpc_rank_longest_first(int B, const int* info_int, int* order, void* stream)
{
    if (B < 0 || B > 12288) return fail(NMPC_ERR_ARG, "bad argument: B=%d (the keys of at most 12288 agents fit the default 48 KB of shared memory)", B);
    if (B == 0) return 0;
    if (!info_int || !order) return fail(NMPC_ERR_ARG, "null pointer argument");
    nmpc::rank_longest_first_kernel<<<(B + nmpc::RANK_THREADS - 1) / nmpc::RANK_THREADS, nmpc::RANK_THREADS, (size_t)B * sizeof(int),
                                      reinterpret_cast<cudaStream_t>(stream)>>>(B, info_int, order);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int nmpc_model_eval_host_f64(int n, const double* z, const double* p, const int* stage, int n_stages, int variant,
                             double* f, double* grad, double* c, double* jc, double* h, double* jh)
{
    if (n <= 0) return 0;
    double *dz, *dp, *df, *dg, *dc, *djc, *dh, *djh;
    int* ds;
    CUDA_TRY(cudaMalloc(&dz, (size_t)n * 17 * 8)); CUDA_TRY(cudaMalloc(&dp, (size_t)n * 130 * 8));
    CUDA_TRY(cudaMalloc(&ds, (size_t)n * 4)); CUDA_TRY(cudaMalloc(&df, (size_t)n * 8));
    CUDA_TRY(cudaMalloc(&dg, (size_t)n * 17 * 8)); CUDA_TRY(cudaMalloc(&dc, (size_t)n * 13 * 8));
    CUDA_TRY(cudaMalloc(&djc, (size_t)n * 221 * 8)); CUDA_TRY(cudaMalloc(&dh, (size_t)n * 30 * 8));
    CUDA_TRY(cudaMalloc(&djh, (size_t)n * 510 * 8));
    CUDA_TRY(cudaMemcpy(dz, z, (size_t)n * 17 * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dp, p, (size_t)n * 130 * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(ds, stage, (size_t)n * 4, cudaMemcpyHostToDevice));
    model_eval_kernel<<<(n + 63) / 64, 64>>>(n, dz, dp, ds, n_stages, variant, df, dg, dc, djc, dh, djh);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(f, df, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(grad, dg, (size_t)n * 17 * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(c, dc, (size_t)n * 13 * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(jc, djc, (size_t)n * 221 * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(h, dh, (size_t)n * 30 * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(jh, djh, (size_t)n * 510 * 8, cudaMemcpyDeviceToHost));
    cudaFree(dz); cudaFree(dp); cudaFree(ds); cudaFree(df); cudaFree(dg); cudaFree(dc); cudaFree(djc); cudaFree(dh); cudaFree(djh);
    return 0;
}

// ---- the reference's solver symbols ---------------------------------------------------------------
// Each exists under two names bound to one internal function: the reference's own name (a planner that links
// libnmpc_b200.so directly) and a library-private one (what the stub archives libFORCESNLPsolver_{normal,final}.a,
// host/forces_stub.c, resolve with dlsym -- the stub itself defines the reference name in the executable, and symbol
// interposition must not send the library's call back into it).
solver_int32_default FORCESNLPsolver_normal_solve(FORCESNLPsolver_normal_params* params, FORCESNLPsolver_normal_output* output,
                                                  FORCESNLPsolver_normal_info* info, FILE* fs,
                                                  FORCESNLPsolver_normal_extfunc /*ignored: device model built in*/)
{
    return forces_entry<FORCESNLPsolver_normal_params, FORCESNLPsolver_normal_output, FORCESNLPsolver_normal_info>(params, output, info, fs, 0);
}
solver_int32_default nmpc_forces_normal_solve(FORCESNLPsolver_normal_params* params, FORCESNLPsolver_normal_output* output,
                                              FORCESNLPsolver_normal_info* info, FILE* fs, FORCESNLPsolver_normal_extfunc)
{
    return forces_entry<FORCESNLPsolver_normal_params, FORCESNLPsolver_normal_output, FORCESNLPsolver_normal_info>(params, output, info, fs, 0);
}
solver_int32_default FORCESNLPsolver_final_solve(FORCESNLPsolver_final_params* params, FORCESNLPsolver_final_output* output,
                                                 FORCESNLPsolver_final_info* info, FILE* fs, FORCESNLPsolver_final_extfunc /*ignored*/)
{
    return forces_entry<FORCESNLPsolver_final_params, FORCESNLPsolver_final_output, FORCESNLPsolver_final_info>(params, output, info, fs, 1);
}
solver_int32_default nmpc_forces_final_solve(FORCESNLPsolver_final_params* params, FORCESNLPsolver_final_output* output,
                                             FORCESNLPsolver_final_info* info, FILE* fs, FORCESNLPsolver_final_extfunc)
{
    return forces_entry<FORCESNLPsolver_final_params, FORCESNLPsolver_final_output, FORCESNLPsolver_final_info>(params, output, info, fs, 1);
}

}  // extern "C"
