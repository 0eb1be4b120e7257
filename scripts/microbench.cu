// Dependent-chain latencies of the instructions the per-warp sweeps are made of (single warp, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench scripts/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void chain(double* out, long long* cyc, int iters)
{
    __shared__ double sm[64];
    const int lane = threadIdx.x;
    double a = 1.0 + lane * 1e-9, b = 1.0000001, c = 1e-9;
    float fa = 1.0f + lane * 1e-6f, fb = 1.000001f, fc = 1e-6f;
    sm[lane] = a; sm[lane + 32] = b;
    __syncwarp();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (MODE == 0) a = fma(a, b, c);                                     // DFMA
            if (MODE == 1) a = a + b;                                            // DADD
            if (MODE == 2) a = __shfl_xor_sync(0xffffffffu, a, 1);               // 64-bit shuffle (2 SHFL)
            if (MODE == 3) { a = sm[(lane + (int)a) & 31]; }                     // LDS (address dependent; includes F2I)
            if (MODE == 4) { sm[lane] = a; __syncwarp(); a = sm[(lane + 1) & 31] ; __syncwarp(); }   // STS -> LDS round trip
            if (MODE == 5) fa = fmaf(fa, fb, fc);                                // FFMA
            if (MODE == 6) fa = __shfl_xor_sync(0xffffffffu, fa, 1);             // 32-bit shuffle
            if (MODE == 7) a = fma(a, b, c) + __shfl_xor_sync(0xffffffffu, a, 1); // DFMA + shuffle + DADD
            if (MODE == 8) a = a * b;                                            // DMUL
        }
    }
    long long t1 = clock64();
    if (lane == 0) { cyc[0] = t1 - t0; }
    out[lane] = a + fa;
}

template <int MODE> void run(const char* name, double* out, long long* cyc)
{
    const int iters = 2000;
    chain<MODE><<<1, 32>>>(out, cyc, iters);
    chain<MODE><<<1, 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-28s %.1f cycles per dependent op\n", name, (double)h / (iters * 16.0));
}

int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 64 * sizeof(double)); cudaMalloc(&cyc, 8);
    run<0>("DFMA", out, cyc);
    run<1>("DADD", out, cyc);
    run<8>("DMUL", out, cyc);
    run<2>("SHFL 64-bit (2 x SHFL)", out, cyc);
    run<3>("LDS (+F2I +IADD)", out, cyc);
    run<4>("STS -> syncwarp -> LDS", out, cyc);
    run<5>("FFMA", out, cyc);
    run<6>("SHFL 32-bit", out, cyc);
    run<7>("DFMA + SHFL64 + DADD", out, cyc);
    return 0;
}
