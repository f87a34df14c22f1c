// FP64 pipe microbenchmark for B200 (sm_100a): dependent-chain latency of DFMA / DMUL / MUFU.RCP64H + refinement,
// and throughput of W warps per SM sub-partition with C independent chains each.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void chain_kernel(double* out, long long* cyc, int iters, double a, double b)
{
    double x[C];
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = threadIdx.x + c;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int c = 0; c < C; ++c) x[c] = fma(x[c], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += x[c];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__device__ __forceinline__ double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double e2 = fma(e, e, e);
    return fma(y, e2, y);
}

__global__ void rcp_kernel(double* out, long long* cyc, int iters, double a)
{
    double x = 1.0 + threadIdx.x * 1e-3;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = fast_rcp(x) + a;
    }
    const long long t1 = clock64();
    if (x == 123.456) out[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void div_kernel(double* out, long long* cyc, int iters, double a)
{
    double x = 1.0 + threadIdx.x * 1e-3;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) x = a / x + a;
    }
    const long long t1 = clock64();
    if (x == 123.456) out[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int C>
void run(const char* name, int threads, double* d_out, long long* d_cyc)
{
    const int iters = 256;
    chain_kernel<C><<<1, threads>>>(d_out, d_cyc, iters, 0.999999, 1e-9);
    chain_kernel<C><<<1, threads>>>(d_out, d_cyc, iters, 0.999999, 1e-9);
    long long c = 0;
    cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double per = (double)c / (iters * 16.0);
    printf("%s threads=%4d chains=%d: %.2f cycles per round of %d DFMA/thread -> %.2f cyc per DFMA per warp; "
           "SMSP DFMA warp-inst/cycle = %.3f\n", name, threads, C, per, C, per / C,
           (threads / 128.0 > 1 ? threads / 128.0 : (threads >= 32 ? 1.0 : 0.0)) * C / per);
}

int main()
{
    double* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 8); cudaMalloc(&d_cyc, 8);
    // one warp: dependent latency
    run<1>("dfma", 32, d_out, d_cyc);
    run<2>("dfma", 32, d_out, d_cyc);
    run<4>("dfma", 32, d_out, d_cyc);
    run<8>("dfma", 32, d_out, d_cyc);
    // one warp per SMSP (128 threads) .. 8 warps per SMSP (1024 threads)
    for (int t : {128, 256, 384, 512, 768, 1024}) {
        run<1>("dfma", t, d_out, d_cyc);
        run<2>("dfma", t, d_out, d_cyc);
        run<4>("dfma", t, d_out, d_cyc);
    }
    long long c = 0;
    rcp_kernel<<<1, 32>>>(d_out, d_cyc, 256, 1e-3);
    rcp_kernel<<<1, 32>>>(d_out, d_cyc, 256, 1e-3);
    cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("fast_rcp + DADD dependent chain: %.1f cycles per (rcp+add)\n", (double)c / (256 * 16.0));
    div_kernel<<<1, 32>>>(d_out, d_cyc, 256, 1e-3);
    div_kernel<<<1, 32>>>(d_out, d_cyc, 256, 1e-3);
    cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("a/x + a dependent chain: %.1f cycles per (div+add)\n", (double)c / (256 * 16.0));
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
