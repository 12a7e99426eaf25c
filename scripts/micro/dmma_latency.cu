// Dependent-chain latency of DMMA (m8n8k4 f64) and DFMA on sm_100a; one warp per SM, clock64 timing.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CHAINS> __global__ void kd(double *out, long long *cyc, int iters, double a, double b)
{
    double c[2 * CHAINS];
    for (int i = 0; i < 2 * CHAINS; i++) c[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) dmma(c[2 * i], c[2 * i + 1], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 2 * CHAINS; i++) s += c[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CHAINS> __global__ void kf(double *out, long long *cyc, int iters, double a, double b)
{
    double c[CHAINS];
    for (int i = 0; i < CHAINS; i++) c[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) c[i] = fma(c[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CHAINS; i++) s += c[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    const int iters = 4096;
#define RUN(K, N, name) K<N><<<1, 32>>>(out, cyc, iters, 1.0000001, 0.5); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%s chains=%d: %.2f cycles per instruction (1 warp)\n", name, N, (double)h / iters / N);
    RUN(kd, 1, "DMMA") RUN(kd, 2, "DMMA") RUN(kd, 4, "DMMA") RUN(kd, 8, "DMMA")
    RUN(kf, 1, "DFMA") RUN(kf, 2, "DFMA") RUN(kf, 4, "DFMA") RUN(kf, 8, "DFMA")
    return 0;
}
