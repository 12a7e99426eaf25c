// Microbenchmark: FP64 DFMA vs DMMA (mma.sync.m8n8k4.f64) issue rate on sm_100a, and their overlap.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE> __global__ void k(double *out, int iters, double a, double b)
{
    double c[16];
    for (int i = 0; i < 16; i++) c[i] = threadIdx.x * 1e-3 + i;
    double f[8];
    for (int i = 0; i < 8; i++) f[i] = i + threadIdx.x;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) dmma(c[2 * i], c[2 * i + 1], a, b);
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) f[i] = fma(f[i], a, b);
        }
    }
    double s = 0;
    for (int i = 0; i < 16; i++) s += c[i];
    for (int i = 0; i < 8; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, int warps)
{
    double *out; cudaMalloc(&out, 148 * 1024 * 8);
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, warps * 32>>>(out, 100, 1.0000001, 0.999999);
    cudaEventRecord(e0);
    k<MODE><<<148, warps * 32>>>(out, iters, 1.0000001, 0.999999);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double nd = (MODE != 1) ? 148.0 * warps * iters * 8 : 0, nf = (MODE != 0) ? 148.0 * warps * iters * 8 : 0;
    double tf = (nd * 8 * 8 * 4 * 2 + nf * 32 * 2) / (ms * 1e-3) / 1e12;
    printf("%-12s warps/SM %2d: %.3f ms  %.1f TFLOP/s  (DMMA/clk/SM %.3f, DFMA warp-instr/clk/SM %.3f at 1.965GHz)\n", name, warps, ms, tf,
           nd / 148 / (ms * 1e-3 * 1.965e9), nf / 148 / (ms * 1e-3 * 1.965e9));
    cudaFree(out);
}
int main()
{
    for (int w : {4, 8, 16, 32}) { run<0>("DMMA", w); run<1>("DFMA", w); run<2>("DMMA+DFMA", w); }
    return 0;
}
