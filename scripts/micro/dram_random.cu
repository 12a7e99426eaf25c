// HBM bandwidth for scattered accesses of CHUNK bytes (read, and read+write back), to decide whether the projector
// kernel's sphere gather (x-runs of ~64 B in a different DRAM page each) sits at the memory's random-access rate.
// usage: dram_random  (prints GB/s per chunk size)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// LANES lanes cooperate on one chunk of LANES*8 bytes; every iteration a new random chunk
template <int LANES, bool WRITE>
__global__ void scatter_kernel(double *buf, size_t nchunks, int iters, double *sink)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int group = tid / LANES, l = tid % LANES;
    double acc = 0;
#pragma unroll 8
    for (int it = 0; it < iters; it++) {
        const size_t c = mix((uint64_t)group * 1315423911ULL + it) & (nchunks - 1); /* nchunks is a power of two */
        double *p = buf + c * LANES + l;
        const double v = *p;
        if (WRITE) *p = v + 1.0; else acc += v;
    }
    if (!WRITE && acc == 123.456) *sink = acc;
}

template <int LANES, bool WRITE>
void run(double *buf, size_t bytes, double *sink)
{
    const size_t nchunks = bytes / (LANES * 8);
    const int threads = 256, blocks = 148 * 8, iters = 256;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    scatter_kernel<LANES, WRITE><<<blocks, threads>>>(buf, nchunks, iters, sink);
    cudaEventRecord(a);
    for (int r = 0; r < 5; r++) scatter_kernel<LANES, WRITE><<<blocks, threads>>>(buf, nchunks, iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double useful = 5.0 * blocks * threads * (double)iters * 8 * (WRITE ? 2 : 1);
    printf("chunk %4d B  %s : %.0f GB/s useful\n", LANES * 8, WRITE ? "read+write" : "read      ", useful / ms / 1e6);
}

int main()
{
    const size_t bytes = (size_t)16 << 30;
    double *buf, *sink;
    cudaMalloc(&buf, bytes); cudaMalloc(&sink, 8);
    cudaMemset(buf, 0, bytes);
    run<4, false>(buf, bytes, sink);  run<8, false>(buf, bytes, sink);  run<16, false>(buf, bytes, sink);  run<32, false>(buf, bytes, sink);
    run<4, true>(buf, bytes, sink);   run<8, true>(buf, bytes, sink);   run<16, true>(buf, bytes, sink);   run<32, true>(buf, bytes, sink);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
