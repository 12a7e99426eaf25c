/*
 * util.cu -- small support kernels of libchefsi_b200.so.
 *
 * fill_random: the synthetic start vectors of the benchmark workload, a
 * counter-based U(-0.5,0.5) generator (splitmix64 finaliser) keyed on (seed, global column,
 * element index), the same interval Init_orbital draws from (orbitalElecDensInit.c:388-392).
 * oracle_random_value() in oracle/chefsi_oracle.c and problem.random_columns() are its CPU twins.
 */
#include "chefsi_internal.h"

namespace {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__global__ void fill_random_kernel(double *__restrict__ buf, size_t n_per_col, size_t ld, long long first_col,
                                   unsigned long long seed)
{
    const int n = blockIdx.y;
    const uint64_t key = mix64(seed + 0x632BE59BD9B4E019ULL * (uint64_t)(first_col + n + 1));
    double *col = buf + (size_t)n * ld;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_col; i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t h = mix64(key + (uint64_t)i);
        col[i] = (double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5;
    }
}

}  // namespace

int launch_fill_random(chefsi_ctx *ctx, double *buf, size_t n_per_col, size_t ld_doubles, int ncol, long long first_col,
                       unsigned long long seed)
{
    if (ncol <= 0) return 0;
    int launched = 0;
    for (int c0 = 0; c0 < ncol; c0 += 65535) {
        const int nc = (ncol - c0 < 65535) ? ncol - c0 : 65535;
        dim3 grid(2 * ctx->num_sms, (unsigned)nc);
        fill_random_kernel<<<grid, 256, 0, ctx->stream>>>(buf + (size_t)c0 * ld_doubles, n_per_col, ld_doubles,
                                                          first_col + c0, seed);
        launched++;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "fill_random launch: %s", cudaGetErrorString(e)); return -1; }
    return launched;
}
