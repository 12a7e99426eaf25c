/*
 * util.cu -- support kernel of libchefsi_b200.so: start vectors.
 *
 * fill_random     the synthetic start vectors of the benchmark workload: a
 *                 counter-based U(-0.5,0.5) generator (splitmix64 finaliser) keyed on (seed, global
 *                 column, element index), the interval Init_orbital draws from
 *                 (orbitalElecDensInit.c:388-392).  oracle_random_value() and
 *                 problem.random_columns() are its CPU twins.
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

/* one thread per x-run of a (j,k) row keeps index math cheap; blockIdx.y = column */
template <int WORDS> /* doubles per element: 1 real, 2 complex */
__global__ void fill_random_kernel(double *__restrict__ buf, const Layout L, long long first_col,
                                   unsigned long long seed)
{
    const int n = blockIdx.y;
    const uint64_t key = mix64(seed + 0x632BE59BD9B4E019ULL * (uint64_t)(first_col + n + 1));
    double *col = buf + (size_t)n * L.ld * WORDS;
    const size_t Nd = (size_t)L.Nx * L.Ny * L.Nz;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < Nd; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e % L.Nx);
        const size_t r = e / L.Nx;
        const int j = (int)(r % L.Ny), k = (int)(r / L.Ny);
        const size_t p = lay_pos(L, i, j, k);
#pragma unroll
        for (int w = 0; w < WORDS; w++) {
            const uint64_t h = mix64(key + (uint64_t)(e * WORDS + w));
            col[p * WORDS + w] = (double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5;
        }
    }
}

int check(chefsi_ctx *ctx, const char *what, int launched)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "%s launch: %s", what, cudaGetErrorString(e)); return -1; }
    return launched;
}

}  // namespace

int launch_fill_random(chefsi_ctx *ctx, void *buf, int ncol, long long first_col, unsigned long long seed,
                       bool is_complex)
{
    if (ncol <= 0) return 0;
    int launched = 0;
    const size_t words = is_complex ? 2 : 1;
    for (int c0 = 0; c0 < ncol; c0 += 65535) {
        const int nc = (ncol - c0 < 65535) ? ncol - c0 : 65535;
        dim3 grid(2 * ctx->num_sms, (unsigned)nc);
        double *b = (double *)buf + (size_t)c0 * ctx->lay.ld * words;
        if (is_complex) fill_random_kernel<2><<<grid, 256, 0, ctx->stream>>>(b, ctx->lay, first_col + c0, seed);
        else fill_random_kernel<1><<<grid, 256, 0, ctx->stream>>>(b, ctx->lay, first_col + c0, seed);
        launched++;
    }
    return check(ctx, "fill_random", launched);
}
