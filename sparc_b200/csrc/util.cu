/*
 * util.cu -- support kernels of libchefsi_b200.so: layout conversion, halo pads, start vectors.
 *
 * pack / unpack   dense reference layout (column n at n*ld, x fastest; eigenSolver.c:744) <-> the
 *                 internal device layout (chefsi_internal.h: Layout), which pads every xy-plane with
 *                 the halo the streaming kernel's TMA boxes read.
 * halo_prepare    fills those pads with what the reference's x_ex copy would hold at np = 1:
 *                 periodic images (times the Bloch phase exp(i k.L) for complex data) or zeros on
 *                 Dirichlet faces (lapVecRoutines.c:536-577, lapVecRoutinesKpt.c:370-462).
 * fill_random     the synthetic start vectors of the benchmark workload: a
 *                 counter-based U(-0.5,0.5) generator (splitmix64 finaliser) keyed on (seed, global
 *                 column, element index), the interval Init_orbital draws from
 *                 (orbitalElecDensInit.c:388-392).  oracle_random_value() and
 *                 problem.random_columns() are its CPU twins.
 */
#include "chefsi_internal.h"
#include "cplx.cuh"

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

template <typename T, bool TO_PACKED>
__global__ void repack_kernel(T *__restrict__ packed, T *__restrict__ dense, const Layout L, size_t ld_dense)
{
    const int n = blockIdx.y;
    T *pc = packed + (size_t)n * L.ld;
    T *dc = dense + (size_t)n * ld_dense;
    const size_t Nd = (size_t)L.Nx * L.Ny * L.Nz;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < Nd; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e % L.Nx);
        const size_t r = e / L.Nx;
        const int j = (int)(r % L.Ny), k = (int)(r / L.Ny);
        const size_t p = lay_pos(L, i, j, k);
        if (TO_PACKED) pc[p] = dc[e]; else dc[e] = pc[p];
    }
}

struct HaloDesc {
    int bcx, bcy;
    double ph_re[9], ph_im[9]; /* index (oy+1)*3 + (ox+1) */
};

template <typename T>
__global__ void halo_prepare_kernel(T *__restrict__ buf, const Layout L, const HaloDesc h, int zero_only)
{
    const int n = blockIdx.y;
    T *col = buf + (size_t)n * L.ld;
    const int padw = 2 * L.px, padh = 2 * L.py;
    /* pad elements of one plane: two full-width strips of py rows + two px-wide side strips */
    const size_t per_plane = (size_t)padh * L.Nxp + (size_t)padw * L.Ny;
    const size_t total = per_plane * L.Nz;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(e / per_plane);
        size_t q = e - (size_t)k * per_plane;
        int ip, jp;
        if (q < (size_t)padh * L.Nxp) {
            jp = (int)(q / L.Nxp);
            ip = (int)(q - (size_t)jp * L.Nxp);
            if (jp >= L.py) jp += L.Ny; /* upper strip */
        } else {
            q -= (size_t)padh * L.Nxp;
            const int jj = (int)(q / padw);
            ip = (int)(q - (size_t)jj * padw);
            if (ip >= L.px) ip += L.Nx;
            jp = L.py + jj;
        }
        int i = ip - L.px, j = jp - L.py, ox = 0, oy = 0;
        if (i < 0) { i += L.Nx; ox = -1; } else if (i >= L.Nx) { i -= L.Nx; ox = 1; }
        if (j < 0) { j += L.Ny; oy = -1; } else if (j >= L.Ny) { j -= L.Ny; oy = 1; }
        T v = cplx::zero<T>();
        const bool dead = zero_only || (ox && h.bcx) || (oy && h.bcy) || i < 0 || i >= L.Nx || j < 0 || j >= L.Ny;
        if (!dead) {
            v = col[lay_pos(L, i, j, k)];
            const int pq = (oy + 1) * 3 + (ox + 1);
            v = cplx::mul_phase(v, h.ph_re[pq], h.ph_im[pq]);
        }
        col[((size_t)k * L.Nyp + jp) * L.Nxp + ip] = v;
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

template <typename T, bool TO_PACKED>
static int repack_t(chefsi_ctx *ctx, void *packed, void *dense, size_t ld_dense, int ncol)
{
    int launched = 0;
    for (int c0 = 0; c0 < ncol; c0 += 65535) {
        const int nc = (ncol - c0 < 65535) ? ncol - c0 : 65535;
        dim3 grid(2 * ctx->num_sms, (unsigned)nc);
        repack_kernel<T, TO_PACKED><<<grid, 256, 0, ctx->stream>>>((T *)packed + (size_t)c0 * ctx->lay.ld,
                                                                   (T *)dense + (size_t)c0 * ld_dense, ctx->lay, ld_dense);
        launched++;
    }
    return check(ctx, "repack", launched);
}

int launch_pack(chefsi_ctx *ctx, const void *dense, size_t ld_dense, void *packed, int ncol, bool is_complex)
{
    if (ncol <= 0) return 0;
    return is_complex ? repack_t<double2, true>(ctx, packed, const_cast<void *>(dense), ld_dense, ncol)
                      : repack_t<double, true>(ctx, packed, const_cast<void *>(dense), ld_dense, ncol);
}
int launch_unpack(chefsi_ctx *ctx, const void *packed, void *dense, size_t ld_dense, int ncol, bool is_complex)
{
    if (ncol <= 0) return 0;
    return is_complex ? repack_t<double2, false>(ctx, const_cast<void *>(packed), dense, ld_dense, ncol)
                      : repack_t<double, false>(ctx, const_cast<void *>(packed), dense, ld_dense, ncol);
}

int launch_halo_prepare(chefsi_ctx *ctx, void *buf, int ncol, bool is_complex, int zero_only)
{
    const Layout &L = ctx->lay;
    if (ncol <= 0 || (L.px == 0 && L.py == 0)) return 0;
    HaloDesc h;
    h.bcx = ctx->grid.BCx;
    h.bcy = ctx->grid.BCy;
    for (int oy = -1; oy <= 1; oy++)
        for (int ox = -1; ox <= 1; ox++) {
            const int q27 = 9 + (oy + 1) * 3 + (ox + 1); /* oz = 0 slice of the 27-entry table */
            h.ph_re[(oy + 1) * 3 + (ox + 1)] = ctx->desc.ph_re[q27];
            h.ph_im[(oy + 1) * 3 + (ox + 1)] = ctx->desc.ph_im[q27];
        }
    int launched = 0;
    for (int c0 = 0; c0 < ncol; c0 += 65535) {
        const int nc = (ncol - c0 < 65535) ? ncol - c0 : 65535;
        dim3 grid(ctx->num_sms, (unsigned)nc);
        if (is_complex)
            halo_prepare_kernel<double2><<<grid, 256, 0, ctx->stream>>>((double2 *)buf + (size_t)c0 * L.ld, L, h, zero_only);
        else
            halo_prepare_kernel<double><<<grid, 256, 0, ctx->stream>>>((double *)buf + (size_t)c0 * L.ld, L, h, zero_only);
        launched++;
    }
    return check(ctx, "halo_prepare", launched);
}
