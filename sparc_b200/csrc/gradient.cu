/*
 * gradient.cu -- (D_dir + c) x for a block of columns (sm_100a): the first-derivative finite-difference stencil
 * along ONE lattice direction,
 *
 *     Dx[p] = c x[p] + sum_{r=1..FDn} w[r] (x[p + r e_dir] - x[p - r e_dir]),      w = D1_stencil_coeffs_{x,y,z}
 *
 * which is what Gradient_vectors_dir / Gradient_vec_dir (src/gradVecRoutines.c:32-311) and their k-point twins
 * (src/gradVecRoutinesKpt.c:35-340) compute: a copy of x extended by FDn points along dir only -- wrapped on a
 * periodic axis, zero on a Dirichlet axis (np = 1 branch, gradVecRoutines.c:262-300), times the Bloch phase
 * exp(-/+ i k L) on the low / high side for k-points (gradVecRoutinesKpt.c:179-191,301-311) -- then Calc_DX
 * (gradVecRoutines.c:318-409): temp = c x; temp += (x[+r] - x[-r]) w[r] for r = 1..FDn in that order, which is the
 * order of the fused multiply-adds below.
 *
 * Three kernels, all plain HBM streaming work (8 B read + 8 B written per grid point and column):
 *   march_kernel  (dir = y or z, FDn = 6): a thread owns one (x, other-axis) line of a column and walks along dir with
 *                 the 13 values of its stencil in registers, so every value is read once (+ 12 halo values per chunk);
 *                 lanes run along x, all loads and stores are coalesced;
 *   xline_kernel  (dir = x, FDn = 6): a warp owns one grid row at a time: the row and its 2 x 6 wrapped / zero halo
 *                 values go to the warp's shared-memory line once (coalesced), then every lane forms two neighbouring
 *                 outputs from one 14-value window (7 LDS.128) -- complex: one output from 13 LDS.128;
 *   gather_kernel (any other FDn, or rows too long for shared memory): a thread per point, the 2 FDn neighbours
 *                 come from the lines the neighbouring lanes load (L1 hits).
 */
#include "chefsi_internal.h"
#include "cplx.cuh"

namespace {

struct GradArgs {
    const void *x;
    void *out;
    size_t ld;      /* elements of T between columns (input and output blocks share the internal layout) */
    int ncol;
    int Nx, Ny, Nz;
    int dir, F, bc; /* bc: 1 = Dirichlet (zero halo) along dir */
    double c;
    double w[CHEFSI_MAXR + 1];
    double ph_re, ph_im; /* phase of the LOW-side halo, exp(-i k L); the high side takes the conjugate */
};

template <typename T>
__device__ __forceinline__ T load_wrapped(const T *__restrict__ line, int q, int N, size_t stride, const GradArgs &a)
{
    /* value of the extended line at position q in [-F, N + F) */
    if (q >= 0 && q < N) return line[(size_t)q * stride];
    if (a.bc) return cplx::zero<T>();
    if (q < 0) return cplx::mul_phase(line[(size_t)(q + N) * stride], a.ph_re, a.ph_im);
    return cplx::mul_phase(line[(size_t)(q - N) * stride], a.ph_re, -a.ph_im);
}

template <typename T>
__global__ void __launch_bounds__(256) gather_kernel(const __grid_constant__ GradArgs a)
{
    const size_t Nd = (size_t)a.Nx * a.Ny * a.Nz, total = Nd * a.ncol;
    const int N = a.dir == 0 ? a.Nx : (a.dir == 1 ? a.Ny : a.Nz);
    const size_t stride = a.dir == 0 ? 1 : (a.dir == 1 ? (size_t)a.Nx : (size_t)a.Nx * a.Ny);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t n = t / Nd, p = t - n * Nd;
        const int i = (int)(p % a.Nx), j = (int)((p / a.Nx) % a.Ny), k = (int)(p / ((size_t)a.Nx * a.Ny));
        const int q = a.dir == 0 ? i : (a.dir == 1 ? j : k);
        const T *__restrict__ line = reinterpret_cast<const T *>(a.x) + n * a.ld + (p - (size_t)q * stride);
        T acc = cplx::mul(line[(size_t)q * stride], a.c);
        for (int r = 1; r <= a.F; r++)
            acc = cplx::fma(cplx::sub(load_wrapped(line, q + r, N, stride, a), load_wrapped(line, q - r, N, stride, a)), a.w[r], acc);
        reinterpret_cast<T *>(a.out)[n * a.ld + p] = acc;
    }
}

/* dir = 1 (march along y, tile = x x z) or dir = 2 (march along z, tile = x x y); F = 6.  blockDim = (32, 8);
 * grid = (tiles of the two other axes, chunks along dir, columns). */
template <typename T, int DIR>
__global__ void __launch_bounds__(256, sizeof(T) == 8 ? 4 : 2) march_kernel(const __grid_constant__ GradArgs a, const int chunk,
                                                                            const int tiles_x)
{
    constexpr int F = 6, U = 4; /* U outputs per step: their U new values are loaded together (independent loads in flight) */
    const int No = DIR == 1 ? a.Nz : a.Ny;  /* the other tiled axis */
    const int N = DIR == 1 ? a.Ny : a.Nz;   /* the marching axis */
    const int i = (blockIdx.x % tiles_x) * 32 + threadIdx.x;
    const int o = (blockIdx.x / tiles_x) * 8 + threadIdx.y;
    if (i >= a.Nx || o >= No) return;
    const size_t plane = (size_t)a.Nx * a.Ny;
    const size_t stride = DIR == 1 ? (size_t)a.Nx : plane;
    const size_t base = DIR == 1 ? (size_t)o * plane + i : (size_t)o * a.Nx + i;
    const int q0 = blockIdx.y * chunk, q1 = min(N, q0 + chunk);
    for (int n = blockIdx.z; n < a.ncol; n += gridDim.z) {
        const T *__restrict__ line = reinterpret_cast<const T *>(a.x) + (size_t)n * a.ld + base;
        T *__restrict__ out = reinterpret_cast<T *>(a.out) + (size_t)n * a.ld + base;
        T v[2 * F + U]; /* v[s] = value at q - F + s */
#pragma unroll
        for (int s = 0; s < 2 * F; s++) v[s] = load_wrapped(line, q0 - F + s, N, stride, a);
        for (int q = q0; q < q1; q += U) {
#pragma unroll
            for (int u = 0; u < U; u++) v[2 * F + u] = (q + u < q1) ? load_wrapped(line, q + F + u, N, stride, a) : cplx::zero<T>();
#pragma unroll
            for (int u = 0; u < U; u++) {
                T acc = cplx::mul(v[F + u], a.c);
#pragma unroll
                for (int r = 1; r <= F; r++) acc = cplx::fma(cplx::sub(v[F + u + r], v[F + u - r]), a.w[r], acc);
                if (q + u < q1) out[(size_t)(q + u) * stride] = acc;
            }
#pragma unroll
            for (int s = 0; s < 2 * F; s++) v[s] = v[s + U];
        }
    }
}

/* dir = 0, F = 6.  blockDim = 256 (8 warps); dynamic shared memory: 8 lines of (Nx + 12) elements, padded to 16 bytes.
 * Rows (y, z, column) are dealt to the warps of the grid in a grid-stride loop. */
template <typename T>
__global__ void __launch_bounds__(256) xline_kernel(const __grid_constant__ GradArgs a, const int line_elems)
{
    constexpr int F = 6;
    extern __shared__ __align__(16) unsigned char xline_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T *row = reinterpret_cast<T *>(xline_smem) + (size_t)warp * line_elems;
    const int Nx = a.Nx;
    const size_t rows_per_col = (size_t)a.Ny * a.Nz, nrows = rows_per_col * a.ncol;
    for (size_t r = (size_t)blockIdx.x * 8 + warp; r < nrows; r += (size_t)gridDim.x * 8) {
        const size_t n = r / rows_per_col, off = (r - n * rows_per_col) * Nx;
        const T *__restrict__ x = reinterpret_cast<const T *>(a.x) + n * a.ld + off;
        T *__restrict__ out = reinterpret_cast<T *>(a.out) + n * a.ld + off;
        for (int i = lane; i < Nx; i += 32) row[F + i] = x[i];
        if (lane < 2 * F) { /* lanes 0..5: low-side halo, lanes 6..11: high-side halo */
            const int q = lane < F ? lane - F : Nx + lane - F;
            row[F + q] = load_wrapped(x, q, Nx, 1, a);
        }
        __syncwarp();
        if (cplx::is_complex<T>::value) {
            for (int i = lane; i < Nx; i += 32) {
                T acc = cplx::mul(row[F + i], a.c);
#pragma unroll
                for (int p = 1; p <= F; p++) acc = cplx::fma(cplx::sub(row[F + i + p], row[F + i - p]), a.w[p], acc);
                out[i] = acc;
            }
        } else {
            for (int i = 2 * lane; i < Nx; i += 64) {
                T v[2 * F + 2]; /* v[s] = value at i - F + s; row[i] is 16-byte aligned because i is even */
#pragma unroll
                for (int s = 0; s < 2 * F + 2; s += 2) {
                    const double2 t = *reinterpret_cast<const double2 *>(reinterpret_cast<const double *>(row) + i + s);
                    reinterpret_cast<double *>(v)[s] = t.x;
                    reinterpret_cast<double *>(v)[s + 1] = t.y;
                }
                T acc0 = cplx::mul(v[F], a.c), acc1 = cplx::mul(v[F + 1], a.c);
#pragma unroll
                for (int p = 1; p <= F; p++) {
                    acc0 = cplx::fma(cplx::sub(v[F + p], v[F - p]), a.w[p], acc0);
                    acc1 = cplx::fma(cplx::sub(v[F + 1 + p], v[F + 1 - p]), a.w[p], acc1);
                }
                out[i] = acc0;
                if (i + 1 < Nx) out[i + 1] = acc1;
            }
        }
        __syncwarp();
    }
}

template <typename T>
int launch_t(chefsi_ctx *ctx, const GradArgs &a)
{
    if (a.dir != 0 && a.F == 6) {
        const int No = a.dir == 1 ? a.Nz : a.Ny, N = a.dir == 1 ? a.Ny : a.Nz;
        const int tiles_x = (a.Nx + 31) / 32, tiles = tiles_x * ((No + 7) / 8);
        /* chunks along the marching axis: enough CTAs for four waves of 148 SMs x 8 CTAs, never shorter than 16 planes
           (the 12 halo values of a chunk are re-read) */
        const int cols = a.ncol < 65535 ? a.ncol : 65535;
        int nchunk = (int)((4L * 148 * 8 + (long)tiles * cols - 1) / ((long)tiles * cols));
        if (nchunk > (N + 15) / 16) nchunk = (N + 15) / 16;
        if (nchunk < 1) nchunk = 1;
        const int chunk = (N + nchunk - 1) / nchunk;
        const dim3 grid(tiles, (N + chunk - 1) / chunk, cols), block(32, 8);
        if (a.dir == 1) march_kernel<T, 1><<<grid, block, 0, ctx->stream>>>(a, chunk, tiles_x);
        else            march_kernel<T, 2><<<grid, block, 0, ctx->stream>>>(a, chunk, tiles_x);
    } else if (a.dir == 0 && a.F == 6 && (size_t)8 * ((a.Nx + 2 * 6 + 2) & ~1) * sizeof(T) <= (size_t)200 << 10) {
        const int line_elems = (a.Nx + 2 * 6 + 2) & ~1; /* an odd Nx reads one element past its halo: keep it inside the line */
        const size_t smem = (size_t)8 * line_elems * sizeof(T);
        if (smem > ((size_t)48 << 10)) /* per device, so not cached in a static: a multi-device context launches on each of its GPUs */
            CHEFSI_CUDA(ctx, cudaFuncSetAttribute(xline_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
        const size_t nrows = (size_t)a.Ny * a.Nz * a.ncol;
        size_t blocks = (nrows + 7) / 8;
        if (blocks > (size_t)148 * 32) blocks = (size_t)148 * 32;
        xline_kernel<T><<<(unsigned)blocks, 256, smem, ctx->stream>>>(a, line_elems);
    } else {
        const size_t total = (size_t)a.Nx * a.Ny * a.Nz * a.ncol;
        size_t blocks = (total + 255) / 256;
        if (blocks > (size_t)148 * 64) blocks = (size_t)148 * 64;
        gather_kernel<T><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a);
    }
    CHEFSI_CUDA(ctx, cudaGetLastError());
    return 0;
}

}  // namespace

/* out = (D_dir + c) x on ncol columns in the internal layout; kdir: the k-point component along dir (complex data).
 * Returns the number of kernel launches, or -1 after chefsi_fail. */
int launch_gradient(chefsi_ctx *ctx, const void *x, void *out, int ncol, int dir, double c, double kdir, bool is_complex)
{
    const chefsi_grid_t &g = ctx->grid;
    GradArgs a{};
    a.x = x;
    a.out = out;
    a.ld = ctx->ld;
    a.ncol = ncol;
    a.Nx = g.Nx; a.Ny = g.Ny; a.Nz = g.Nz;
    a.dir = dir;
    a.F = g.FDn;
    a.bc = dir == 0 ? g.BCx : (dir == 1 ? g.BCy : g.BCz);
    a.c = c;
    const double *w = dir == 0 ? g.D1_x : (dir == 1 ? g.D1_y : g.D1_z);
    for (int r = 0; r <= g.FDn; r++) a.w[r] = w[r];
    const double L = dir == 0 ? g.range_x : (dir == 1 ? g.range_y : g.range_z);
    a.ph_re = cos(kdir * L);   /* phase_fac_l = cos(k L) - i sin(k L), gradVecRoutinesKpt.c:179-181 */
    a.ph_im = -sin(kdir * L);
    const int rc = is_complex ? launch_t<double2>(ctx, a) : launch_t<double>(ctx, a);
    return rc == 0 ? 1 : -1;
}
