/*
 * stencil_general.cu -- general fused Chebyshev/Hamiltonian step kernel (sm_100a).
 *
 *   out = s1 * ( (-1/2 Lap + Veff + c) x ) - s2 * xprev
 *
 * for every cell type SPARC's filter supports (cell_typ 0 and 11..17), periodic or
 * Dirichlet faces, real (Gamma) or complex (k-point, Bloch-phase halos) data.  One CTA owns
 * a TX x TY x TZ brick of one orbital column:
 *
 *   1. the brick plus an FDn-wide halo is staged in shared memory; halo points that leave the
 *      cell take the periodically wrapped value (times exp(i k.L) in the complex case) or zero
 *      on Dirichlet faces -- the rule of the reference's x_ex copy at np = 1
 *      (lapVecRoutines.c:536-577,1183-1204; lapVecRoutinesKpt.c:370-462,647-865);
 *   2. for non-orthogonal cells the intermediate first-derivative fields of the reference's
 *      two-stage mixed-derivative composition (Calc_DX gradVecRoutines.c:318, Calc_DX1_DX2
 *      lapVecRoutines.c:1429) are formed in shared memory on the box extended along one axis;
 *   3. every thread evaluates the star stencil (stencil_3axis_thread_v2 lapVecRoutines.c:257,
 *      stencil_4comp :1481, stencil_5comp :1540) for its points, adds (Veff + c) x and applies the
 *      three-term recurrence scaling (eigenSolver.c:763-768,787-794) before the single store.
 *
 * This is the catch-all path (small grids, odd sizes, non-orthogonal and k-point runs).  The
 * large orthogonal workload goes through stencil_stream_dense.cu instead.
 */
#include "chefsi_internal.h"
#include "cplx.cuh"

namespace {

constexpr int kThreads = 256;

template <typename T, int TX, int TY, int TZ, int RT>
__global__ void __launch_bounds__(kThreads)
stencil_general_kernel(const __grid_constant__ StencilDesc d, const StepArgs a, const int nbx,
                       const int nby)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int F = RT ? RT : d.F;
    const int EX = TX + 2 * F, EY = TY + 2 * F, EZ = TZ + 2 * F;
    const int EXY = EX * EY;
    T *f = reinterpret_cast<T *>(smem_raw);

    const int tid = threadIdx.x;
    const int bid = blockIdx.x;
    const int bx = bid % nbx, by = (bid / nbx) % nby, bz = bid / (nbx * nby);
    const int x0 = bx * TX, y0 = by * TY, z0 = bz * TZ;
    const size_t col = (size_t)blockIdx.y * a.ld;
    const T *__restrict__ x = reinterpret_cast<const T *>(a.x) + col;
    const int Nx = d.Nx, Ny = d.Ny, Nz = d.Nz;

    /* ---- 1. stage brick + halo ---------------------------------------------------------- */
    for (int idx = tid; idx < EXY * EZ; idx += kThreads) {
        const int ip = idx % EX, jp = (idx / EX) % EY, kp = idx / EXY;
        int i = x0 - F + ip, j = y0 - F + jp, k = z0 - F + kp;
        int ox = 0, oy = 0, oz = 0;
        if (i < 0) { i += Nx; ox = -1; } else if (i >= Nx) { i -= Nx; ox = 1; }
        if (j < 0) { j += Ny; oy = -1; } else if (j >= Ny) { j -= Ny; oy = 1; }
        if (k < 0) { k += Nz; oz = -1; } else if (k >= Nz) { k -= Nz; oz = 1; }
        const bool dead = (ox && d.bc[0]) || (oy && d.bc[1]) || (oz && d.bc[2]) ||
                          i < 0 || i >= Nx || j < 0 || j >= Ny || k < 0 || k >= Nz;
        T v = cplx::zero<T>();
        if (!dead) {
            v = x[lay_pos(d.lay, i, j, k)];
            if (cplx::is_complex<T>::value && (ox | oy | oz)) {
                const int q = (oz + 1) * 9 + (oy + 1) * 3 + (ox + 1);
                v = cplx::mul_phase(v, d.ph_re[q], d.ph_im[q]);
            }
        }
        f[idx] = v;
    }
    __syncthreads();

    /* ---- 2. intermediate derivative fields for the mixed terms --------------------------- */
    T *D[2] = {f + EXY * EZ, nullptr};
    int dR0[2] = {0, 0}, dR01[2] = {0, 0}, dsd[2] = {0, 0}, doff[2][3] = {{0, 0, 0}, {0, 0, 0}};
    const int hs[3] = {1, EX, EXY};
    for (int q = 0; q < d.nmix; q++) {
        const MixedComp &mc = d.mix[q];
        int R[3] = {TX, TY, TZ};
        R[mc.ext] += 2 * F;
        int O[3] = {F, F, F};
        O[mc.ext] = 0;
        const int n = R[0] * R[1] * R[2];
        if (q == 1) D[1] = D[0] + dR01[0] * ((d.mix[0].ext == 2) ? TZ + 2 * F : TZ);
        dR0[q] = R[0];
        dR01[q] = R[0] * R[1];
        dsd[q] = (mc.ext == 0) ? 1 : (mc.ext == 1 ? R[0] : R[0] * R[1]);
        doff[q][0] = (mc.ext == 0) ? F : 0;
        doff[q][1] = (mc.ext == 1) ? F : 0;
        doff[q][2] = (mc.ext == 2) ? F : 0;
        const int s1 = hs[mc.ax1];
        const int s2 = (mc.ax2 >= 0) ? hs[mc.ax2] : 0;
        for (int idx = tid; idx < n; idx += kThreads) {
            const int a0 = idx % R[0], a1 = (idx / R[0]) % R[1], a2 = idx / (R[0] * R[1]);
            const int p = (a2 + O[2]) * EXY + (a1 + O[1]) * EX + (a0 + O[0]);
            T t1 = cplx::zero<T>(), t2 = cplx::zero<T>();
#pragma unroll
            for (int r = 1; r <= (RT ? RT : CHEFSI_MAXR); r++) {
                if (!RT && r > F) break;
                t1 = cplx::fma(cplx::sub(f[p + r * s1], f[p - r * s1]), mc.c1[r], t1);
                if (mc.ax2 >= 0) t2 = cplx::fma(cplx::sub(f[p + r * s2], f[p - r * s2]), mc.c2[r], t2);
            }
            D[q][idx] = (mc.ax2 >= 0) ? cplx::add(t1, t2) : t1;
        }
    }
    if (d.nmix) __syncthreads();

    /* ---- 3. star stencil + potential + recurrence ---------------------------------------- */
    const T *__restrict__ xprev = reinterpret_cast<const T *>(a.xprev);
    T *__restrict__ out = reinterpret_cast<T *>(a.out) + col;
    const double diag0 = d.coef0 + a.c;
    for (int l = tid; l < TX * TY * TZ; l += kThreads) {
        const int lx = l % TX, ly = (l / TX) % TY, lz = l / (TX * TY);
        const int i = x0 + lx, j = y0 + ly, k = z0 + lz;
        if (i >= Nx || j >= Ny || k >= Nz) continue;
        const int p = (lz + F) * EXY + (ly + F) * EX + (lx + F);
        const size_t g = lay_pos(d.lay, i, j, k);
        const T xc = f[p];
        T res = cplx::mul(xc, diag0);
#pragma unroll
        for (int r = 1; r <= (RT ? RT : CHEFSI_MAXR); r++) {
            if (!RT && r > F) break;
            T acc = cplx::mul(cplx::add(f[p - r], f[p + r]), d.wx[r]);
            acc = cplx::fma(cplx::add(f[p - r * EX], f[p + r * EX]), d.wy[r], acc);
            acc = cplx::fma(cplx::add(f[p - r * EXY], f[p + r * EXY]), d.wz[r], acc);
            for (int q = 0; q < d.nmix; q++) {
                const int pd = (lz + doff[q][2]) * dR01[q] + (ly + doff[q][1]) * dR0[q] + lx + doff[q][0];
                acc = cplx::fma(cplx::sub(D[q][pd + r * dsd[q]], D[q][pd - r * dsd[q]]), d.mix[q].wm[r], acc);
            }
            res = cplx::add(res, acc);
        }
        if (a.veff) res = cplx::fma(xc, a.veff[g], res);
        T o = cplx::mul(res, a.s1);
        if (a.s2 != 0.0) o = cplx::fma(xprev[col + g], -a.s2, o);
        out[g] = o;
    }
}

template <typename T, int TX, int TY, int TZ>
size_t smem_bytes(const StencilDesc &d)
{
    const int F = d.F;
    size_t n = (size_t)(TX + 2 * F) * (TY + 2 * F) * (TZ + 2 * F);
    for (int q = 0; q < d.nmix; q++) {
        int R[3] = {TX, TY, TZ};
        R[d.mix[q].ext] += 2 * F;
        n += (size_t)R[0] * R[1] * R[2];
    }
    return n * sizeof(T);
}

template <typename T, int TX, int TY, int TZ>
int launch_t(chefsi_ctx *ctx, const StepArgs &a)
{
    const StencilDesc &d = ctx->desc;
    const int nbx = (d.Nx + TX - 1) / TX, nby = (d.Ny + TY - 1) / TY, nbz = (d.Nz + TZ - 1) / TZ;
    const size_t smem = smem_bytes<T, TX, TY, TZ>(d);
    if (smem > ctx->max_smem_optin)
        return chefsi_fail(ctx, "general stencil: FD radius %d needs %zu B of shared memory (> %zu)", d.F, smem,
                           ctx->max_smem_optin) ? -1 : -1;
    dim3 grid((unsigned)(nbx * nby * nbz), (unsigned)a.ncol);
    auto k6 = stencil_general_kernel<T, TX, TY, TZ, 6>;
    auto k0 = stencil_general_kernel<T, TX, TY, TZ, 0>;
    auto kern = (d.F == 6) ? k6 : k0;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { chefsi_fail(ctx, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
    /* ncol may exceed the 65535 limit of gridDim.y: launch in slabs of columns */
    int launched = 0;
    for (int c0 = 0; c0 < a.ncol; c0 += 65535) {
        StepArgs b = a;
        const int nc = (a.ncol - c0 < 65535) ? a.ncol - c0 : 65535;
        const size_t off = (size_t)c0 * a.ld * sizeof(T);
        b.x = (const char *)a.x + off;
        b.out = (char *)a.out + off;
        if (a.xprev) b.xprev = (const char *)a.xprev + off;
        b.ncol = nc;
        grid.y = (unsigned)nc;
        kern<<<grid, kThreads, smem, ctx->stream>>>(d, b, nbx, nby);
        launched++;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "general stencil launch: %s", cudaGetErrorString(e)); return -1; }
    return launched;
}

}  // namespace

int launch_stencil_general(chefsi_ctx *ctx, const StepArgs &a, bool is_complex)
{
    if (a.ncol <= 0) return 0;
    if (is_complex) return launch_t<double2, 16, 8, 4>(ctx, a);
    return launch_t<double, 16, 8, 8>(ctx, a);
}
