/*
 * nloc.cu -- Kleinman-Bylander nonlocal projector apply (sm_100a), real and complex.
 *
 *   out += scale * Vnl x ,   Vnl x = sum_J conj(b_J) Chi_J Gamma ( sum_{J' of atom(J)} b_J' dV Chi_J'^T x[sphere_J'] )
 *
 * replacing Vnl_vec_mult (nlocVecRoutines.c:798-883) and Vnl_vec_mult_kpt (:889-999):
 *
 *   project  (gather + contraction, :807-831 / :908-941)
 *            one CTA per (atom, 8 columns); a warp owns one column, its lanes stride over the
 *            sphere's grid points (coalesced Chi reads, gathered x reads), partial inner products
 *            for 8 projectors at a time live in registers and are combined with warp shuffles.
 *            All periodic images of an atom are handled by the same CTA, so alpha needs no
 *            atomics and the result is deterministic (the reference accumulates images with
 *            beta = 1 in dgemm, :821-827).
 *   expand   (Gamma scale :841-863, contraction + scatter-add :866-881 / :968-997)
 *            one CTA per (image, 8 columns); Gamma*alpha for the 8 columns sits in shared memory,
 *            each thread owns sphere points and adds into `out` -- with FP64 atomics only when
 *            the setup pass found overlapping spheres (small cells such as Si8).
 *
 * alpha layout is the reference's: [atom][column][projector].
 */
#include "chefsi_internal.h"
#include "cplx.cuh"

namespace {

constexpr int kProjWarps = 8;  /* columns per project-CTA */
constexpr int kProjChunk = 8;  /* projectors accumulated per pass */
constexpr int kExpCols = 8;    /* columns per expand-CTA */
constexpr int kExpThreads = 128;

struct NlocView {
    const int *IP_displ;
    const double *gamma;
    const int *img_atom, *img_ndc;
    const long long *pos_off, *chi_off;
    const int *grid_pos;
    const double *chi;
    const double2 *img_phase;
    const int *atom_img_off, *atom_img;
};

template <typename T> __device__ __forceinline__ T warp_sum(T v);
template <> __device__ __forceinline__ double warp_sum<double>(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <> __device__ __forceinline__ double2 warp_sum<double2>(double2 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}

template <typename T>
__global__ void __launch_bounds__(kProjWarps * 32)
nloc_project_kernel(const NlocView nl, const T *__restrict__ x, const size_t ld, const int ncol,
                    T *__restrict__ alpha, const double dV)
{
    const int atom = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.y * kProjWarps + warp;
    if (n >= ncol) return;
    const int ip0 = nl.IP_displ[atom];
    const int nproj = nl.IP_displ[atom + 1] - ip0;
    const int j0 = nl.atom_img_off[atom], j1 = nl.atom_img_off[atom + 1];
    const T *__restrict__ xc = x + (size_t)n * ld;
    T *__restrict__ ablk = alpha + (size_t)ip0 * ncol + (size_t)n * nproj;

    for (int pc = 0; pc < nproj; pc += kProjChunk) {
        T tot[kProjChunk];
#pragma unroll
        for (int q = 0; q < kProjChunk; q++) tot[q] = cplx::zero<T>();
        for (int jj = j0; jj < j1; jj++) {
            const int J = nl.atom_img[jj];
            const int ndc = nl.img_ndc[J];
            const int *__restrict__ pos = nl.grid_pos + nl.pos_off[J];
            const double *__restrict__ chi = nl.chi + nl.chi_off[J] + (size_t)pc * ndc;
            T acc[kProjChunk];
#pragma unroll
            for (int q = 0; q < kProjChunk; q++) acc[q] = cplx::zero<T>();
            for (int i = lane; i < ndc; i += 32) {
                const T xv = xc[pos[i]];
#pragma unroll
                for (int q = 0; q < kProjChunk; q++)
                    if (pc + q < nproj) acc[q] = cplx::fma(xv, chi[(size_t)q * ndc + i], acc[q]);
            }
            if (cplx::is_complex<T>::value) {
                const double2 ph = nl.img_phase[J];
#pragma unroll
                for (int q = 0; q < kProjChunk; q++) tot[q] = cplx::add(tot[q], cplx::mul_phase(acc[q], ph.x, ph.y));
            } else {
#pragma unroll
                for (int q = 0; q < kProjChunk; q++) tot[q] = cplx::add(tot[q], acc[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < kProjChunk; q++) {
            const T s = warp_sum<T>(tot[q]);
            if (lane == 0 && pc + q < nproj) ablk[pc + q] = cplx::mul(s, dV);
        }
    }
}

__device__ __forceinline__ void accumulate(double *p, double v, bool atomic)
{
    if (atomic) atomicAdd(p, v); else *p += v;
}
__device__ __forceinline__ void accumulate(double2 *p, double2 v, bool atomic)
{
    if (atomic) {
        atomicAdd(&p->x, v.x);
        atomicAdd(&p->y, v.y);
    } else {
        double2 o = *p;
        o.x += v.x;
        o.y += v.y;
        *p = o;
    }
}

template <typename T>
__global__ void __launch_bounds__(kExpThreads)
nloc_expand_kernel(const NlocView nl, const T *__restrict__ alpha, T *__restrict__ out, const size_t ld,
                   const int ncol, const double scale, const int use_atomics)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *beta = reinterpret_cast<T *>(smem_raw); /* [kExpCols][nproj] */
    const int J = blockIdx.x;
    const int atom = nl.img_atom[J];
    const int ip0 = nl.IP_displ[atom];
    const int nproj = nl.IP_displ[atom + 1] - ip0;
    if (nproj == 0) return;
    const int n0 = blockIdx.y * kExpCols;
    const int nc = min(kExpCols, ncol - n0);
    const int ndc = nl.img_ndc[J];

    for (int t = threadIdx.x; t < kExpCols * nproj; t += kExpThreads) {
        const int c = t / nproj, p = t - c * nproj;
        T b = cplx::zero<T>();
        if (c < nc) {
            b = cplx::mul(alpha[(size_t)ip0 * ncol + (size_t)(n0 + c) * nproj + p], nl.gamma[ip0 + p] * scale);
            if (cplx::is_complex<T>::value) {
                const double2 ph = nl.img_phase[J];
                b = cplx::mul_phase(b, ph.x, -ph.y); /* cos(theta) - i sin(theta), nlocVecRoutines.c:982 */
            }
        }
        beta[t] = b;
    }
    __syncthreads();

    const int *__restrict__ pos = nl.grid_pos + nl.pos_off[J];
    const double *__restrict__ chi = nl.chi + nl.chi_off[J];
    for (int i = threadIdx.x; i < ndc; i += kExpThreads) {
        T v[kExpCols];
#pragma unroll
        for (int c = 0; c < kExpCols; c++) v[c] = cplx::zero<T>();
        for (int p = 0; p < nproj; p++) {
            const double ch = chi[(size_t)p * ndc + i];
#pragma unroll
            for (int c = 0; c < kExpCols; c++) v[c] = cplx::fma(beta[c * nproj + p], ch, v[c]);
        }
        const size_t g = (size_t)pos[i];
#pragma unroll
        for (int c = 0; c < kExpCols; c++)
            if (c < nc) accumulate(out + (size_t)(n0 + c) * ld + g, v[c], use_atomics != 0);
    }
}

template <typename T>
int launch_t(chefsi_ctx *ctx, const void *x, void *out, size_t ld, int ncol, double scale)
{
    NlocDev &d = ctx->nl;
    const size_t need = (size_t)d.ntot * ncol * sizeof(T);
    if (need > ctx->alpha_bytes) {
        if (ctx->d_alpha) cudaFree(ctx->d_alpha);
        ctx->d_alpha = nullptr;
        ctx->alpha_bytes = 0;
        cudaError_t e = cudaMalloc(&ctx->d_alpha, need);
        if (e != cudaSuccess) { chefsi_fail(ctx, "cudaMalloc(alpha, %zu): %s", need, cudaGetErrorString(e)); return -1; }
        ctx->alpha_bytes = need;
    }
    NlocView v{d.IP_displ, d.gamma, d.img_atom, d.img_ndc, d.pos_off, d.chi_off,
               d.grid_pos, d.chi, d.img_phase, d.atom_img_off, d.atom_img};
    int launched = 0;
    /* gridDim.y is limited to 65535: slab the columns (never hit in practice) */
    const int slab = 65535 * kExpCols;
    for (int c0 = 0; c0 < ncol; c0 += slab) {
        const int nc = (ncol - c0 < slab) ? ncol - c0 : slab;
        if (c0 != 0) { chefsi_fail(ctx, "nloc: more than %d columns per call not supported", slab); return -1; }
        dim3 g1((unsigned)d.n_atom, (unsigned)((nc + kProjWarps - 1) / kProjWarps));
        nloc_project_kernel<T><<<g1, kProjWarps * 32, 0, ctx->stream>>>(
            v, reinterpret_cast<const T *>(x), ld, nc, reinterpret_cast<T *>(ctx->d_alpha), ctx->grid.dV);
        dim3 g2((unsigned)d.n_img, (unsigned)((nc + kExpCols - 1) / kExpCols));
        const size_t smem = (size_t)kExpCols * d.max_nproj * sizeof(T);
        nloc_expand_kernel<T><<<g2, kExpThreads, smem, ctx->stream>>>(
            v, reinterpret_cast<const T *>(ctx->d_alpha), reinterpret_cast<T *>(out), ld, nc, scale, d.overlap);
        launched += 2;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "nloc launch: %s", cudaGetErrorString(e)); return -1; }
    return launched;
}

}  // namespace

int launch_nloc_apply(chefsi_ctx *ctx, const void *x, void *out, size_t ld, int ncol, double scale,
                      bool is_complex)
{
    if (ctx->nl.n_img == 0 || ctx->nl.ntot == 0 || ncol <= 0) return 0;
    return is_complex ? launch_t<double2>(ctx, x, out, ld, ncol, scale)
                      : launch_t<double>(ctx, x, out, ld, ncol, scale);
}
