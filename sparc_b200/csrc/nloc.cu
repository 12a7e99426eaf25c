/*
 * nloc.cu -- Kleinman-Bylander nonlocal projector apply (sm_100a), real and complex.
 *
 *   out += scale * Vnl x ,   Vnl x = sum_J conj(b_J) Chi_J Gamma ( sum_{J' of atom(J)} b_J' dV Chi_J'^T x[sphere_J'] )
 *
 * replacing Vnl_vec_mult (nlocVecRoutines.c:798-883) and Vnl_vec_mult_kpt (:889-999).  Both halves are
 * dense FP64 contractions (the reference calls dgemm/zgemm, :821,:872,:931,:988), so they run on the FP64
 * tensor cores (mma.sync m8n8k4 -> DMMA.8x8x4) with operand fragments loaded straight from global / L2:
 * the fragment shapes (8 rows x 4 consecutive elements) are exactly sector-sized runs of Chi and of the
 * gathered sphere points, so no shared-memory staging is needed.
 *
 *   project  alpha[atom][col][proj] = dV * sum_images phase * Chi^T x[sphere]        (:807-831 / :908-941)
 *            one CTA per (atom, 64 columns); a warp owns 8 real (4 complex) columns and all projectors:
 *            C(8 proj x 8 col) += A(8 proj x 4 pts) * B(4 pts x 8 col) per DMMA.  All periodic images of an
 *            atom are reduced by the same warp, so alpha needs no atomics and is deterministic.
 *   expand   out[sphere] += scale * conj(phase) * Chi (Gamma .* alpha)     (:841-881 / :951-997)
 *            one CTA per (image, 64 columns); C(8 pts x 8 col) = sum_k A(8 pts x 4 proj) * B(4 proj x 8 col)
 *            with the Gamma-scaled alpha fragments preloaded in registers; results are added into `out`
 *            with FP64 atomics only when the setup pass found overlapping spheres (small cells, Si8).
 *   patch    sphere points that are mirrored in the halo pads of the internal layout get their image
 *            refreshed (only with the streaming layout on periodic faces).
 *
 * Complex data (k-points) reuses the same kernels: Chi is real (nlocVecRoutines.c:731), so a complex
 * column is two real MMA columns (re, im); the per-image Bloch factor is applied lane-locally because one
 * lane holds the (re, im) pair of an accumulator element.
 */
#include "chefsi_internal.h"

namespace {

constexpr int kWarps = 8;
constexpr int kColsPerCta = 64; /* real columns (32 complex) */
constexpr int kMaxMT = 4;       /* projector tiles of 8  -> nproj <= 32 */
constexpr int kMaxKT = 8;       /* projector chunks of 4 */

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

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

/* x is addressed as real words: element (col, pos) of a complex block is at 2*(col*ld + pos) + {0,1}.
 * WORDS = 1 (real) or 2 (complex); the warp's 8 MMA columns are 8/WORDS data columns. */
template <int WORDS>
__global__ void __launch_bounds__(kWarps * 32)
nloc_project_kernel(const NlocView nl, const double *__restrict__ x, const size_t ld, const int ncol,
                    double *__restrict__ alpha, const double dV)
{
    const int atom = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lr = lane >> 2, lc = lane & 3;
    /* MMA column j (0..7) of this warp -> data column and word */
    const int mcol0 = blockIdx.y * kColsPerCta + warp * 8;        /* in MMA (real-word) columns */
    if (mcol0 / WORDS >= ncol) return;                            /* warp-uniform */
    const int bcol = mcol0 + lr;                                  /* B fragment: column index lr */
    const int bdata = bcol / WORDS, bword = bcol % WORDS;
    const bool bvalid = bdata < ncol;
    const double *__restrict__ xb = x + ((size_t)bdata * ld) * WORDS + bword;

    const int ip0 = nl.IP_displ[atom];
    const int nproj = nl.IP_displ[atom + 1] - ip0;
    const int MT = (nproj + 7) >> 3;
    const int j0 = nl.atom_img_off[atom], j1 = nl.atom_img_off[atom + 1];

    double tot[kMaxMT][2];
#pragma unroll
    for (int m = 0; m < kMaxMT; m++) tot[m][0] = tot[m][1] = 0.0;

    for (int jj = j0; jj < j1; jj++) {
        const int J = nl.atom_img[jj];
        const int ndc = nl.img_ndc[J];
        const int *__restrict__ pos = nl.grid_pos + nl.pos_off[J];
        const double *__restrict__ chi = nl.chi + nl.chi_off[J];
        double c[kMaxMT][2];
#pragma unroll
        for (int m = 0; m < kMaxMT; m++) c[m][0] = c[m][1] = 0.0;
#pragma unroll 2
        for (int i0 = 0; i0 < ndc; i0 += 4) {
            const int pt = i0 + lc;
            const bool pv = pt < ndc;
            double b = 0.0;
            if (pv && bvalid) b = xb[(size_t)pos[pt] * WORDS];
#pragma unroll
            for (int m = 0; m < kMaxMT; m++) {
                if (m < MT) {
                    const int pr = 8 * m + lr;
                    const double a = (pv && pr < nproj) ? chi[(size_t)pr * ndc + pt] : 0.0;
                    dmma(c[m][0], c[m][1], a, b);
                }
            }
        }
        if (WORDS == 2) { /* (c0, c1) = (re, im) of one complex accumulator: multiply by the Bloch factor */
            const double2 ph = nl.img_phase[J];
#pragma unroll
            for (int m = 0; m < kMaxMT; m++) {
                tot[m][0] += c[m][0] * ph.x - c[m][1] * ph.y;
                tot[m][1] += c[m][0] * ph.y + c[m][1] * ph.x;
            }
        } else {
#pragma unroll
            for (int m = 0; m < kMaxMT; m++) { tot[m][0] += c[m][0]; tot[m][1] += c[m][1]; }
        }
    }
    /* C fragment: row lr (projector), MMA columns 2*lc, 2*lc+1 */
#pragma unroll
    for (int m = 0; m < kMaxMT; m++) {
        const int pr = 8 * m + lr;
        if (m < MT && pr < nproj) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int mc = mcol0 + 2 * lc + e;
                const int dc = mc / WORDS, w = mc % WORDS;
                if (dc < ncol) alpha[(((size_t)ip0 * ncol + (size_t)dc * nproj + pr)) * WORDS + w] = tot[m][e] * dV;
            }
        }
    }
}

template <int WORDS>
__global__ void __launch_bounds__(kWarps * 32)
nloc_expand_kernel(const NlocView nl, const double *__restrict__ alpha, double *__restrict__ out, const size_t ld,
                   const int ncol, const double scale, const int use_atomics)
{
    const int J = blockIdx.x;
    const int atom = nl.img_atom[J];
    const int ip0 = nl.IP_displ[atom];
    const int nproj = nl.IP_displ[atom + 1] - ip0;
    if (nproj == 0) return;
    const int KT = (nproj + 3) >> 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lr = lane >> 2, lc = lane & 3;
    const int mcol0 = blockIdx.y * kColsPerCta + warp * 8;
    if (mcol0 / WORDS >= ncol) return;
    const int ndc = nl.img_ndc[J];

    /* B fragments: row lc (projector 4k+lc), column lr.  beta = scale * conj(phase) * Gamma * alpha */
    double bf[kMaxKT];
    {
        const int mc = mcol0 + lr;
        const int dc = mc / WORDS, w = mc % WORDS;
        double2 ph = make_double2(1.0, 0.0);
        if (WORDS == 2) ph = nl.img_phase[J];
#pragma unroll
        for (int k = 0; k < kMaxKT; k++) {
            const int pr = 4 * k + lc;
            double v = 0.0;
            if (k < KT && pr < nproj && dc < ncol) {
                const double g = nl.gamma[ip0 + pr] * scale;
                const size_t ai = ((size_t)ip0 * ncol + (size_t)dc * nproj + pr) * WORDS;
                if (WORDS == 2) {
                    const double ar = alpha[ai], aim = alpha[ai + 1];
                    /* (ar + i aim) * (cos - i sin), nlocVecRoutines.c:982 */
                    v = (w == 0) ? g * (ar * ph.x + aim * ph.y) : g * (aim * ph.x - ar * ph.y);
                } else {
                    v = g * alpha[ai];
                }
            }
            bf[k] = v;
        }
    }
    const int *__restrict__ pos = nl.grid_pos + nl.pos_off[J];
    const double *__restrict__ chi = nl.chi + nl.chi_off[J];
    /* this lane's two output MMA columns */
    const int oc0 = mcol0 + 2 * lc;
    const int od0 = oc0 / WORDS, ow0 = oc0 % WORDS;
    const int od1 = (oc0 + 1) / WORDS, ow1 = (oc0 + 1) % WORDS;
    double *__restrict__ o0 = out + ((size_t)od0 * ld) * WORDS + ow0;
    double *__restrict__ o1 = out + ((size_t)od1 * ld) * WORDS + ow1;
    const bool v0 = od0 < ncol, v1 = od1 < ncol;

    for (int i0 = 0; i0 < ndc; i0 += 8) {
        const int pt = i0 + lr;
        const bool pv = pt < ndc;
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int k = 0; k < kMaxKT; k++) {
            if (k < KT) {
                const int pr = 4 * k + lc;
                const double a = (pv && pr < nproj) ? chi[(size_t)pr * ndc + pt] : 0.0;
                dmma(c0, c1, a, bf[k]);
            }
        }
        if (pv) {
            const size_t g = (size_t)pos[pt] * WORDS;
            if (use_atomics) {
                if (v0) atomicAdd(o0 + g, c0);
                if (v1) atomicAdd(o1 + g, c1);
            } else {
                if (v0) o0[g] += c0;
                if (v1) o1[g] += c1;
            }
        }
    }
}

template <int WORDS>
__global__ void nloc_patch_kernel(double *__restrict__ out, const size_t ld, const int *__restrict__ src,
                                  const int *__restrict__ dst, const int n)
{
    double *col = out + (size_t)blockIdx.y * ld * WORDS;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
#pragma unroll
        for (int w = 0; w < WORDS; w++) col[(size_t)dst[t] * WORDS + w] = col[(size_t)src[t] * WORDS + w];
    }
}

template <int WORDS>
int launch_t(chefsi_ctx *ctx, const void *x, void *out, size_t ld, int ncol, double scale)
{
    NlocDev &d = ctx->nl;
    const size_t need = (size_t)d.ntot * ncol * sizeof(double) * WORDS;
    if (need > ctx->alpha_bytes) {
        if (ctx->d_alpha) cudaFree(ctx->d_alpha);
        ctx->d_alpha = nullptr;
        ctx->alpha_bytes = 0;
        cudaError_t e = cudaMalloc(&ctx->d_alpha, need);
        if (e != cudaSuccess) { chefsi_fail(ctx, "cudaMalloc(alpha, %zu): %s", need, cudaGetErrorString(e)); return -1; }
        ctx->alpha_bytes = need;
    }
    if (d.max_nproj > 8 * kMaxMT) { chefsi_fail(ctx, "nloc: more than %d projectors per atom not supported", 8 * kMaxMT); return -1; }
    NlocView v{d.IP_displ, d.gamma, d.img_atom, d.img_ndc, d.pos_off, d.chi_off,
               d.grid_pos, d.chi, d.img_phase, d.atom_img_off, d.atom_img};
    const int mma_cols = ncol * WORDS;
    const unsigned gy = (unsigned)((mma_cols + kColsPerCta - 1) / kColsPerCta);
    if (gy > 65535) { chefsi_fail(ctx, "nloc: too many columns per call"); return -1; }
    nloc_project_kernel<WORDS><<<dim3((unsigned)d.n_atom, gy), kWarps * 32, 0, ctx->stream>>>(
        v, reinterpret_cast<const double *>(x), ld, ncol, reinterpret_cast<double *>(ctx->d_alpha), ctx->grid.dV);
    nloc_expand_kernel<WORDS><<<dim3((unsigned)d.n_img, gy), kWarps * 32, 0, ctx->stream>>>(
        v, reinterpret_cast<const double *>(ctx->d_alpha), reinterpret_cast<double *>(out), ld, ncol, scale, d.overlap);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "nloc launch: %s", cudaGetErrorString(e)); return -1; }
    return 2;
}

}  // namespace

int launch_nloc_apply(chefsi_ctx *ctx, const void *x, void *out, size_t ld, int ncol, double scale,
                      bool is_complex)
{
    if (ctx->nl.n_img == 0 || ctx->nl.ntot == 0 || ncol <= 0) return 0;
    return is_complex ? launch_t<2>(ctx, x, out, ld, ncol, scale) : launch_t<1>(ctx, x, out, ld, ncol, scale);
}

int launch_nloc_halo_patch(chefsi_ctx *ctx, void *out, size_t ld, int ncol, bool is_complex)
{
    NlocDev &d = ctx->nl;
    if (d.n_patch == 0 || ncol <= 0) return 0;
    if (ncol > 65535) { chefsi_fail(ctx, "nloc patch: too many columns per call"); return -1; }
    dim3 grid((unsigned)((d.n_patch + 255) / 256), (unsigned)ncol);
    if (is_complex) nloc_patch_kernel<2><<<grid, 256, 0, ctx->stream>>>((double *)out, ld, d.patch_src, d.patch_dst, d.n_patch);
    else nloc_patch_kernel<1><<<grid, 256, 0, ctx->stream>>>((double *)out, ld, d.patch_src, d.patch_dst, d.n_patch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "nloc patch launch: %s", cudaGetErrorString(e)); return -1; }
    return 1;
}
