/*
 * nloc.cu -- Kleinman-Bylander nonlocal projector apply (sm_100a), real and complex.
 *
 *   out += scale * Vnl x ,   Vnl x = sum_J conj(b_J) Chi_J Gamma ( sum_{J' of atom(J)} b_J' dV Chi_J'^T x[sphere_J'] )
 *
 * replacing Vnl_vec_mult (nlocVecRoutines.c:798-883) and Vnl_vec_mult_kpt (:889-999).
 *
 * The first ncu capture of round 1 (profiles/r1_ncu_stream_nloc_summary.md) showed that the projector
 * step is NOT compute-limited (FP64 pipe < 1 % active with the DMMA kernels of the first draft) but
 * latency- and sector-limited by the sphere gather/scatter, and that the separate project and expand
 * passes touched the sphere points three times per Chebyshev step.  So the contraction runs on the
 * plain FP64 FMA pipe and the kernel is organised around the memory access instead:
 *
 *   one kernel, one CTA per (sphere image J, 32 real columns), marching the sphere in chunks of 64 points:
 *     A  gather   lanes <-> sphere points (runs along x are contiguous), cp.async 8 B elements into a
 *                 double-buffered shared tile V[column][point]; the Chi chunk (stored point-major,
 *                 zero-padded to NP projectors at set_projectors time) arrives with 16 B cp.async;
 *     B  contract lanes <-> columns, warps <-> points: a lane keeps beta[NP] = scale*Gamma*alpha_prev and
 *                 acc[NP] in registers; per point one conflict-free LDS of V, NP/2 broadcast LDS.128 of
 *                 Chi, NP FMAs "expand" (v += Chi beta), NP FMAs "project" (acc += Chi v);
 *     C  scatter  lanes <-> points again, the updated tile goes back with coalesced stores.
 *   Modes: PROJECT (alpha of the input only), FUSED (expand with alpha_prev into `out`, then project the
 *   now final `out` for the NEXT Chebyshev step: one read + one write of the sphere points per step),
 *   EXPAND (last step / single H apply), EXPAND_ATOMIC (spheres overlap: FP64 atomics, no read).
 *
 * alpha is kept per IMAGE (a partial sum over that image's points, already times dV and the Bloch
 * factor); the consumer adds the partials of all images of the atom in a fixed order, so the result is
 * deterministic and needs no atomics or zero fill (the reference accumulates images with beta = 1,
 * nlocVecRoutines.c:821-827).
 *
 * Complex data (k-points): Chi is real (nlocVecRoutines.c:731), so a complex column is two real "word"
 * columns (re, im); the Bloch factors (nlocVecRoutines.c:911-921, :982-989) are applied where beta is
 * formed and where alpha is written.
 */
#include <algorithm>

#include "chefsi_internal.h"

namespace {

constexpr int kCols = 32;            /* real word columns per CTA (lanes) */

/* Pipeline shape: T threads per CTA, S-deep ring of KP-point chunks.  The gather is latency-bound
 * (ncu: long_scoreboard + barrier stalls dominate, FP64 pipe ~20 % busy), so what matters is how many
 * gather bytes an SM keeps in flight: CTAs per SM x (S-1) chunks x KP x 32 columns x 8 B. */
template <int T_, int S_, int KP_> struct Shape {
    static constexpr int T = T_, S = S_, KP = KP_;
    static constexpr int kWarps = T / 32;
    static constexpr int kPitch = KP + 1;       /* odd pitch: conflict-free column-strided LDS */
    static constexpr int kPtsPerWarp = KP / kWarps;
    static constexpr int H = KP / 32;           /* points per lane in the gather / scatter phases */
    static constexpr int kColsPerWarp = kCols / kWarps;
    static_assert(KP % 32 == 0 && KP % kWarps == 0 && kCols % kWarps == 0, "shape");
};

enum { MODE_PROJECT = 0, MODE_FUSED = 1, MODE_EXPAND = 2, MODE_EXPAND_ATOMIC = 3 };

struct NlocView {
    const int *IP_displ;
    const double *gamma;
    const int *img_atom, *img_ndc, *img_aoff;
    const long long *pos_off, *chiT_off;
    const int *grid_pos;
    const double *chiT;
    const double2 *img_phase;
    const int *atom_img_off, *atom_img;
    const double *alpha_sum; /* != NULL: per-atom sums of the partials (alpha_reduce_kernel), atom a at IP_displ[a] * ncol * WORDS */
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async8(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NP, class SH> struct Smem {
    double V[SH::S][kCols][SH::kPitch];
    double chi[SH::S][SH::KP][NP];
    int pos[SH::S][SH::KP];
};

/* vec: MODE_PROJECT -> the input block x (read only); otherwise the output block (read-modify-write).
 * Word column wc of the block lives at vec + (wc / WORDS) * ld * WORDS + (wc % WORDS), element stride WORDS. */
template <int NP, int WORDS, int MODE, class SH, bool SUMMED>
__global__ void __launch_bounds__(SH::T, ((NP <= 20) ? 2 : 1) * (256 / SH::T))
nloc_kernel(const NlocView nl, const double *__restrict__ alpha_prev, double *__restrict__ alpha_next,
            double *__restrict__ vec, const size_t ld, const int ncol, const int ngroups, const double scale,
            const double dV)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kThreads = SH::T, kWarps = SH::kWarps, kP = SH::KP, kPtsPerWarp = SH::kPtsPerWarp, NS = SH::S;
    constexpr int H = SH::H, kCPW = SH::kColsPerWarp;
    Smem<NP, SH> &S = *reinterpret_cast<Smem<NP, SH> *>(smem_raw);

    const int J = blockIdx.x / ngroups, grp = blockIdx.x % ngroups;
    const int atom = nl.img_atom[J];
    const int ip0 = nl.IP_displ[atom];
    const int nproj = nl.IP_displ[atom + 1] - ip0;
    const int ndc = nl.img_ndc[J];
    if (nproj == 0 || ndc == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwc = ncol * WORDS;                /* word columns in the block */
    const int wc0 = grp * kCols;                 /* first word column of this CTA */
    const int *__restrict__ pos = nl.grid_pos + nl.pos_off[J];
    const double *__restrict__ chiT = nl.chiT + nl.chiT_off[J];
    const int nchunks = (ndc + kP - 1) / kP;

    /* ---- beta = scale * conj(phase_J) * Gamma * sum_{images of the atom} alpha_prev ------------------ */
    double beta[NP], acc[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) { beta[p] = 0.0; acc[p] = 0.0; }
    const int mywc = wc0 + lane;
    const bool colvalid = mywc < nwc;
    /* alpha partials are stored [image/segment][data column][projector] */
    if (MODE != MODE_PROJECT && colvalid) {
        const int dc = mywc / WORDS, w = mywc % WORDS;
        double sre[NP], sim[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) { sre[p] = 0.0; sim[p] = 0.0; }
        if (SUMMED) { /* per-atom sums prepared by alpha_reduce_kernel (same [column][projector] layout per atom) */
            const double *ap = nl.alpha_sum + ((size_t)ip0 * ncol + (size_t)dc * nproj) * WORDS;
#pragma unroll
            for (int p = 0; p < NP; p++)
                if (p < nproj) {
                    sre[p] = ap[p * WORDS];
                    if (WORDS == 2) sim[p] = ap[p * WORDS + 1];
                }
        } else {
            for (int jj = nl.atom_img_off[atom]; jj < nl.atom_img_off[atom + 1]; jj++) {
                const int J2 = nl.atom_img[jj];
                const double *ap = alpha_prev + ((size_t)nl.img_aoff[J2] * ncol + (size_t)dc * nproj) * WORDS;
#pragma unroll
                for (int p = 0; p < NP; p++)
                    if (p < nproj) {
                        sre[p] += ap[p * WORDS];
                        if (WORDS == 2) sim[p] += ap[p * WORDS + 1];
                    }
            }
        }
        double2 ph = make_double2(1.0, 0.0);
        if (WORDS == 2) ph = nl.img_phase[J];
#pragma unroll
        for (int p = 0; p < NP; p++)
            if (p < nproj) {
                const double g = nl.gamma[ip0 + p] * scale;
                /* (ar + i ai) * (cos - i sin), nlocVecRoutines.c:982 */
                if (WORDS == 2) beta[p] = (w == 0) ? g * (sre[p] * ph.x + sim[p] * ph.y) : g * (sim[p] * ph.x - sre[p] * ph.y);
                else beta[p] = g * sre[p];
            }
    }

    /* ---- chunk pipeline ------------------------------------------------------------------------------ */
    /* phase A/C mapping: lane <-> points (lane, lane+32, ..), warp <-> kCPW word columns.  The grid
       positions of the chunk after the one being issued are prefetched into registers (psn) so the
       cp.async addresses never wait on a dependent global load. */
    int psn[H];
    auto load_pos = [&](int k) {
#pragma unroll
        for (int h = 0; h < H; h++) {
            const int pt = k * kP + lane + 32 * h;
            psn[h] = (k < nchunks && pt < ndc) ? __ldg(pos + pt) : -1;
        }
    };
    auto issue_chunk = [&](int k, int buf) {
        const int pt0 = k * kP;
        int ps[H];
#pragma unroll
        for (int h = 0; h < H; h++) ps[h] = psn[h];
        load_pos(k + 1);
        if (warp == 0) {
#pragma unroll
            for (int h = 0; h < H; h++) S.pos[buf][lane + 32 * h] = ps[h];
        }
        if (MODE != MODE_EXPAND_ATOMIC) {
#pragma unroll
            for (int c = 0; c < kCPW; c++) {
                const int cl = warp * kCPW + c;
                const int wc = wc0 + cl;
                const double *base = vec + ((size_t)(wc / WORDS) * ld) * WORDS + (wc % WORDS);
#pragma unroll
                for (int h = 0; h < H; h++) {
                    double *dst = &S.V[buf][cl][lane + 32 * h];
                    if (ps[h] >= 0 && wc < nwc) cp_async8(dst, base + (size_t)ps[h] * WORDS);
                    else *dst = 0.0;
                }
            }
        }
        /* Chi chunk: kP x NP doubles, contiguous in the point-major table */
        const int npts = (ndc - pt0 < kP) ? ndc - pt0 : kP;
        const double *src = chiT + (size_t)pt0 * NP;
        double *dstc = &S.chi[buf][0][0];
        for (int t = threadIdx.x; t < kP * NP / 2; t += kThreads) {
            if (2 * t < npts * NP) cp_async16(dstc + 2 * t, src + 2 * t);
            else { dstc[2 * t] = 0.0; dstc[2 * t + 1] = 0.0; }
        }
    };

    load_pos(0);
#pragma unroll
    for (int s = 0; s < NS - 1; s++) {
        if (s < nchunks) issue_chunk(s, s);
        cp_async_commit();
    }
    for (int k = 0; k < nchunks; k++) {
        const int buf = k % NS;
        cp_async_wait<NS - 2>();
        __syncthreads(); /* chunk k has landed; everybody is done with phase C of chunk k-1 (its buffer is refilled now) */
        if (k + NS - 1 < nchunks) issue_chunk(k + NS - 1, (k + NS - 1) % NS); /* in flight during phases B and C */
        cp_async_commit();

        /* ---- B: lanes <-> columns, warp handles points warp*kPtsPerWarp .. ---- */
#pragma unroll 1
        for (int i = 0; i < kPtsPerWarp; i++) {
            const int pt = warp * kPtsPerWarp + i;
            const double2 *ch = reinterpret_cast<const double2 *>(&S.chi[buf][pt][0]);
            double v = (MODE == MODE_EXPAND_ATOMIC) ? 0.0 : S.V[buf][lane][pt];
            if (MODE != MODE_PROJECT) {
                double d0 = 0.0, d1 = 0.0;
#pragma unroll
                for (int p = 0; p < NP / 2; p++) {
                    const double2 c2 = ch[p];
                    d0 = fma(c2.x, beta[2 * p], d0);
                    d1 = fma(c2.y, beta[2 * p + 1], d1);
                }
                v += d0 + d1;
                S.V[buf][lane][pt] = v;
            }
            if (MODE == MODE_PROJECT || MODE == MODE_FUSED) {
#pragma unroll
                for (int p = 0; p < NP / 2; p++) {
                    const double2 c2 = ch[p];
                    acc[2 * p] = fma(c2.x, v, acc[2 * p]);
                    acc[2 * p + 1] = fma(c2.y, v, acc[2 * p + 1]);
                }
            }
        }
        __syncthreads();

        /* ---- C: scatter back (lanes <-> points) ---- */
        if (MODE != MODE_PROJECT) {
#pragma unroll
            for (int h = 0; h < H; h++) {
                const int pl = lane + 32 * h;
                const int ps = S.pos[buf][pl];
                if (ps < 0) continue;
#pragma unroll
                for (int c = 0; c < kCPW; c++) {
                    const int cl = warp * kCPW + c;
                    const int wc = wc0 + cl;
                    if (wc >= nwc) continue;
                    double *base = vec + ((size_t)(wc / WORDS) * ld) * WORDS + (wc % WORDS);
                    const double v = S.V[buf][cl][pl];
                    if (MODE == MODE_EXPAND_ATOMIC) {
                        atomicAdd(base + (size_t)ps * WORDS, v);
                    } else {
                        base[(size_t)ps * WORDS] = v;
                    }
                }
            }
        }
    }
    cp_async_wait<0>();

    /* ---- alpha_next[J] = dV * phase_J * sum over warps of acc ----------------------------------------- */
    if (MODE == MODE_PROJECT || MODE == MODE_FUSED) {
        __syncthreads();
        double *red = reinterpret_cast<double *>(smem_raw); /* [kWarps][NP][kCols+1] */
#pragma unroll
        for (int p = 0; p < NP; p++) red[(warp * NP + p) * (kCols + 1) + lane] = acc[p];
        __syncthreads();
        double2 ph = make_double2(1.0, 0.0);
        if (WORDS == 2) ph = nl.img_phase[J];
        double *an = alpha_next + (size_t)nl.img_aoff[J] * ncol * WORDS;
        for (int t = threadIdx.x; t < nproj * (kCols / WORDS); t += kThreads) {
            const int p = t % nproj, cl = t / nproj; /* cl: data column inside the CTA */
            const int dc = wc0 / WORDS + cl;
            if (dc >= ncol) continue;
            double sr = 0.0, si = 0.0;
#pragma unroll
            for (int w8 = 0; w8 < kWarps; w8++) {
                sr += red[(w8 * NP + p) * (kCols + 1) + cl * WORDS];
                if (WORDS == 2) si += red[(w8 * NP + p) * (kCols + 1) + cl * WORDS + 1];
            }
            double *dst = an + ((size_t)dc * nproj + p) * WORDS;
            if (WORDS == 2) {
                dst[0] = dV * (sr * ph.x - si * ph.y);
                dst[1] = dV * (sr * ph.y + si * ph.x);
            } else {
                dst[0] = dV * sr;
            }
        }
    }
}

/* Per-atom sums of the alpha partials (layout per atom as in a partial: [column][projector]), summed over the atom's images/segments in list
 * order (deterministic).  Launched only when some atom has many partials (segmented spheres of small systems,
 * atoms with many periodic images): every consumer CTA then reads one row set instead of looping over them. */
__global__ void alpha_reduce_kernel(const NlocView nl, const double *__restrict__ alpha, double *__restrict__ asum,
                                    const size_t rowlen)
{
    const int atom = blockIdx.y;
    const int ip0 = nl.IP_displ[atom];
    const int nproj = nl.IP_displ[atom + 1] - ip0;
    const size_t n = (size_t)nproj * rowlen;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int jj = nl.atom_img_off[atom]; jj < nl.atom_img_off[atom + 1]; jj++)
            acc += alpha[(size_t)nl.img_aoff[nl.atom_img[jj]] * rowlen + t];
        asum[(size_t)ip0 * rowlen + t] = acc;
    }
}

template <int NP, int WORDS, int MODE, class SH>
int launch_shape(chefsi_ctx *ctx, const NlocView &v, const double *aprev, double *anext, double *vec, size_t ld, int ncol,
                 double scale)
{
    constexpr int kThreads = SH::T;
    constexpr size_t red_bytes = (size_t)SH::kWarps * NP * (kCols + 1) * sizeof(double);
    constexpr size_t smem = sizeof(Smem<NP, SH>) > red_bytes ? sizeof(Smem<NP, SH>) : red_bytes;
    auto kern = v.alpha_sum ? nloc_kernel<NP, WORDS, MODE, SH, true> : nloc_kernel<NP, WORDS, MODE, SH, false>;
    {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { chefsi_fail(ctx, "cudaFuncSetAttribute(nloc): %s", cudaGetErrorString(e)); return -1; }
    }
    const int ngroups = (ncol * WORDS + kCols - 1) / kCols;
    const long long nblk = (long long)ctx->nl.n_img * ngroups;
    if (nblk > 0x7fffffffLL) { chefsi_fail(ctx, "nloc: too many CTAs"); return -1; }
    kern<<<(unsigned)nblk, kThreads, smem, ctx->stream>>>(v, aprev, anext, vec, ld, ncol, ngroups, scale, ctx->grid.dV);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "nloc launch: %s", cudaGetErrorString(e)); return -1; }
    return 1;
}

template <int NP, int WORDS, int MODE>
int launch_mode(chefsi_ctx *ctx, const NlocView &v, const double *aprev, double *anext, double *vec, size_t ld, int ncol,
                double scale)
{
    /* pipeline shape: 256 threads, 2-stage ring of 64-point chunks.  128-thread CTAs, 3- and 4-stage rings and
       128-point chunks were measured and are 3-30 % slower (profiles/r1_exp_nloc_shapes.log): the gather sits at
       the HBM random-access rate, not at a latency the pipeline depth could hide. */
    return launch_shape<NP, WORDS, MODE, Shape<256, 2, 64>>(ctx, v, aprev, anext, vec, ld, ncol, scale);
}

template <int NP, int WORDS>
int launch_np(chefsi_ctx *ctx, int mode, const NlocView &v, const double *aprev, double *anext, double *vec, size_t ld,
              int ncol, double scale)
{
    switch (mode) {
    case MODE_PROJECT: return launch_mode<NP, WORDS, MODE_PROJECT>(ctx, v, aprev, anext, vec, ld, ncol, scale);
    case MODE_FUSED: return launch_mode<NP, WORDS, MODE_FUSED>(ctx, v, aprev, anext, vec, ld, ncol, scale);
    case MODE_EXPAND: return launch_mode<NP, WORDS, MODE_EXPAND>(ctx, v, aprev, anext, vec, ld, ncol, scale);
    default: return launch_mode<NP, WORDS, MODE_EXPAND_ATOMIC>(ctx, v, aprev, anext, vec, ld, ncol, scale);
    }
}

}  // namespace

/* projector padding the kernels are instantiated for (chosen from the largest nproj of any atom) */
int nloc_padded_nproj(int max_nproj)
{
    if (max_nproj <= 8) return 8;
    if (max_nproj <= 14) return 14;
    if (max_nproj <= 20) return 20;
    if (max_nproj <= 26) return 26;
    if (max_nproj <= 32) return 32;
    return -1;
}

static int ensure_alpha(chefsi_ctx *ctx, int ncol, int words)
{
    const size_t need = (size_t)ctx->nl.img_proj_total * ncol * sizeof(double) * words;
    if (need > ctx->alpha_bytes) {
        for (int i = 0; i < 2; i++) { cudaFree(ctx->d_alpha[i]); ctx->d_alpha[i] = nullptr; }
        ctx->alpha_bytes = 0;
        for (int i = 0; i < 2; i++) {
            cudaError_t e = cudaMalloc(&ctx->d_alpha[i], need);
            if (e != cudaSuccess) { chefsi_fail(ctx, "cudaMalloc(alpha, %zu): %s", need, cudaGetErrorString(e)); return -1; }
        }
        ctx->alpha_bytes = need;
    }
    const size_t ns = (ctx->nl.max_parts > ctx->alpha_reduce_min) ? (size_t)ctx->nl.ntot * ncol * sizeof(double) * words : 0;
    if (ns > ctx->alpha_sum_bytes) {
        cudaFree(ctx->d_alpha_sum); ctx->d_alpha_sum = nullptr;
        ctx->alpha_sum_bytes = 0;
        cudaError_t e = cudaMalloc(&ctx->d_alpha_sum, ns);
        if (e != cudaSuccess) { chefsi_fail(ctx, "cudaMalloc(alpha sums, %zu): %s", ns, cudaGetErrorString(e)); return -1; }
        ctx->alpha_sum_bytes = ns;
    }
    return 0;
}

int nloc_ensure_alpha(chefsi_ctx *ctx, int ncol, bool is_complex) { return ensure_alpha(ctx, ncol, is_complex ? 2 : 1); }

/* per-atom sums of the current alpha partials into ctx->d_alpha_sum (see alpha_reduce_kernel) */
int launch_alpha_reduce(chefsi_ctx *ctx, int ncol, bool is_complex)
{
    NlocDev &d = ctx->nl;
    const int words = is_complex ? 2 : 1;
    if (ensure_alpha(ctx, ncol, words)) return -1;
    if (!ctx->d_alpha_sum) { chefsi_fail(ctx, "alpha reduce: no sum buffer"); return -1; }
    NlocView v{};
    v.IP_displ = d.IP_displ; v.img_aoff = d.img_aoff; v.atom_img_off = d.atom_img_off; v.atom_img = d.atom_img;
    const size_t rowlen = (size_t)ncol * words;
    const unsigned gx = (unsigned)std::min<size_t>(64, ((size_t)d.max_nproj * rowlen + 255) / 256);
    alpha_reduce_kernel<<<dim3(gx ? gx : 1, (unsigned)d.n_atom), 256, 0, ctx->stream>>>(
        v, reinterpret_cast<const double *>(ctx->d_alpha[ctx->alpha_cur]), (double *)ctx->d_alpha_sum, rowlen);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "alpha reduce launch: %s", cudaGetErrorString(e)); return -1; }
    return 1;
}

/* mode: NLOC_PROJECT  alpha(cur) = dV * phase * Chi^T vec                       (vec is read only)
 *       NLOC_FUSED    vec += scale * Chi Gamma alpha(cur); alpha(next) = proj(vec); cur <-> next
 *       NLOC_EXPAND   vec += scale * Chi Gamma alpha(cur)   (atomics when spheres overlap)            */
int launch_nloc(chefsi_ctx *ctx, int mode, void *vec, size_t ld, int ncol, double scale, bool is_complex)
{
    NlocDev &d = ctx->nl;
    if (d.n_img == 0 || d.ntot == 0 || ncol <= 0) return 0;
    const int words = is_complex ? 2 : 1;
    if (ensure_alpha(ctx, ncol, words)) return -1;
    NlocView v{d.IP_displ, d.gamma, d.img_atom, d.img_ndc, d.img_aoff, d.pos_off, d.chiT_off, d.grid_pos, d.chiT,
               d.img_phase, d.atom_img_off, d.atom_img, nullptr};
    int kmode = mode;
    if (mode == NLOC_FUSED && d.overlap) { chefsi_fail(ctx, "nloc: fused mode needs disjoint spheres"); return -1; }
    if (mode == NLOC_EXPAND && d.overlap) kmode = MODE_EXPAND_ATOMIC;
    if (mode != NLOC_PROJECT) { ctx->stats.last_nloc_atomic = (kmode == MODE_EXPAND_ATOMIC); ctx->stats.last_alpha_reduced = 0; }
    const double *aprev = reinterpret_cast<const double *>(ctx->d_alpha[ctx->alpha_cur]);
    int n_extra = 0;
    if (mode != NLOC_PROJECT && d.max_parts > ctx->alpha_reduce_min) { /* consumers read per-atom sums instead of looping over many partials */
        if (!ctx->alpha_sum_external) {
            n_extra = launch_alpha_reduce(ctx, ncol, is_complex);
            if (n_extra < 0) return -1;
        }
        v.alpha_sum = (const double *)ctx->d_alpha_sum;
        ctx->stats.last_alpha_reduced = 1;
    }
    double *anext = reinterpret_cast<double *>(ctx->d_alpha[mode == NLOC_FUSED ? ctx->alpha_cur ^ 1 : ctx->alpha_cur]);
    double *p = reinterpret_cast<double *>(vec);
    int n = -1;
#define CHEFSI_NLOC_CASE(NP)                                                                              \
    case NP:                                                                                              \
        n = is_complex ? launch_np<NP, 2>(ctx, kmode, v, aprev, anext, p, ld, ncol, scale)                \
                       : launch_np<NP, 1>(ctx, kmode, v, aprev, anext, p, ld, ncol, scale);               \
        break;
    switch (d.np_pad) {
        CHEFSI_NLOC_CASE(8)
        CHEFSI_NLOC_CASE(14)
        CHEFSI_NLOC_CASE(20)
        CHEFSI_NLOC_CASE(26)
        CHEFSI_NLOC_CASE(32)
    default: chefsi_fail(ctx, "nloc: more than 32 projectors per atom not supported"); return -1;
    }
#undef CHEFSI_NLOC_CASE
    if (n < 0) return -1;
    if (mode == NLOC_FUSED) ctx->alpha_cur ^= 1;
    return n + n_extra;
}
