/*
 * stencil_stream_kpt.cu -- streaming fused Chebyshev step for orthogonal cells, k-point (complex) data (sm_100a).
 *
 *   out = s1 * ( (-1/2 Lap + Veff + c) x ) - s2 * xprev      (complex FP64, radius-6 star stencil, real weights)
 *
 * The k-point sibling of stencil_stream_dense.cu: it replaces, per Chebyshev degree, the reference's haloed copy
 * with Bloch-phase halos + stencil_3axis_thread_complex_v2 (lapVecRoutinesKpt.c:120-164,179-513) + the three
 * scale/axpy/swap passes of ChebyshevFiltering_kpt (eigenSolverKpt.c:500-531).  All stencil weights are real, so
 * a complex column (re, im interleaved = C99 double _Complex) is treated as a real array with 2 Nx doubles per
 * row whose x neighbours are 2 r doubles away; only the halo values that cross a periodic face need complex
 * arithmetic: they are multiplied by exp(-+ i k_d L_d) (lapVecRoutinesKpt.c:370-382,452-462).
 *
 * Same machinery as the real kernel (persistent CTA per SM, producer warpgroup + 5-stage TMA/mbarrier ring,
 * setmaxnreg, z in registers, round barrier); what differs:
 *   - a thread owns ONE complex point (a 16-byte chunk) of two rows, i.e. again 4 doubles of queues;
 *     a warp covers 16 complex points x 4 rows, the tile is 16 complex x 32 rows;
 *   - the haloed tile is 16 + 2*6 complex wide (58 doubles pitch); the x window of a point is 13 chunks per row;
 *   - periodic-x strips are 7 complex wide; Veff is a real tile (one double per point);
 *   - phases: the x strips and the wrapped y rows of a stage are multiplied by their face's phase in shared
 *     memory by the three spare warps of the producer warpgroup, between the TMA completion and the hand-over to
 *     the consumers; the planes beyond a periodic z face are multiplied when their centre values are read.  All
 *     consumers run the same instruction stream as the real kernel (with a 13-chunk x window).
 * Because complex rows are always 16-byte aligned there is no parity restriction on Nx.
 */
#include <cuda.h>

#include "chefsi_internal.h"
#include "tma_ring.cuh"

namespace {

using namespace tma_ring;

constexpr int R = 6;        /* FD radius (complex points) */
constexpr int HT = 8;       /* top halo rows held in the tile (6 used) */
constexpr int TXC = 16;     /* tile width in complex points */
constexpr int TY = 32;      /* tile height */
constexpr int SWC = 7;      /* periodic-x strip width in complex points (7 chunks: odd) */
constexpr int kStages = 5;

struct Cfg {
    static constexpr int YP = 2 * TXC + 4 * R + 2;   /* haloed tile pitch in doubles: 29 chunks */
    static constexpr int YROWS = HT + TY + R;
    static constexpr int SP = 2 * SWC;               /* strip pitch in doubles */
    static constexpr int XP = 2 * TXC + 2;           /* xprev tile pitch in doubles */
    static constexpr int VP = TXC + 2;               /* Veff tile pitch in doubles */
    static constexpr int Y_BYTES = ((YP * YROWS * 8 + 127) / 128) * 128;
    static constexpr int S_BYTES = ((SP * TY * 8 + 127) / 128) * 128;
    static constexpr int X_BYTES = ((XP * TY * 8 + 127) / 128) * 128;
    static constexpr int V_BYTES = ((VP * TY * 8 + 127) / 128) * 128;
    static constexpr int OFF_L = Y_BYTES;
    static constexpr int OFF_R = OFF_L + S_BYTES;
    static constexpr int OFF_V = OFF_R + S_BYTES;
    static constexpr int OFF_X = OFF_V + V_BYTES;
    static constexpr int STAGE_BYTES = OFF_X + X_BYTES;
    static constexpr int CONSUMER_WARPS = 8;
    static constexpr int THREADS = (CONSUMER_WARPS + 4) * 32; /* + producer warpgroup, see stencil_stream_dense.cu */
    static constexpr int PRODUCER_REGS = 40, CONSUMER_REGS = 232;
    static constexpr size_t SMEM = (size_t)kStages * STAGE_BYTES + 3 * kStages * sizeof(unsigned long long);
    static_assert((YP * HT * 8) % 128 == 0 && (YP * TY * 8) % 128 == 0, "box starts must be 128-byte aligned");
    static_assert(TY == 4 * CONSUMER_WARPS, "16 points x 4 rows per warp");
};

struct KptDesc {
    int Nx, Ny, Nz;          /* complex points */
    int bc[3];
    int ntx, nty;
    double coef0;            /* s1 * (coef0 + c) */
    double wx[R + 1], wy[R + 1], wz[R + 1]; /* s1 * weights */
    /* Bloch phases of halo values taken from across the low (m) / high (p) periodic face of each axis */
    double phm_re[3], phm_im[3], php_re[3], php_im[3];
};

struct KptMaps {
    CUtensorMap y_full, y_top, y_body, y_bot, y_strip, xprev, veff;
};

/* (re, im) * (pr + i pi) */
__device__ __forceinline__ double2 cmul(double2 a, double pr, double pi)
{
    return make_double2(a.x * pr - a.y * pi, a.x * pi + a.y * pr);
}

/* ---- Bloch phases of a freshly loaded stage -----------------------------------------------------
 * Every chunk of an x strip and of the 6 wrapped top / bottom rows was fetched from the other side of the cell:
 * the three spare warps of the producer warpgroup multiply them in place by the face's phase between the TMA
 * completion (`landed` barrier) and the hand-over to the consumers (`full` barrier), a few planes ahead of them;
 * the consumers of boundary and interior tiles run the same instruction stream. */
__device__ __forceinline__ void fix_chunk(unsigned char *p, double pr, double pi)
{
    double2 *q = reinterpret_cast<double2 *>(p);
    *q = cmul(*q, pr, pi);
}
__device__ __forceinline__ void fix_phases(const KptDesc &d, unsigned char *stage, int ft, int nthreads, int x0,
                                           bool need_l, bool need_r, bool wrap_top, bool wrap_bot)
{
    /* x strips: multiplied by the face's phase and written into the (zero-filled) halo columns of the tile, so that
       the consumers read every x window with constant offsets */
    double *tile = reinterpret_cast<double *>(stage);
    if (need_l) {
        const double *sl = reinterpret_cast<const double *>(stage + Cfg::OFF_L);
        for (int c = ft; c < R * TY; c += nthreads) {
            const int row = c / R, j = c % R, gi = x0 - R + j; /* tile chunk j <-> complex point gi */
            if (gi < 0)
                *reinterpret_cast<double2 *>(tile + (HT + row) * Cfg::YP + 2 * j) =
                    cmul(*reinterpret_cast<const double2 *>(sl + row * Cfg::SP + 2 * (gi + SWC)), d.phm_re[0], d.phm_im[0]);
        }
    }
    if (need_r) {
        const double *sr = reinterpret_cast<const double *>(stage + Cfg::OFF_R);
        for (int c = ft; c < R * TY; c += nthreads) {
            const int row = c / R, j = c % R, col = d.Nx + j - (x0 - R); /* tile chunk of complex point Nx + j */
            if (col < Cfg::YP / 2)
                *reinterpret_cast<double2 *>(tile + (HT + row) * Cfg::YP + 2 * col) =
                    cmul(*reinterpret_cast<const double2 *>(sr + row * Cfg::SP + 2 * j), d.php_re[0], d.php_im[0]);
        }
    }
    constexpr int CPR = Cfg::YP / 2; /* chunks per tile row */
    if (wrap_top || wrap_bot) {
        for (int c = ft; c < R * CPR; c += nthreads) {
            const int row = c / CPR, col = c % CPR;
            if (wrap_top) fix_chunk(stage + ((HT - R + row) * Cfg::YP + 2 * col) * 8, d.phm_re[1], d.phm_im[1]);
            if (wrap_bot) fix_chunk(stage + ((HT + TY + row) * Cfg::YP + 2 * col) * 8, d.php_re[1], d.php_im[1]);
        }
    }
}

/* ---- one plane step of a consumer thread -------------------------------------------------------
 * The thread owns the complex point xp of rows r0 and r0 + 1; value index = 2 * row + (0: re, 1: im).
 * U = (p + 7) mod 7 (compile time): register-queue rotation by renaming.
 * The periodic-x strips (phase applied) have been merged into the tile and the wrapped y rows multiplied by their
 * phase in shared memory (fix_phases) by now: every window is read with constant offsets. */
/* yoff / voff / xoff: the thread's offsets (doubles) into the haloed tile, the Veff tile and the xprev tile, kept in
 * registers by the caller; HASV / HASX: the step has a Veff / an xprev operand; STEADY: R <= p < Nz, the plane is
 * inside the grid and so is the plane it completes (no boundary case distinctions); dst: running output pointer. */
template <int U, bool HASV, bool HASX, bool STEADY>
__device__ __forceinline__ void consume_plane(const KptDesc &d, const StepArgs &a, const unsigned char *stage, int p,
                                              bool active, bool act0, bool act1, uint32_t yoff, uint32_t voff, uint32_t xoff,
                                              double *__restrict__ &dst,
                                              size_t plane_doubles, double (&in)[7][4], double (&acc)[7][4],
                                              bool plane_is_zero)
{
    const int Nz = d.Nz;
    const bool interior = STEADY || ((p >= 0) && (p < Nz));
    const int o = p - R;
    const bool emit = STEADY || (o >= 0 && o < Nz);
    const double *ytile = reinterpret_cast<const double *>(stage);
    const double *vtile = reinterpret_cast<const double *>(stage + Cfg::OFF_V);
    const double *xtile = reinterpret_cast<const double *>(stage + Cfg::OFF_X);

    double v[4] = {0, 0, 0, 0};
    if (active && (STEADY || !plane_is_zero)) {
        const double *cp = ytile + yoff; /* this point, row r0: (r0 + HT) * YP + 2 xp + 2 R */
        if (interior) {
            double ve[2] = {0.0, 0.0};
            if (HASV) {
                ve[0] = vtile[voff]; /* r0 * VP + xp */
                ve[1] = vtile[voff + Cfg::VP];
            }
            /* d.w*, d.coef0 carry the recurrence scale s1 (and the shift c), see launch */
            double sx[4], sy[4], sz[4];
#pragma unroll
            for (int i = 0; i < 4; i++) sz[i] = d.wz[1] * in[(U - 1 + 7) % 7][i];
#pragma unroll
            for (int r = 2; r <= R; r++)
#pragma unroll
                for (int i = 0; i < 4; i++) sz[i] = fma(d.wz[r], in[(U - r + 7) % 7][i], sz[i]);
            /* x: one row at a time (13 chunks = 26 doubles live), chunk t = point x - 6 + t */
#pragma unroll
            for (int row = 0; row < 2; row++) {
                double2 w[13];
#pragma unroll
                for (int t = 0; t < 13; t++) {
                    w[t] = *reinterpret_cast<const double2 *>(cp + row * Cfg::YP + 2 * (t - 6));
                }
                v[2 * row] = w[6].x;
                v[2 * row + 1] = w[6].y;
                const double diag = HASV ? fma(a.s1, ve[row], d.coef0) : d.coef0;
                double s0 = fma(d.wx[1], w[5].x + w[7].x, diag * w[6].x);
                double s1 = fma(d.wx[1], w[5].y + w[7].y, diag * w[6].y);
#pragma unroll
                for (int r = 2; r <= R; r++) {
                    s0 = fma(d.wx[r], w[6 - r].x + w[6 + r].x, s0);
                    s1 = fma(d.wx[r], w[6 - r].y + w[6 + r].y, s1);
                }
                sx[2 * row] = s0;
                sx[2 * row + 1] = s1;
            }
            /* y: up[k] = row r0-k, dn[k] = row r0+1+k (k = 1..6); row r0 pairs up[k] with (k == 1 ? own row 1 : dn[k-1]),
               row r0+1 pairs (k == 1 ? own row 0 : up[k-1]) with dn[k]; the y chains start from the z sums */
            double2 up[R + 1], dn[R + 1];
            up[0] = make_double2(v[0], v[1]);
            dn[0] = make_double2(v[2], v[3]);
#pragma unroll
            for (int k = 1; k <= R; k++) {
                up[k] = *reinterpret_cast<const double2 *>(cp - k * Cfg::YP);
                dn[k] = *reinterpret_cast<const double2 *>(cp + (1 + k) * Cfg::YP);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) sy[i] = sz[i];
#pragma unroll
            for (int k = 1; k <= R; k++) {
                const double a0 = up[k].x + dn[k - 1].x, a1 = up[k].y + dn[k - 1].y;
                const double b0 = up[k - 1].x + dn[k].x, b1 = up[k - 1].y + dn[k].y;
                sy[0] = fma(d.wy[k], a0, sy[0]); sy[1] = fma(d.wy[k], a1, sy[1]);
                sy[2] = fma(d.wy[k], b0, sy[2]); sy[3] = fma(d.wy[k], b1, sy[3]);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) acc[U][i] = sx[i] + sy[i];
        } else { /* plane beyond a periodic z face: only its z contributions are needed, times the face's phase */
            double2 c0 = *reinterpret_cast<const double2 *>(cp);
            double2 c1 = *reinterpret_cast<const double2 *>(cp + Cfg::YP);
            const double pr = p < 0 ? d.phm_re[2] : d.php_re[2], pi = p < 0 ? d.phm_im[2] : d.php_im[2];
            c0 = cmul(c0, pr, pi);
            c1 = cmul(c1, pr, pi);
            v[0] = c0.x; v[1] = c0.y; v[2] = c1.x; v[3] = c1.y;
        }
    }
    if (STEADY || p >= 0) { /* scatter the z terms into the 6 accumulators behind this plane */
#pragma unroll
        for (int r = 1; r <= R; r++)
#pragma unroll
            for (int i = 0; i < 4; i++) acc[(U - r + 7) % 7][i] = fma(d.wz[r], v[i], acc[(U - r + 7) % 7][i]);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) in[U][i] = v[i];

    if (emit && active) {
        double res[4];
        if (HASX) {
            const double2 q0 = *reinterpret_cast<const double2 *>(xtile + xoff); /* r0 * XP + 2 xp */
            const double2 q1 = *reinterpret_cast<const double2 *>(xtile + xoff + Cfg::XP);
            res[0] = fma(-a.s2, q0.x, acc[(U + 1) % 7][0]);
            res[1] = fma(-a.s2, q0.y, acc[(U + 1) % 7][1]);
            res[2] = fma(-a.s2, q1.x, acc[(U + 1) % 7][2]);
            res[3] = fma(-a.s2, q1.y, acc[(U + 1) % 7][3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) res[i] = acc[(U + 1) % 7][i];
        }
        if (act0) stg128(dst, res[0], res[1]);
        if (act1) stg128(dst + 2 * d.Nx, res[2], res[3]);
    }
    if (emit) dst += plane_doubles; /* planes are emitted in order */
}

template <bool HASV, bool HASX>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
stream_kpt_kernel(const __grid_constant__ KptMaps maps, const __grid_constant__ KptDesc d, const StepArgs a, const int nitems,
                  unsigned int *__restrict__ sync_counter, const unsigned int sync_base)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *ring = smem_raw;
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)kStages * Cfg::STAGE_BYTES);
    uint64_t *empty = full + kStages;
    uint64_t *landed = empty + kStages; /* TMA completion; the fix-up warps turn it into `full` */

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&landed[s], 1);
            mbar_init(&empty[s], Cfg::CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int Nx = d.Nx, Ny = d.Ny, Nz = d.Nz;
    const bool xper = (d.bc[0] == 0), yper = (d.bc[1] == 0), zper = (d.bc[2] == 0);
    const size_t plane_doubles = (size_t)2 * Nx * Ny;
    /* ring position (stage, parity of its current use), continues across work items */
    uint32_t rs = 0, rpar = 0;
#define CHEFSI_RING_ADVANCE() do { if (++rs == (uint32_t)kStages) { rs = 0; rpar ^= 1u; } } while (0)

    if (warp >= Cfg::CONSUMER_WARPS) {
        /* ================= producer warpgroup (one elected lane issues the TMA boxes) ================= */
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::PRODUCER_REGS));
        if (warp == Cfg::CONSUMER_WARPS && lane == 0) {
            unsigned int round = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, round++) {
                const int tile = item % (d.ntx * d.nty), n = item / (d.ntx * d.nty);
                const int x0 = tile_origin(tile % d.ntx, TXC, Nx), y0 = tile_origin(tile / d.ntx, TY, Ny);
                const bool wrap_top = yper && (y0 - R < 0), wrap_bot = yper && (y0 + TY + R > Ny);
                const bool split_y = wrap_top || wrap_bot;
                const int ytop = wrap_top ? y0 - HT + Ny : y0 - HT;
                const int ybot = wrap_bot ? y0 + TY - Ny : y0 + TY;
                const bool need_l = xper && (x0 - R < 0), need_r = xper && (x0 + TXC + R > Nx);
                const uint32_t ybytes = (uint32_t)(Cfg::YP * Cfg::YROWS * 8 + (need_l ? Cfg::SP * TY * 8 : 0) +
                                                   (need_r ? Cfg::SP * TY * 8 : 0));
                if (sync_counter) { /* round barrier between the producers of all CTAs: see stencil_stream_dense.cu */
                    const unsigned int in_round = (unsigned int)min((long long)gridDim.x, (long long)nitems - (long long)round * gridDim.x);
                    const unsigned int done_before = round * gridDim.x;
                    __threadfence();
                    atomicAdd(sync_counter, 1u);
                    const unsigned int target = sync_base + done_before + in_round;
                    unsigned int spins = 0;
                    while ((int)(*(volatile unsigned int *)sync_counter - target) < 0 && ++spins < (1u << 22)) __nanosleep(64);
                    if (spins >= (1u << 22)) atomicAdd(sync_counter + 1, 1u); /* gave up: reported by chefsi_get_stats */
                }
                for (int p = -R; p < Nz + R; p++) {
                    int kz = p;
                    const bool interior = (p >= 0 && p < Nz);
                    if (p < 0) kz += Nz; else if (p >= Nz) kz -= Nz;
                    const int o = p - R;
                    const bool need_y = interior || zper;
                    const bool need_v = interior && HASV;
                    const bool need_x = (o >= 0 && o < Nz) && HASX;
                    if (!need_y && !need_x) continue;
                    const int s = (int)rs;
                    unsigned char *stage = ring + (size_t)s * Cfg::STAGE_BYTES;
                    mbar_wait(&empty[s], rpar ^ 1u);
                    mbar_expect_tx(&landed[s], (need_y ? ybytes : 0u) + (uint32_t)((need_v ? Cfg::VP * TY * 8 : 0) +
                                                                                (need_x ? Cfg::XP * TY * 8 : 0)));
                    const int xd = 2 * (x0 - R); /* x coordinates of the complex maps are in doubles */
                    if (need_y) {
                        if (!split_y) {
                            tma_load_4d(stage, &maps.y_full, xd, y0 - HT, kz, n, &landed[s]);
                        } else {
                            tma_load_4d(stage, &maps.y_top, xd, ytop, kz, n, &landed[s]);
                            tma_load_4d(stage + Cfg::YP * HT * 8, &maps.y_body, xd, y0, kz, n, &landed[s]);
                            tma_load_4d(stage + Cfg::YP * (HT + TY) * 8, &maps.y_bot, xd, ybot, kz, n, &landed[s]);
                        }
                        if (need_l) tma_load_4d(stage + Cfg::OFF_L, &maps.y_strip, 2 * (Nx - SWC), y0, kz, n, &landed[s]);
                        if (need_r) tma_load_4d(stage + Cfg::OFF_R, &maps.y_strip, 0, y0, kz, n, &landed[s]);
                    }
                    if (need_v) tma_load_4d(stage + Cfg::OFF_V, &maps.veff, x0, y0, p, 0, &landed[s]);
                    if (need_x) tma_load_4d(stage + Cfg::OFF_X, &maps.xprev, 2 * x0, y0, o, n, &landed[s]);
                    CHEFSI_RING_ADVANCE();
                }
            }
        } else if (warp > Cfg::CONSUMER_WARPS) {
            /* ---- fix-up warps: same (item, plane) sequence as the producer lane ---- */
            const int ft = (int)threadIdx.x - (Cfg::CONSUMER_WARPS + 1) * 32; /* 0 .. 95 */
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int tile = item % (d.ntx * d.nty);
                const int x0 = tile_origin(tile % d.ntx, TXC, Nx), y0 = tile_origin(tile / d.ntx, TY, Ny);
                const bool wrap_top = yper && (y0 - R < 0), wrap_bot = yper && (y0 + TY + R > Ny);
                const bool need_l = xper && (x0 - R < 0), need_r = xper && (x0 + TXC + R > Nx);
                const bool fix_any = wrap_top || wrap_bot || need_l || need_r;
                for (int p = -R; p < Nz + R; p++) {
                    const bool interior = (p >= 0 && p < Nz);
                    const int o = p - R;
                    const bool need_y = interior || zper;
                    const bool need_x = (o >= 0 && o < Nz) && HASX;
                    if (!need_y && !need_x) continue;
                    const int s = (int)rs;
                    unsigned char *stage = ring + (size_t)s * Cfg::STAGE_BYTES;
                    mbar_wait(&landed[s], rpar);
                    if (fix_any && need_y) fix_phases(d, stage, ft, 96, x0, need_l, need_r, wrap_top, wrap_bot);
                    asm volatile("bar.sync 2, 96;" ::: "memory");
                    if (ft == 0) mbar_arrive(&full[s]);
                    CHEFSI_RING_ADVANCE();
                }
            }
        }
    } else {
        /* ================= consumer warps ================= */
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::CONSUMER_REGS));
        const int xp = lane & 15;                    /* complex point inside the tile */
        const int r0 = warp * 4 + 2 * (lane >> 4);   /* first of the thread's two rows */
        /* made opaque so that they live in registers instead of being re-derived from %tid in every plane step */
        uint32_t yoff = (uint32_t)((r0 + HT) * Cfg::YP + 2 * xp + 2 * R), voff = (uint32_t)(r0 * Cfg::VP + xp);
        uint32_t xoff = (uint32_t)(r0 * Cfg::XP + 2 * xp), lane0 = (lane == 0) ? 1u : 0u;
        asm volatile("" : "+r"(yoff), "+r"(voff), "+r"(xoff), "+r"(lane0));
        double in[7][4], acc[7][4];
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int tile = item % (d.ntx * d.nty), n = item / (d.ntx * d.nty);
            const int tx = tile % d.ntx, ty = tile / d.ntx;
            const int x0 = tile_origin(tx, TXC, Nx), y0 = tile_origin(ty, TY, Ny);
            const int gx = x0 + xp, gy = y0 + r0;
            /* a shifted last tile overlaps its neighbour: only the not yet covered points are computed */
            const bool act0 = (gx >= tx * TXC) && (gy >= ty * TY);
            const bool act1 = (gx >= tx * TXC) && (gy + 1 >= ty * TY);
            const bool active = act1;
            double *dst = reinterpret_cast<double *>(a.out) + 2 * ((size_t)n * a.ld + (size_t)gy * Nx + gx); /* output plane 0 */
#pragma unroll
            for (int u = 0; u < 7; u++)
#pragma unroll
                for (int j = 0; j < 4; j++) { in[u][j] = 0.0; acc[u][j] = 0.0; }

#define CHEFSI_STEP(U, STEADY)                                                                           \
    if ((STEADY) || p + (U) < Nz + R) {                                                                  \
        const int pp = p + (U);                                                                          \
        const bool zplane = !(STEADY) && !zper && (pp < 0 || pp >= Nz); /* Dirichlet z: the plane is zero */ \
        const bool use_stage = (STEADY) || !zplane || (pp - R >= 0 && pp - R < Nz && HASX);              \
        const unsigned char *stage = ring;                                                               \
        int s = 0;                                                                                       \
        if (use_stage) {                                                                                 \
            s = (int)rs;                                                                                 \
            stage = ring + (size_t)s * Cfg::STAGE_BYTES;                                                 \
            mbar_wait(&full[s], rpar);                                                                   \
        }                                                                                                \
        consume_plane<(U), HASV, HASX, (STEADY)>(d, a, stage, pp, active, act0, act1, yoff, voff, xoff, dst, \
                                                 plane_doubles, in, acc, zplane);                       \
        if (use_stage) {                                                                                 \
            __syncwarp();                                                                                \
            if (lane0) mbar_arrive(&empty[s]);                                                           \
            CHEFSI_RING_ADVANCE();                                                                       \
        }                                                                                                \
    }
#define CHEFSI_GROUP(STEADY)                                                                             \
    {                                                                                                    \
        if ((STEADY) || p + 0 >= -R) { CHEFSI_STEP(0, STEADY) }                                          \
        CHEFSI_STEP(1, STEADY)                                                                           \
        CHEFSI_STEP(2, STEADY)                                                                           \
        CHEFSI_STEP(3, STEADY)                                                                           \
        CHEFSI_STEP(4, STEADY)                                                                           \
        CHEFSI_STEP(5, STEADY)                                                                           \
        CHEFSI_STEP(6, STEADY)                                                                           \
    }
            /* groups whose seven planes all lie in [R, Nz) take a body without the boundary case distinctions */
            int p = -R - 1;
            for (; p < R && p < Nz + R; p += 7) CHEFSI_GROUP(false)
            for (; p + 6 < Nz; p += 7) CHEFSI_GROUP(true)
            for (; p < Nz + R; p += 7) CHEFSI_GROUP(false)
#undef CHEFSI_GROUP
#undef CHEFSI_STEP
        }
    }
#undef CHEFSI_RING_ADVANCE
}

/* ---- host side ---------------------------------------------------------------------------- */
/* 4-D view (x in doubles, y, z, column) of a block of dense columns of `words` doubles per point (2: complex
 * orbitals, 1: the real Veff); elements outside the grid read as zero */
bool make_map(CUtensorMap *map, const void *base, const Layout &L, int words, int ncol, int box_x, int box_y, int promo)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[4] = {(cuuint64_t)L.Nx * words, (cuuint64_t)L.Ny, (cuuint64_t)L.Nz, (cuuint64_t)(ncol > 0 ? ncol : 1)};
    cuuint64_t strides[3] = {(cuuint64_t)L.Nx * 8 * words, (cuuint64_t)L.plane * 8 * words, (cuuint64_t)L.ld * 8 * words};
    cuuint32_t box[4] = {(cuuint32_t)box_x, (cuuint32_t)box_y, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void *>(base), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

/* orthogonal cell, FD radius 6, dense layout, at least one full 16 x 32 tile per plane, and -- for a periodic y
 * face -- no 6-row halo box straddling the face (Ny mod 32 is 0 or >= 6); a real Veff row must be 16-byte
 * aligned for its tensor map (Nx even).  Everything else goes through the z-march kernel. */
bool stream_kpt_supported(const chefsi_ctx *ctx)
{
    const chefsi_grid_t &g = ctx->grid;
    if (ctx->force_general) return false;
    if (g.cell_typ != 0 || g.FDn != R) return false;
    if (g.Nx % 2 != 0) return false;
    if (g.Nx < TXC || g.Ny < TY || g.Nz < 2 * R) return false;
    if (g.BCy == 0 && g.Ny % TY != 0 && g.Ny % TY < R) return false;
    return true;
}

int launch_stencil_stream_kpt(chefsi_ctx *ctx, const StepArgs &a)
{
    if (a.ncol <= 0) return 0;
    const chefsi_grid_t &g = ctx->grid;
    const Layout &L = ctx->lay;
    KptDesc d;
    d.Nx = g.Nx; d.Ny = g.Ny; d.Nz = g.Nz;
    d.bc[0] = g.BCx; d.bc[1] = g.BCy; d.bc[2] = g.BCz;
    d.ntx = (g.Nx + TXC - 1) / TXC;
    d.nty = (g.Ny + TY - 1) / TY;
    /* s1 (and c) applied through the weights: out = (s1 H') x - s2 xprev */
    d.coef0 = a.s1 * (ctx->desc.coef0 + a.c);
    for (int r = 0; r <= R; r++) { d.wx[r] = a.s1 * ctx->desc.wx[r]; d.wy[r] = a.s1 * ctx->desc.wy[r]; d.wz[r] = a.s1 * ctx->desc.wz[r]; }
    /* StencilDesc::ph_* index (oz+1)*9 + (oy+1)*3 + (ox+1), o = -1: value taken from across the low face */
    const int qm[3] = {12, 10, 4}, qp[3] = {14, 16, 22};
    for (int ax = 0; ax < 3; ax++) {
        d.phm_re[ax] = ctx->desc.ph_re[qm[ax]]; d.phm_im[ax] = ctx->desc.ph_im[qm[ax]];
        d.php_re[ax] = ctx->desc.ph_re[qp[ax]]; d.php_im[ax] = ctx->desc.ph_im[qp[ax]];
    }
    const long long nitems = (long long)a.ncol * d.ntx * d.nty;
    if (nitems > 0x7fffffffLL) { chefsi_fail(ctx, "k-point stream kernel: too many work items"); return -1; }

    KptMaps m;
    const void *xp = a.xprev ? a.xprev : a.x; /* never dereferenced when s2 == 0 */
    const int promo = ctx->tma_l2promo;
    if (!make_map(&m.y_full, a.x, L, 2, a.ncol, Cfg::YP, Cfg::YROWS, promo) || !make_map(&m.y_top, a.x, L, 2, a.ncol, Cfg::YP, HT, promo) ||
        !make_map(&m.y_body, a.x, L, 2, a.ncol, Cfg::YP, TY, promo) || !make_map(&m.y_bot, a.x, L, 2, a.ncol, Cfg::YP, R, promo) ||
        !make_map(&m.y_strip, a.x, L, 2, a.ncol, Cfg::SP, TY, promo) || !make_map(&m.xprev, xp, L, 2, a.ncol, Cfg::XP, TY, promo) ||
        !make_map(&m.veff, ctx->d_veff, L, 1, 1, Cfg::VP, TY, promo)) {
        chefsi_fail(ctx, "cuTensorMapEncodeTiled failed (k-point)");
        return -1;
    }
    /* with / without the Veff tile and the xprev tile: compile-time, so the plane step carries no selects for them */
    const bool hv = a.veff != nullptr, hx = a.s2 != 0.0;
    auto kern = hv ? (hx ? stream_kpt_kernel<true, true> : stream_kpt_kernel<true, false>)
                   : (hx ? stream_kpt_kernel<false, true> : stream_kpt_kernel<false, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) { chefsi_fail(ctx, "cudaFuncSetAttribute(k-point stream): %s", cudaGetErrorString(e)); return -1; }
    const int grid = (int)((nitems < ctx->num_sms) ? nitems : ctx->num_sms);
    unsigned int *counter = nullptr;
    unsigned int base = 0;
    if (ctx->stream_gridsync && nitems > grid) { /* round barrier between the producers, as in the real kernel */
        if (!ctx->d_sync) {
            if (cudaMalloc((void **)&ctx->d_sync, 256) != cudaSuccess || cudaMemset(ctx->d_sync, 0, 256) != cudaSuccess) {
                chefsi_fail(ctx, "k-point stream kernel: cannot allocate the round-barrier counter");
                return -1;
            }
        }
        counter = ctx->d_sync;
        base = ctx->sync_arrivals;
        ctx->sync_arrivals += (unsigned int)nitems;
    }
    int nit = (int)nitems;
    if (counter) {
        void *args[] = {(void *)&m, (void *)&d, (void *)&a, (void *)&nit, (void *)&counter, (void *)&base};
        e = cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(Cfg::THREADS), args, Cfg::SMEM, ctx->stream);
        if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
            cudaGetLastError();
            ctx->sync_arrivals = base;
            ctx->stream_gridsync = 0;
            counter = nullptr;
            kern<<<grid, Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(m, d, a, nit, counter, base);
        } else if (e != cudaSuccess) { chefsi_fail(ctx, "k-point stream kernel cooperative launch: %s", cudaGetErrorString(e)); return -1; }
    } else {
        kern<<<grid, Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(m, d, a, nit, counter, base);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "k-point stream kernel launch: %s", cudaGetErrorString(e)); return -1; }
    return 1;
}
