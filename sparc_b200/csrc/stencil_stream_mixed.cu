/*
 * stencil_stream_mixed.cu -- streaming fused Chebyshev step for NON-ORTHOGONAL cells (cell_typ 11..17), real data,
 * on the skeleton of stencil_stream_dense.cu (sm_100a).
 *
 *   out = s1 * ( (-1/2 Lap + Veff + c) x ) - s2 * xprev ,
 *   Lap = T11 dxx + T22 dyy + T33 dzz + 2 T12 dx dy + 2 T13 dx dz + 2 T23 dy dz          (FP64, FD radius 6)
 *
 * It replaces Lap_plus_diag_vec_mult_nonorth (lapVecRoutines.c:940-1331): the 26-region haloed copy, the
 * first-derivative passes Calc_DX / Calc_DX1_DX2 (gradVecRoutines.c:318, lapVecRoutines.c:1429) and
 * stencil_4comp / stencil_5comp (:1481, :1540), plus the recurrence passes of ChebyshevFiltering
 * (eigenSolver.c:764-768,787-794).  The reference evaluates every mixed term as a TWO-STAGE product of
 * first-derivative stencils (never as a 13 x 13 cross stencil); so does this kernel -- the same discrete
 * operator, only the order in which the two commuting one-dimensional differences are applied is chosen per
 * axis pair so that the streaming direction z is always the OUTER one:
 *
 *   x-y  (in-plane)   D = sum_r b[r] (f(j+r) - f(j-r)) for the tile rows on the x-extended columns, into a per-warp
 *                     shared tile (a warp owns 4 complete rows, so only __syncwarp is needed); then
 *                     sum_p a[p] (D(i+p) - D(i-p)).  The tile's own points reuse the y-window values the star
 *                     stencil has loaded anyway; the 2 x 6 halo columns are done by the 6 edge lanes of each half warp.
 *   x-z, y-z          W = sum_p gx[p] (f(i+p) - f(i-p)) + sum_p gy[p] (f(j+p) - f(j-p)) of the arriving plane
 *                     (differences of the x- and y-window values that are in registers for the star stencil), then
 *                     the z difference sum_r cz[r] (W(k+r) - W(k-r)) through the register z-queue.  The reference
 *                     differentiates along z first on an x- or y-extended box; difference operators along
 *                     different axes commute exactly (also across Dirichlet zero fill), so the results differ by
 *                     rounding only (tests: <= 1e-13 against the oracle).
 *
 * z never touches shared memory: a thread keeps ONE queue of 13 partial outputs for each of its 2 x 2 points
 * (planes p-6 .. p+6 around the arriving plane p); plane p scatters its star-z and W-z contributions into all 12
 * neighbours and its in-plane terms into the centre, the oldest entry is emitted and the queue shifts (register
 * moves: the kernel is FP64-pipe bound at ~96 FP64 instructions per point, the moves issue in its shadow).
 *
 * Halo handling as in stencil_stream_dense.cu (TMA out-of-bounds zero fill on Dirichlet faces, split boxes on a
 * periodic y face, periodic-x strips merged into the tile's halo columns by the spare producer-group warps), except
 * that the strips cover the halo ROWS as well: the x-y term reads the tile's corners.
 */
#include <cuda.h>

#include <cmath>
#include <cstring>

#include "chefsi_internal.h"
#include "tma_ring.cuh"

namespace {

using namespace tma_ring;

constexpr int R = 6;        /* FD radius */
constexpr int kStages = 4;  /* shared memory ring depth (5 measured 2 % slower: profiles/r2_exp_mixed_kernel_variants.log) */
constexpr int HT = 8;       /* top halo rows held in the tile (6 used; 8 keep box starts 128-byte aligned) */
constexpr int SW = 10;      /* width of the periodic-x strips */
constexpr int Q = 2 * R + 1; /* depth of the z queue */

struct Cfg {
    static constexpr int TX = 32, TY = 32;
    static constexpr int YP = TX + 2 * R + 2;   /* 46: haloed tile pitch (odd number of 16-byte chunks) */
    static constexpr int YROWS = HT + TY + R;   /* 46 */
    static constexpr int XP = TX + 2;           /* xprev / Veff tile pitch */
    static constexpr int Y_BYTES = ((YP * YROWS * 8 + 127) / 128) * 128;
    static constexpr int S_BYTES = ((SW * YROWS * 8 + 127) / 128) * 128;
    static constexpr int X_BYTES = ((XP * TY * 8 + 127) / 128) * 128;
    static constexpr int OFF_L = Y_BYTES;
    static constexpr int OFF_R = OFF_L + S_BYTES;
    static constexpr int OFF_V = OFF_R + S_BYTES;
    static constexpr int OFF_X = OFF_V + X_BYTES;
    static constexpr int STAGE_BYTES = OFF_X + X_BYTES;
    static constexpr int CONSUMER_WARPS = 8;
    static constexpr int THREADS = (CONSUMER_WARPS + 4) * 32;
    static constexpr int PRODUCER_REGS = 40, CONSUMER_REGS = 232;
    static constexpr int DP = YP;                       /* pitch of the per-warp D tile (4 rows x 44 columns) */
    static constexpr int D_BYTES = 4 * DP * 8;          /* per warp */
    static constexpr size_t SMEM = (size_t)kStages * STAGE_BYTES + (size_t)CONSUMER_WARPS * D_BYTES + 3 * kStages * sizeof(unsigned long long);
    static_assert((YP / 2) % 2 == 1 && (XP / 2) % 2 == 1, "row pitches must be an odd number of 16-byte chunks");
    static_assert((YP * HT * 8) % 128 == 0 && (YP * TY * 8) % 128 == 0, "box starts must be 128-byte aligned");
    static_assert((SW * HT * 8) % 128 == 0 && (SW * TY * 8) % 128 == 0, "strip box starts must be 128-byte aligned");
    static_assert(D_BYTES % 16 == 0 && STAGE_BYTES % 128 == 0, "alignment");
};

struct MixDesc {
    int Nx, Ny, Nz;
    int bc[3];
    int ntx, nty;
    double coef0;                       /* s1 * (a (D2x[0] + D2y[0] + D2z[0]) + c) */
    double wx[R + 1], wy[R + 1], wz[R + 1]; /* s1 * a * D2_*                    */
    double axy[R + 1], bxy[R + 1];      /* x-y term: outer x weights (s1 * a folded in), inner y weights */
    double cz[R + 1];                   /* z weights of the W term */
    double gxw[R + 1], gyw[R + 1];      /* in-plane weights of W (s1 * a and the z-weight ratio folded in) */
};

struct MixMaps {
    CUtensorMap y_full, y_top, y_body, y_bot;   /* YP x {YROWS, HT, TY, R} */
    CUtensorMap s_full, s_top, s_body, s_bot;   /* SW x {YROWS, HT, TY, R}: periodic-x strips */
    CUtensorMap xprev, veff;                    /* XP x TY */
};

/* ---- one plane step of a consumer thread (2 x 2 patch: x pair xp of rows r0, r0 + 1) ------------------------- */
template <bool XY, bool ZW>
__device__ __forceinline__ void consume_plane(const MixDesc &d, const StepArgs &a, const unsigned char *stage, double *dtile,
                                              const int p, const bool act0, const bool act1, const int xp, const int r0,
                                              const int lane, double *__restrict__ out_row, const size_t plane_elems,
                                              double (&acc)[Q][4], const bool plane_is_zero)
{
    const int Nz = d.Nz;
    const bool interior = (p >= 0) && (p < Nz);
    const int o = p - R;
    const bool emit = o >= 0 && o < Nz;
    const double *ytile = reinterpret_cast<const double *>(stage);
    const double *vtile = reinterpret_cast<const double *>(stage + Cfg::OFF_V);
    const double *xtile = reinterpret_cast<const double *>(stage + Cfg::OFF_X);

    if (!plane_is_zero) {
        double v[4], W[4] = {0, 0, 0, 0};
        const double *cp = ytile + (r0 + HT) * Cfg::YP + 2 * xp + R; /* centre chunk of row r0 */
        if (interior || ZW) {
            /* x windows of the two rows (the periodic-x strips were merged into the halo columns) */
            double xr[2][14];
#pragma unroll
            for (int t = 0; t < 7; t++) {
                const double2 w0 = *reinterpret_cast<const double2 *>(cp + 2 * (t - 3));
                const double2 w1 = *reinterpret_cast<const double2 *>(cp + 2 * (t - 3) + Cfg::YP);
                xr[0][2 * t] = w0.x; xr[0][2 * t + 1] = w0.y;
                xr[1][2 * t] = w1.x; xr[1][2 * t + 1] = w1.y;
            }
#pragma unroll
            for (int i = 0; i < 4; i++) v[i] = xr[i >> 1][R + (i & 1)];
            /* y window: up[k] = row r0-k, dn[k] = row r0+1+k */
            double2 up[R + 1], dn[R + 1];
            up[0] = make_double2(v[0], v[1]);
            dn[0] = make_double2(v[2], v[3]);
#pragma unroll
            for (int k = 1; k <= R; k++) {
                up[k] = *reinterpret_cast<const double2 *>(cp - k * Cfg::YP);
                dn[k] = *reinterpret_cast<const double2 *>(cp + (1 + k) * Cfg::YP);
            }
            double inpl[4] = {0, 0, 0, 0}; /* in-plane terms of plane p itself */
            double dy[4] = {0, 0, 0, 0};   /* inner y derivative of the x-y term */
            if (interior) {
                double ve[4] = {0, 0, 0, 0};
                if (a.veff) {
                    const double2 w0 = *reinterpret_cast<const double2 *>(vtile + r0 * Cfg::XP + 2 * xp);
                    const double2 w1 = *reinterpret_cast<const double2 *>(vtile + (r0 + 1) * Cfg::XP + 2 * xp);
                    ve[0] = w0.x; ve[1] = w0.y; ve[2] = w1.x; ve[3] = w1.y;
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const double diag = a.veff ? fma(a.s1, ve[i], d.coef0) : d.coef0;
                    inpl[i] = diag * v[i];
                }
            }
            /* x direction: sums for the star, differences for W */
#pragma unroll
            for (int r = 1; r <= R; r++)
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int row = i >> 1, j = i & 1;
                    const double lo = xr[row][R + j - r], hi = xr[row][R + j + r];
                    if (interior) inpl[i] = fma(d.wx[r], lo + hi, inpl[i]);
                    if (ZW) W[i] = fma(d.gxw[r], hi - lo, W[i]);
                }
            /* y direction: row r0 pairs up[k] with (k == 1 ? own row 1 : dn[k-1]); row r0+1 pairs (k == 1 ? own row 0 : up[k-1]) with dn[k] */
#pragma unroll
            for (int k = 1; k <= R; k++) {
                const double lo0 = up[k].x, lo1 = up[k].y, hi0 = dn[k - 1].x, hi1 = dn[k - 1].y;     /* row r0   */
                const double lo2 = up[k - 1].x, lo3 = up[k - 1].y, hi2 = dn[k].x, hi3 = dn[k].y;     /* row r0+1 */
                if (interior) {
                    inpl[0] = fma(d.wy[k], lo0 + hi0, inpl[0]); inpl[1] = fma(d.wy[k], lo1 + hi1, inpl[1]);
                    inpl[2] = fma(d.wy[k], lo2 + hi2, inpl[2]); inpl[3] = fma(d.wy[k], lo3 + hi3, inpl[3]);
                }
                if (ZW || (XY && interior)) {
                    const double d0 = hi0 - lo0, d1 = hi1 - lo1, d2 = hi2 - lo2, d3 = hi3 - lo3;
                    if (ZW) {
                        W[0] = fma(d.gyw[k], d0, W[0]); W[1] = fma(d.gyw[k], d1, W[1]);
                        W[2] = fma(d.gyw[k], d2, W[2]); W[3] = fma(d.gyw[k], d3, W[3]);
                    }
                    if (XY && interior) {
                        dy[0] = fma(d.bxy[k], d0, dy[0]); dy[1] = fma(d.bxy[k], d1, dy[1]);
                        dy[2] = fma(d.bxy[k], d2, dy[2]); dy[3] = fma(d.bxy[k], d3, dy[3]);
                    }
                }
            }
            if (XY && interior) {
                /* ---- x-y term: D on the warp's 4 rows x 44 columns, then the x difference of D ---- */
                const int rl = 2 * (lane >> 4);                       /* local row of this thread's first row (0 or 2) */
                double *drow = dtile + rl * Cfg::DP;
                *reinterpret_cast<double2 *>(drow + R + 2 * xp) = make_double2(dy[0], dy[1]);
                *reinterpret_cast<double2 *>(drow + Cfg::DP + R + 2 * xp) = make_double2(dy[2], dy[3]);
                if (xp < 3 || xp >= 13) {
                    /* halo column pair hx of the haloed tile: columns 2 hx, 2 hx + 1 (left 0..5, right 38..43) */
                    const int hx = (xp < 3) ? xp : xp + 6;
                    const double *hp = ytile + (r0 + HT) * Cfg::YP + 2 * hx;
                    double e0 = 0, e1 = 0, e2 = 0, e3 = 0;
                    double2 hu[R + 1], hd[R + 1];
#pragma unroll
                    for (int k = 0; k <= R; k++) {
                        hu[k] = *reinterpret_cast<const double2 *>(hp - k * Cfg::YP);
                        hd[k] = *reinterpret_cast<const double2 *>(hp + (1 + k) * Cfg::YP);
                    }
#pragma unroll
                    for (int k = 1; k <= R; k++) {
                        e0 = fma(d.bxy[k], hd[k - 1].x - hu[k].x, e0); e1 = fma(d.bxy[k], hd[k - 1].y - hu[k].y, e1);
                        e2 = fma(d.bxy[k], hd[k].x - hu[k - 1].x, e2); e3 = fma(d.bxy[k], hd[k].y - hu[k - 1].y, e3);
                    }
                    *reinterpret_cast<double2 *>(drow + 2 * hx) = make_double2(e0, e1);
                    *reinterpret_cast<double2 *>(drow + Cfg::DP + 2 * hx) = make_double2(e2, e3);
                }
                __syncwarp();
                double dr[2][14];
#pragma unroll
                for (int t = 0; t < 7; t++) {
                    const double2 w0 = *reinterpret_cast<const double2 *>(drow + 2 * xp + 2 * t);
                    const double2 w1 = *reinterpret_cast<const double2 *>(drow + Cfg::DP + 2 * xp + 2 * t);
                    dr[0][2 * t] = w0.x; dr[0][2 * t + 1] = w0.y;
                    dr[1][2 * t] = w1.x; dr[1][2 * t + 1] = w1.y;
                }
                __syncwarp(); /* the tile is rewritten by the next plane */
#pragma unroll
                for (int r = 1; r <= R; r++)
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int row = i >> 1, j = i & 1;
                        inpl[i] = fma(d.axy[r], dr[row][R + j + r] - dr[row][R + j - r], inpl[i]);
                    }
            }
            if (interior) {
#pragma unroll
                for (int i = 0; i < 4; i++) acc[R][i] += inpl[i];
            }
        } else {
            const double2 w0 = *reinterpret_cast<const double2 *>(cp);
            const double2 w1 = *reinterpret_cast<const double2 *>(cp + Cfg::YP);
            v[0] = w0.x; v[1] = w0.y; v[2] = w1.x; v[3] = w1.y;
        }
        /* scatter the z terms of plane p into the 12 planes around it:
           out(o) has + wz[r] (f(o+r) + f(o-r)) + cz[r] (W(o+r) - W(o-r)) */
#pragma unroll
        for (int r = 1; r <= R; r++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                double lo = fma(d.wz[r], v[i], acc[R - r][i]); /* plane p - r: p is its +r neighbour */
                double hi = fma(d.wz[r], v[i], acc[R + r][i]); /* plane p + r: p is its -r neighbour */
                if (ZW) {
                    lo = fma(d.cz[r], W[i], lo);
                    hi = fma(-d.cz[r], W[i], hi);
                }
                acc[R - r][i] = lo;
                acc[R + r][i] = hi;
            }
    }
    if (emit && act1) {
        double res[4];
        if (a.s2 != 0.0) {
            const double2 w0 = *reinterpret_cast<const double2 *>(xtile + r0 * Cfg::XP + 2 * xp);
            const double2 w1 = *reinterpret_cast<const double2 *>(xtile + (r0 + 1) * Cfg::XP + 2 * xp);
            res[0] = fma(-a.s2, w0.x, acc[0][0]);
            res[1] = fma(-a.s2, w0.y, acc[0][1]);
            res[2] = fma(-a.s2, w1.x, acc[0][2]);
            res[3] = fma(-a.s2, w1.y, acc[0][3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) res[i] = acc[0][i];
        }
        double *dst = out_row + (size_t)o * plane_elems;
        if (act0) stg128(dst, res[0], res[1]);
        stg128(dst + d.Nx, res[2], res[3]);
    }
    /* the queue moves on by one plane */
#pragma unroll
    for (int u = 0; u < Q - 1; u++)
#pragma unroll
        for (int i = 0; i < 4; i++) acc[u][i] = acc[u + 1][i];
#pragma unroll
    for (int i = 0; i < 4; i++) acc[Q - 1][i] = 0.0;
}

template <bool XY, bool ZW>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
stream_mixed_kernel(const __grid_constant__ MixMaps maps, const __grid_constant__ MixDesc d, const StepArgs a, const int nitems,
                    unsigned int *__restrict__ sync_counter, const unsigned int sync_base)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *ring = smem_raw;
    double *dtiles = reinterpret_cast<double *>(ring + (size_t)kStages * Cfg::STAGE_BYTES);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)kStages * Cfg::STAGE_BYTES + (size_t)Cfg::CONSUMER_WARPS * Cfg::D_BYTES);
    uint64_t *empty = full + kStages;
    uint64_t *landed = empty + kStages;

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
    const size_t plane_elems = (size_t)Nx * Ny;
    uint32_t it = 0;

    if (warp >= Cfg::CONSUMER_WARPS) {
        /* ================= producer warpgroup ================= */
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::PRODUCER_REGS));
        if (warp == Cfg::CONSUMER_WARPS && lane == 0) {
            unsigned int round = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, round++) {
                const int tile = item % (d.ntx * d.nty), n = item / (d.ntx * d.nty);
                const int x0 = tile_origin(tile % d.ntx, Cfg::TX, Nx), y0 = tile_origin(tile / d.ntx, Cfg::TY, Ny);
                const bool wrap_top = yper && (y0 - R < 0), wrap_bot = yper && (y0 + Cfg::TY + R > Ny);
                const bool split_y = wrap_top || wrap_bot;
                const int ytop = wrap_top ? y0 - HT + Ny : y0 - HT;
                const int ybot = wrap_bot ? y0 + Cfg::TY - Ny : y0 + Cfg::TY;
                const bool need_l = xper && (x0 - R < 0), need_r = xper && (x0 + Cfg::TX + R > Nx);
                const uint32_t sbytes = (uint32_t)(SW * Cfg::YROWS * 8);
                const uint32_t ybytes = (uint32_t)(Cfg::YP * Cfg::YROWS * 8) + (need_l ? sbytes : 0u) + (need_r ? sbytes : 0u);
                if (sync_counter) { /* round barrier between the producers of all CTAs: see stencil_stream_dense.cu */
                    const unsigned int in_round = (unsigned int)min((long long)gridDim.x, (long long)nitems - (long long)round * gridDim.x);
                    const unsigned int done_before = round * gridDim.x;
                    __threadfence();
                    atomicAdd(sync_counter, 1u);
                    const unsigned int target = sync_base + done_before + in_round;
                    unsigned int spins = 0;
                    while ((int)(*(volatile unsigned int *)sync_counter - target) < 0 && ++spins < (1u << 22)) __nanosleep(64);
                    if (spins >= (1u << 22)) atomicAdd(sync_counter + 1, 1u);
                }
                for (int p = -R; p < Nz + R; p++) {
                    int kz = p;
                    const bool interior = (p >= 0 && p < Nz);
                    if (p < 0) kz += Nz; else if (p >= Nz) kz -= Nz;
                    const int o = p - R;
                    const bool need_y = interior || zper;
                    const bool need_v = interior && a.veff != nullptr;
                    const bool need_x = (o >= 0 && o < Nz) && a.s2 != 0.0;
                    if (!need_y && !need_x) continue;
                    const int s = it % kStages;
                    unsigned char *stage = ring + (size_t)s * Cfg::STAGE_BYTES;
                    mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
                    mbar_expect_tx(&landed[s], (need_y ? ybytes : 0u) + (uint32_t)((need_v ? Cfg::XP * Cfg::TY * 8 : 0) +
                                                                                    (need_x ? Cfg::XP * Cfg::TY * 8 : 0)));
                    if (need_y) {
                        if (!split_y) {
                            tma_load_4d(stage, &maps.y_full, x0 - R, y0 - HT, kz, n, &landed[s]);
                            if (need_l) tma_load_4d(stage + Cfg::OFF_L, &maps.s_full, Nx - SW, y0 - HT, kz, n, &landed[s]);
                            if (need_r) tma_load_4d(stage + Cfg::OFF_R, &maps.s_full, 0, y0 - HT, kz, n, &landed[s]);
                        } else {
                            tma_load_4d(stage, &maps.y_top, x0 - R, ytop, kz, n, &landed[s]);
                            tma_load_4d(stage + Cfg::YP * HT * 8, &maps.y_body, x0 - R, y0, kz, n, &landed[s]);
                            tma_load_4d(stage + Cfg::YP * (HT + Cfg::TY) * 8, &maps.y_bot, x0 - R, ybot, kz, n, &landed[s]);
                            if (need_l) {
                                tma_load_4d(stage + Cfg::OFF_L, &maps.s_top, Nx - SW, ytop, kz, n, &landed[s]);
                                tma_load_4d(stage + Cfg::OFF_L + SW * HT * 8, &maps.s_body, Nx - SW, y0, kz, n, &landed[s]);
                                tma_load_4d(stage + Cfg::OFF_L + SW * (HT + Cfg::TY) * 8, &maps.s_bot, Nx - SW, ybot, kz, n, &landed[s]);
                            }
                            if (need_r) {
                                tma_load_4d(stage + Cfg::OFF_R, &maps.s_top, 0, ytop, kz, n, &landed[s]);
                                tma_load_4d(stage + Cfg::OFF_R + SW * HT * 8, &maps.s_body, 0, y0, kz, n, &landed[s]);
                                tma_load_4d(stage + Cfg::OFF_R + SW * (HT + Cfg::TY) * 8, &maps.s_bot, 0, ybot, kz, n, &landed[s]);
                            }
                        }
                    }
                    if (need_v) tma_load_4d(stage + Cfg::OFF_V, &maps.veff, x0, y0, p, 0, &landed[s]);
                    if (need_x) tma_load_4d(stage + Cfg::OFF_X, &maps.xprev, x0, y0, o, n, &landed[s]);
                    it++;
                }
            }
        } else if (warp > Cfg::CONSUMER_WARPS) {
            /* ---- merge warps: copy the periodic-x strips of a landed stage into the halo columns of its tile (all
               YROWS rows: the x-y term reads the corners), then hand the stage to the consumers ---- */
            const int ft = (int)threadIdx.x - (Cfg::CONSUMER_WARPS + 1) * 32; /* 0 .. 95 */
            uint32_t itf = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int tile = item % (d.ntx * d.nty);
                const int x0 = tile_origin(tile % d.ntx, Cfg::TX, Nx);
                const bool need_l = xper && (x0 - R < 0), need_r = xper && (x0 + Cfg::TX + R > Nx);
                for (int p = -R; p < Nz + R; p++) {
                    const bool interior = (p >= 0 && p < Nz);
                    const int o = p - R;
                    const bool need_y = interior || zper;
                    const bool need_x = (o >= 0 && o < Nz) && a.s2 != 0.0;
                    if (!need_y && !need_x) continue;
                    const int s = itf % kStages;
                    unsigned char *stage = ring + (size_t)s * Cfg::STAGE_BYTES;
                    mbar_wait(&landed[s], (itf / kStages) & 1);
                    if (need_y && (need_l || need_r)) {
                        double *tile_d = reinterpret_cast<double *>(stage);
                        const double *sl = reinterpret_cast<const double *>(stage + Cfg::OFF_L);
                        const double *sr = reinterpret_cast<const double *>(stage + Cfg::OFF_R);
                        if (need_l)
                            for (int c = ft; c < 3 * Cfg::YROWS; c += 96) {
                                const int row = c / 3, j = c % 3, gi = x0 - R + 2 * j;
                                if (gi < 0)
                                    *reinterpret_cast<double2 *>(tile_d + row * Cfg::YP + 2 * j) =
                                        *reinterpret_cast<const double2 *>(sl + row * SW + gi + SW);
                            }
                        if (need_r)
                            for (int c = ft; c < 4 * Cfg::YROWS; c += 96) {
                                const int row = c / 4, j = c % 4, col = Nx + 2 * j - (x0 - R);
                                if (col + 1 < Cfg::YP)
                                    *reinterpret_cast<double2 *>(tile_d + row * Cfg::YP + col) =
                                        *reinterpret_cast<const double2 *>(sr + row * SW + 2 * j);
                            }
                    }
                    asm volatile("bar.sync 2, 96;" ::: "memory");
                    if (ft == 0) mbar_arrive(&full[s]);
                    itf++;
                }
            }
        }
    } else {
        /* ================= consumer warps ================= */
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::CONSUMER_REGS));
        const int qx = lane & 15;
        const int ry = warp * 4 + 2 * (lane >> 4);
        double *dtile = dtiles + warp * (Cfg::D_BYTES / 8);
        double acc[Q][4];
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int tile = item % (d.ntx * d.nty), n = item / (d.ntx * d.nty);
            const int tx = tile % d.ntx, ty = tile / d.ntx;
            const int x0 = tile_origin(tx, Cfg::TX, Nx), y0 = tile_origin(ty, Cfg::TY, Ny);
            const int gx = x0 + 2 * qx, gy = y0 + ry;
            /* a shifted last tile overlaps its neighbour: only the not yet covered points are stored */
            const bool act0 = (gx >= tx * Cfg::TX) && (gy >= ty * Cfg::TY);
            const bool act1 = (gx >= tx * Cfg::TX) && (gy + 1 >= ty * Cfg::TY);
            double *out_row = reinterpret_cast<double *>(a.out) + (size_t)n * a.ld + (size_t)gy * Nx + gx;
#pragma unroll
            for (int u = 0; u < Q; u++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[u][j] = 0.0;
#pragma unroll 1
            for (int p = -R; p < Nz + R; p++) {
                const bool zplane = !zper && (p < 0 || p >= Nz);   /* Dirichlet z: the plane is zero */
                const bool use_stage = !zplane || (p - R >= 0 && p - R < Nz && a.s2 != 0.0);
                const unsigned char *stage = ring;
                int s = 0;
                if (use_stage) {
                    s = it % kStages;
                    stage = ring + (size_t)s * Cfg::STAGE_BYTES;
                    mbar_wait(&full[s], (it / kStages) & 1);
                }
                consume_plane<XY, ZW>(d, a, stage, dtile, p, act0, act1, qx, ry, lane, out_row, plane_elems, acc, zplane);
                if (use_stage) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s]);
                    it++;
                }
            }
        }
    }
}

/* ---- host side ---------------------------------------------------------------------------- */
bool make_map(CUtensorMap *map, const void *base, const Layout &L, int ncol, int box_x, int box_y, int promo)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[4] = {(cuuint64_t)L.Nx, (cuuint64_t)L.Ny, (cuuint64_t)L.Nz, (cuuint64_t)(ncol > 0 ? ncol : 1)};
    cuuint64_t strides[3] = {(cuuint64_t)L.Nx * 8, (cuuint64_t)L.plane * 8, (cuuint64_t)L.ld * 8};
    cuuint32_t box[4] = {(cuuint32_t)box_x, (cuuint32_t)box_y, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void *>(base), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

/* Sort the reference's mixed components (lapVecRoutines.c:1210-1297, flattened in chefsi_set_grid) into the three
 * axis pairs.  Each component is sum_p wm[p] (D(+p ext) - D(-p ext)) with D = sum_r c1[r] d_{ax1,r} f (+ c2[r] d_{ax2,r} f),
 * i.e. one or two tensor products of first-derivative stencils.  Returns false when the tables do not have the
 * structure the kernel folds (both z-outer terms must share their z weights up to a factor: they are all
 * multiples of the same FD weights w1, initialization.c:2164-2177). */
bool build_desc(const chefsi_ctx *ctx, const StepArgs &a, MixDesc &d, bool &has_xy, bool &has_zw)
{
    const chefsi_grid_t &g = ctx->grid;
    const StencilDesc &sd = ctx->desc;
    memset(&d, 0, sizeof(d));
    d.Nx = g.Nx; d.Ny = g.Ny; d.Nz = g.Nz;
    d.bc[0] = g.BCx; d.bc[1] = g.BCy; d.bc[2] = g.BCz;
    d.ntx = (g.Nx + Cfg::TX - 1) / Cfg::TX;
    d.nty = (g.Ny + Cfg::TY - 1) / Cfg::TY;
    d.coef0 = a.s1 * (sd.coef0 + a.c);
    for (int r = 0; r <= R; r++) { d.wx[r] = a.s1 * sd.wx[r]; d.wy[r] = a.s1 * sd.wy[r]; d.wz[r] = a.s1 * sd.wz[r]; }
    has_xy = has_zw = false;
    bool have_gx = false, have_gy = false;
    double czx[R + 1] = {0}, czy[R + 1] = {0}, gx[R + 1] = {0}, gy[R + 1] = {0};
    for (int q = 0; q < sd.nmix; q++) {
        const MixedComp &mc = sd.mix[q];
        for (int t = 0; t < 2; t++) {
            const int ax = t ? mc.ax2 : mc.ax1;
            const double *ci = t ? mc.c2 : mc.c1;
            if (ax < 0) continue;
            const int e = mc.ext;
            if (e == ax) return false;
            const int lo = e < ax ? e : ax, hi = e < ax ? ax : e;
            const double *w_lo = (e == lo) ? mc.wm : ci, *w_hi = (e == hi) ? mc.wm : ci; /* weights along the lower / higher axis */
            if (lo == 0 && hi == 1) {
                if (has_xy) return false;
                has_xy = true;
                for (int r = 0; r <= R; r++) { d.axy[r] = a.s1 * w_lo[r]; d.bxy[r] = w_hi[r]; }
            } else if (lo == 0 && hi == 2) {
                if (have_gx) return false;
                have_gx = true;
                for (int r = 0; r <= R; r++) { gx[r] = a.s1 * w_lo[r]; czx[r] = w_hi[r]; }
            } else {
                if (have_gy) return false;
                have_gy = true;
                for (int r = 0; r <= R; r++) { gy[r] = a.s1 * w_lo[r]; czy[r] = w_hi[r]; }
            }
        }
    }
    has_zw = have_gx || have_gy;
    if (have_gx && have_gy) {
        /* common z weights: czy = kappa * czx */
        int r0 = 1;
        while (r0 <= R && czx[r0] == 0.0) r0++;
        if (r0 > R) return false;
        const double kappa = czy[r0] / czx[r0];
        for (int r = 1; r <= R; r++)
            if (fabs(czy[r] - kappa * czx[r]) > 1e-13 * (fabs(czy[r]) + fabs(kappa * czx[r]) + 1e-300)) return false;
        for (int r = 0; r <= R; r++) { d.cz[r] = czx[r]; d.gxw[r] = gx[r]; d.gyw[r] = kappa * gy[r]; }
    } else if (have_gx) {
        for (int r = 0; r <= R; r++) { d.cz[r] = czx[r]; d.gxw[r] = gx[r]; }
    } else if (have_gy) {
        for (int r = 0; r <= R; r++) { d.cz[r] = czy[r]; d.gyw[r] = gy[r]; }
    }
    return has_xy || has_zw;
}

template <bool XY, bool ZW>
int launch_cfg(chefsi_ctx *ctx, const StepArgs &a, const MixDesc &d)
{
    const Layout &L = ctx->lay;
    const long long nitems = (long long)a.ncol * d.ntx * d.nty;
    if (nitems > 0x7fffffffLL) { chefsi_fail(ctx, "mixed stream kernel: too many work items"); return -1; }
    MixMaps m;
    const void *xp = a.xprev ? a.xprev : a.x; /* never dereferenced when s2 == 0 */
    const int promo = ctx->tma_l2promo;
    if (!make_map(&m.y_full, a.x, L, a.ncol, Cfg::YP, Cfg::YROWS, promo) || !make_map(&m.y_top, a.x, L, a.ncol, Cfg::YP, HT, promo) ||
        !make_map(&m.y_body, a.x, L, a.ncol, Cfg::YP, Cfg::TY, promo) || !make_map(&m.y_bot, a.x, L, a.ncol, Cfg::YP, R, promo) ||
        !make_map(&m.s_full, a.x, L, a.ncol, SW, Cfg::YROWS, promo) || !make_map(&m.s_top, a.x, L, a.ncol, SW, HT, promo) ||
        !make_map(&m.s_body, a.x, L, a.ncol, SW, Cfg::TY, promo) || !make_map(&m.s_bot, a.x, L, a.ncol, SW, R, promo) ||
        !make_map(&m.xprev, xp, L, a.ncol, Cfg::XP, Cfg::TY, promo) || !make_map(&m.veff, ctx->d_veff, L, 1, Cfg::XP, Cfg::TY, promo)) {
        chefsi_fail(ctx, "cuTensorMapEncodeTiled failed");
        return -1;
    }
    auto kern = stream_mixed_kernel<XY, ZW>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) { chefsi_fail(ctx, "cudaFuncSetAttribute(mixed stream): %s", cudaGetErrorString(e)); return -1; }
    const int grid = (int)((nitems < ctx->num_sms) ? nitems : ctx->num_sms);
    unsigned int *counter = nullptr;
    unsigned int base = 0;
    if (ctx->stream_gridsync && nitems > grid) {
        if (!ctx->d_sync) {
            if (cudaMalloc((void **)&ctx->d_sync, 256) != cudaSuccess || cudaMemset(ctx->d_sync, 0, 256) != cudaSuccess) {
                chefsi_fail(ctx, "mixed stream kernel: cannot allocate the round-barrier counter");
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
        } else if (e != cudaSuccess) { chefsi_fail(ctx, "mixed stream kernel cooperative launch: %s", cudaGetErrorString(e)); return -1; }
    } else {
        kern<<<grid, Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(m, d, a, nit, counter, base);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "mixed stream kernel launch: %s", cudaGetErrorString(e)); return -1; }
    return 1;
}

}  // namespace

/* Needs: non-orthogonal cell (11..17), FD radius 6, real data, Nx even (TMA strides are 16-byte multiples), at least
 * one full 32 x 32 tile per plane and no 6-row halo box straddling a periodic y face (Ny mod 32 is 0 or >= 6; the
 * strip boxes that carry the corners follow the same rows).  Everything else goes through the z-march kernel. */
bool stream_mixed_supported(const chefsi_ctx *ctx, bool is_complex)
{
    const chefsi_grid_t &g = ctx->grid;
    if (ctx->force_general || is_complex) return false;
    if (g.cell_typ < 11 || g.cell_typ > 17 || g.FDn != R) return false;
    if (g.Nx % 2 != 0) return false;
    if (g.Nx < Cfg::TX || g.Ny < Cfg::TY || g.Nz < 2 * R) return false;
    if (g.BCy == 0 && g.Ny % Cfg::TY != 0 && g.Ny % Cfg::TY < R) return false;
    return true;
}

/* returns -2 when the coefficient tables do not have the structure the kernel folds (the caller falls back) */
int launch_stencil_stream_mixed(chefsi_ctx *ctx, const StepArgs &a)
{
    if (a.ncol <= 0) return 0;
    MixDesc d;
    bool xy = false, zw = false;
    if (!build_desc(ctx, a, d, xy, zw)) return -2;
    if (xy && zw) return launch_cfg<true, true>(ctx, a, d);
    if (xy) return launch_cfg<true, false>(ctx, a, d);
    return launch_cfg<false, true>(ctx, a, d);
}
