/*
 * stencil_stream_dense.cu -- streaming fused Chebyshev step for orthogonal cells on the DENSE
 * (reference) column layout (sm_100a).
 *
 *   out = s1 * ( (-1/2 Lap + Veff + c) x ) - s2 * xprev          (FP64, radius-6 star stencil)
 *
 * It replaces, per Chebyshev degree, the reference's haloed copy + stencil_3axis_thread_radius6
 * (lapVecRoutines.c:185-227 called from :586) + the three scale/axpy/swap passes of ChebyshevFiltering
 * (eigenSolver.c:764-768,787-794).  The columns are stored exactly as the reference stores them (x fastest, no
 * halo pads), so the DRAM traffic of a step is the algorithmic 24 B per grid point plus halo misses and nothing
 * has to be packed, unpacked or patched around the step.
 *
 * 2.5-D streaming: one persistent CTA per SM; a work item is (column, 32 x 32 xy-tile); the CTA marches the whole
 * z extent.  A producer lane streams per z-plane the haloed (TX+12+2) x (8+TY+6) tile of the input, the Veff tile
 * and the xprev tile of the plane that completes 6 planes later by TMA into a 5-stage shared-memory ring with
 * full / empty mbarriers; the consumer warps keep z in registers (the last 6 input planes and 7 partial outputs
 * of their 2 x 2 points).  Halos without pads:
 *   - Dirichlet faces and tile parts outside the grid come from the TMA out-of-bounds zero fill;
 *   - a periodic y face splits the tile into up to three boxes (top halo rows / body / bottom halo rows)
 *     whose y coordinates are wrapped separately; they land in consecutive rows of the same shared tile
 *     (8 top rows instead of 6 keep every box start 128-byte aligned);
 *   - a periodic x face adds a 10-column strip box from the other side of the grid, which the three spare
 *     warps of the producer warpgroup copy into the zero-filled halo columns of the landed tile before they
 *     hand the stage to the consumers (so every x window is read with constant offsets).
 * Tiles never hang over the grid edge: the last tile of a row/column is shifted inwards (x0 = Nx - TX)
 * and the threads that would recompute its neighbour's points are masked.
 * The producers of all CTAs pass a round barrier (global counter) before each item, so neighbouring tiles of a
 * column are marched at the same z and their xy-halos hit in L2 (DRAM reads 13.9 -> 10.8 GB per launch).
 * See DESIGN.md section 4 for the measurements behind each choice.
 */
#include <cuda.h>

#include "chefsi_internal.h"
#include "tma_ring.cuh"

namespace {

using namespace tma_ring;

constexpr int R = 6;        /* FD radius this kernel is specialised for (FD_ORDER 12) */
constexpr int kStages = 5;  /* shared memory ring depth */
constexpr int HT = 8;       /* top halo rows held in the tile (6 used) */
constexpr int SW = 10;      /* width of the periodic-x strips: 5 x 16 B, an odd number of chunks */

template <int WX, int WY> struct TileCfg {
    static constexpr int TX = 16 * WX;           /* tile width  (points) */
    static constexpr int TY = 8 * WY;            /* tile height (points) */
    static constexpr int YP = TX + 2 * R + 2;    /* haloed tile pitch (doubles): odd number of 16 B chunks */
    static constexpr int YROWS = HT + TY + R;
    static constexpr int XP = TX + 2;            /* xprev / Veff tile pitch */
    static constexpr int Y_BYTES = ((YP * YROWS * 8 + 127) / 128) * 128;
    static constexpr int S_BYTES = ((SW * TY * 8 + 127) / 128) * 128;
    static constexpr int X_BYTES = ((XP * TY * 8 + 127) / 128) * 128;
    static constexpr int OFF_L = Y_BYTES;
    static constexpr int OFF_R = OFF_L + S_BYTES;
    static constexpr int OFF_V = OFF_R + S_BYTES;
    static constexpr int OFF_X = OFF_V + X_BYTES;
    static constexpr int STAGE_BYTES = OFF_X + X_BYTES;
    static constexpr int CONSUMER_WARPS = WX * WY;
    /* consumers + one producer WARPGROUP: setmaxnreg moves registers per warpgroup (4 warps), and the register
       file is per scheduler (16 K registers): with a 9th warp one scheduler hosts 3 warps and ptxas has to cap
       every thread at 168 registers; with 2 consumer warps + 1 producer warp per scheduler the producers give
       their registers back (168 -> 40) and the consumers grow to 232 */
    static constexpr int THREADS = (CONSUMER_WARPS + 4) * 32;
    static constexpr int PRODUCER_REGS = 40, CONSUMER_REGS = 232;
    static constexpr size_t SMEM = (size_t)kStages * STAGE_BYTES + 3 * kStages * sizeof(unsigned long long);
    static_assert((YP / 2) % 2 == 1 && (XP / 2) % 2 == 1, "row pitches must be an odd number of 16-byte chunks");
    static_assert((YP * HT * 8) % 128 == 0 && (YP * TY * 8) % 128 == 0, "box starts must be 128-byte aligned");
};

struct DenseDesc {
    int Nx, Ny, Nz;
    int bc[3];
    int ntx, nty;
    double coef0;
    double wx[R + 1], wy[R + 1], wz[R + 1];
};

struct DenseMaps {
    CUtensorMap y_full;   /* YP x YROWS : the whole haloed tile in one box        */
    CUtensorMap y_top;    /* YP x HT                                               */
    CUtensorMap y_body;   /* YP x TY                                               */
    CUtensorMap y_bot;    /* YP x R                                                */
    CUtensorMap y_strip;  /* SW x TY    : periodic-x strips                        */
    CUtensorMap xprev;    /* XP x TY                                               */
    CUtensorMap veff;     /* XP x TY                                               */
};

/* ---- one plane step of a consumer thread: 2 x 2 thread tile ------------------------------------
 * U = (p + 7) mod 7 (compile time): register-queue rotation by renaming.
 * A thread owns the x pair (2 xp, 2 xp + 1) of rows r0 and r0 + 1 (point index 2*row + j).  Per plane it
 * reads 2 x 7 chunks for the two x windows, 12 chunks for the rows r0-6 .. r0-1 and r0+2 .. r0+7 (each
 * halo row serves both output rows) and 2 + 2 chunks of Veff / xprev: 30 LDS.128 per 4 points.  A quarter
 * warp reads 8 consecutive chunks of one row: conflict free for any pitch.                            */
/* the in-plane part of an interior plane: v = the centre values, acc[U] = x + y terms + the z terms of the 6 planes
   behind (from the input queue) */
template <class Cfg, int U, bool HASV>
__device__ __forceinline__ void interior_plane22(const DenseDesc &d, const StepArgs &a, const double *cp, const double *vtile,
                                                 uint32_t voff, const double (&in)[7][4], double (&acc)[7][4], double (&v)[4])
{
    double xr[2][14];
#pragma unroll
    for (int t = 0; t < 7; t++) {
        /* the periodic-x strips were merged into the tile's halo columns: constant offsets */
        const double2 w0 = *reinterpret_cast<const double2 *>(cp + 2 * (t - 3));
        const double2 w1 = *reinterpret_cast<const double2 *>(cp + 2 * (t - 3) + Cfg::YP);
        xr[0][2 * t] = w0.x; xr[0][2 * t + 1] = w0.y;
        xr[1][2 * t] = w1.x; xr[1][2 * t + 1] = w1.y;
    }
    double ve[4] = {0, 0, 0, 0};
    if (HASV) {
        const double2 w0 = *reinterpret_cast<const double2 *>(vtile + voff); /* r0 * XP + 2 xp */
        const double2 w1 = *reinterpret_cast<const double2 *>(vtile + voff + Cfg::XP);
        ve[0] = w0.x; ve[1] = w0.y; ve[2] = w1.x; ve[3] = w1.y;
    }
    /* d.w*, d.coef0 carry the recurrence scale s1 (and the shift c), see launch_cfg: 40 FP64 instructions per point */
    double sx[4], sy[4], sz[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int row = i >> 1, j = i & 1;
        v[i] = xr[row][R + j];
        const double diag = HASV ? fma(a.s1, ve[i], d.coef0) : d.coef0;
        sx[i] = fma(d.wx[1], xr[row][R + j - 1] + xr[row][R + j + 1], diag * v[i]);
        sz[i] = d.wz[1] * in[(U - 1 + 7) % 7][i];
    }
#pragma unroll
    for (int r = 2; r <= R; r++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int row = i >> 1, j = i & 1;
            sx[i] = fma(d.wx[r], xr[row][R + j - r] + xr[row][R + j + r], sx[i]);
            sz[i] = fma(d.wz[r], in[(U - r + 7) % 7][i], sz[i]);
        }
    /* y: up[k] = row r0-k, dn[k] = row r0+1+k (k = 1..6); row r0 pairs up[k] with (k == 1 ? own row 1 : dn[k-1]),
       row r0+1 pairs (k == 1 ? own row 0 : up[k-1]) with dn[k]; the y chains start from the z sums */
    double2 up[R + 1], dn[R + 1];
    up[0] = make_double2(v[0], v[1]); /* row r0   */
    dn[0] = make_double2(v[2], v[3]); /* row r0+1 */
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
}

template <class Cfg, int U, bool HASV, bool HASX, bool STEADY>
__device__ __forceinline__ void consume_plane22(const DenseDesc &d, const StepArgs &a, const unsigned char *stage, int p,
                                                bool active, bool act0, bool act1, uint32_t yoff, uint32_t voff,
                                                double *__restrict__ &dst, size_t plane_elems,
                                                double (&in)[7][4], double (&acc)[7][4], bool plane_is_zero)
{
    const int Nz = d.Nz;
    const double *ytile = reinterpret_cast<const double *>(stage);
    const double *vtile = reinterpret_cast<const double *>(stage + Cfg::OFF_V);
    const double *xtile = reinterpret_cast<const double *>(stage + Cfg::OFF_X);
    if (STEADY) {
        /* R <= p < Nz: the plane is inside the grid and so is the plane it completes -- one straight block, the xprev
           values are fetched with everything else */
        if (active) {
            double v[4];
            double2 xw0 = make_double2(0.0, 0.0), xw1 = xw0;
            if (HASX) {
                xw0 = *reinterpret_cast<const double2 *>(xtile + voff);
                xw1 = *reinterpret_cast<const double2 *>(xtile + voff + Cfg::XP);
            }
            interior_plane22<Cfg, U, HASV>(d, a, ytile + yoff, vtile, voff, in, acc, v);
#pragma unroll
            for (int r = 1; r <= R; r++)
#pragma unroll
                for (int i = 0; i < 4; i++) acc[(U - r + 7) % 7][i] = fma(d.wz[r], v[i], acc[(U - r + 7) % 7][i]);
#pragma unroll
            for (int i = 0; i < 4; i++) in[U][i] = v[i];
            double res[4];
#pragma unroll
            for (int i = 0; i < 4; i++) res[i] = acc[(U + 1) % 7][i];
            if (HASX) {
                res[0] = fma(-a.s2, xw0.x, res[0]); res[1] = fma(-a.s2, xw0.y, res[1]);
                res[2] = fma(-a.s2, xw1.x, res[2]); res[3] = fma(-a.s2, xw1.y, res[3]);
            }
            if (act0) stg128(dst, res[0], res[1]);
            stg128(dst + d.Nx, res[2], res[3]); /* active == act1 */
        }
        dst += plane_elems;
        return;
    }
    const bool interior = (p >= 0) && (p < Nz);
    const int o = p - R;
    const bool emit = o >= 0 && o < Nz;

    double v[4] = {0, 0, 0, 0};
    if (active && !plane_is_zero) {
        const double *cp = ytile + yoff; /* centre chunk of row r0: (r0 + HT) * YP + 2 xp + R */
        if (interior) {
            interior_plane22<Cfg, U, HASV>(d, a, cp, vtile, voff, in, acc, v);
        } else {
            const double2 w0 = *reinterpret_cast<const double2 *>(cp);
            const double2 w1 = *reinterpret_cast<const double2 *>(cp + Cfg::YP);
            v[0] = w0.x; v[1] = w0.y; v[2] = w1.x; v[3] = w1.y;
        }
    }
    if (p >= 0) { /* scatter the z terms into the 6 accumulators behind this plane */
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
            const double2 w0 = *reinterpret_cast<const double2 *>(xtile + voff);
            const double2 w1 = *reinterpret_cast<const double2 *>(xtile + voff + Cfg::XP);
            res[0] = fma(-a.s2, w0.x, acc[(U + 1) % 7][0]);
            res[1] = fma(-a.s2, w0.y, acc[(U + 1) % 7][1]);
            res[2] = fma(-a.s2, w1.x, acc[(U + 1) % 7][2]);
            res[3] = fma(-a.s2, w1.y, acc[(U + 1) % 7][3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) res[i] = acc[(U + 1) % 7][i];
        }
        if (act0) stg128(dst, res[0], res[1]);
        if (act1) stg128(dst + d.Nx, res[2], res[3]);
    }
    if (emit) dst += plane_elems; /* running output pointer: planes are emitted in order */
}

template <int WX, int WY, bool HASV, bool HASX>
__global__ void __launch_bounds__(TileCfg<WX, WY>::THREADS, 1)
stream_dense_kernel(const __grid_constant__ DenseMaps maps, const __grid_constant__ DenseDesc d, const StepArgs a,
                    const int nitems, unsigned int *__restrict__ sync_counter, const unsigned int sync_base)
{
    using Cfg = TileCfg<WX, WY>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *ring = smem_raw;
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)kStages * Cfg::STAGE_BYTES);
    uint64_t *empty = full + kStages;
    uint64_t *landed = empty + kStages; /* TMA completion; the merge warps turn it into `full` */

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
    /* ring position (stage, parity of its current use), continues across work items; kept as a pair of counters: the
       stage index as it % 5 is a multiply-high chain in front of every barrier wait */
    uint32_t rs = 0, rpar = 0;
#define CHEFSI_RING_ADVANCE() do { if (++rs == (uint32_t)kStages) { rs = 0; rpar ^= 1u; } } while (0)

    if (warp >= Cfg::CONSUMER_WARPS) {
        /* ================= producer warpgroup (one elected lane issues the TMA boxes) ================= */
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::PRODUCER_REGS));
        uint64_t *tb = landed; /* barrier the TMA boxes of a stage complete on */
        if (warp == Cfg::CONSUMER_WARPS && lane == 0) {
            unsigned int round = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, round++) {
                const int tile = item % (d.ntx * d.nty), n = item / (d.ntx * d.nty);
                const int x0 = tile_origin(tile % d.ntx, Cfg::TX, Nx), y0 = tile_origin(tile / d.ntx, Cfg::TY, Ny);
                /* which pieces this tile's haloed plane consists of (constant over z) */
                const bool wrap_top = yper && (y0 - R < 0), wrap_bot = yper && (y0 + Cfg::TY + R > Ny);
                const bool split_y = wrap_top || wrap_bot;
                const int ytop = wrap_top ? y0 - HT + Ny : y0 - HT;
                const int ybot = wrap_bot ? y0 + Cfg::TY - Ny : y0 + Cfg::TY;
                const bool need_l = xper && (x0 - R < 0), need_r = xper && (x0 + Cfg::TX + R > Nx);
                const uint32_t ybytes = (uint32_t)(Cfg::YP * Cfg::YROWS * 8 + (need_l ? SW * Cfg::TY * 8 : 0) +
                                                   (need_r ? SW * Cfg::TY * 8 : 0));
                if (sync_counter) {
                    /* round barrier between the producers of all (co-resident) CTAs: neighbouring tiles of a column are
                       marched at the same z, so their xy-halos hit in L2.  A barrier that does not complete within
                       ~2^22 polls (a CTA not resident: never with the cooperative launch) gives up -- the result is
                       unaffected -- and says so in sync_counter[1], which chefsi_get_stats reports. */
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
                    const bool need_y = interior || zper;          /* Dirichlet z: planes outside are zero */
                    const bool need_v = interior && HASV;
                    const bool need_x = (o >= 0 && o < Nz) && HASX;
                    if (!need_y && !need_x) continue;
                    const int s = (int)rs;
                    unsigned char *stage = ring + (size_t)s * Cfg::STAGE_BYTES;
                    mbar_wait(&empty[s], rpar ^ 1u);
                    mbar_expect_tx(&tb[s], (need_y ? ybytes : 0u) + (uint32_t)((need_v ? Cfg::XP * Cfg::TY * 8 : 0) +
                                                                                (need_x ? Cfg::XP * Cfg::TY * 8 : 0)));
                    if (need_y) {
                        if (!split_y) {
                            tma_load_4d(stage, &maps.y_full, x0 - R, y0 - HT, kz, n, &tb[s]);
                        } else {
                            tma_load_4d(stage, &maps.y_top, x0 - R, ytop, kz, n, &tb[s]);
                            tma_load_4d(stage + Cfg::YP * HT * 8, &maps.y_body, x0 - R, y0, kz, n, &tb[s]);
                            tma_load_4d(stage + Cfg::YP * (HT + Cfg::TY) * 8, &maps.y_bot, x0 - R, ybot, kz, n, &tb[s]);
                        }
                        if (need_l) tma_load_4d(stage + Cfg::OFF_L, &maps.y_strip, Nx - SW, y0, kz, n, &tb[s]);
                        if (need_r) tma_load_4d(stage + Cfg::OFF_R, &maps.y_strip, 0, y0, kz, n, &tb[s]);
                    }
                    if (need_v) tma_load_4d(stage + Cfg::OFF_V, &maps.veff, x0, y0, p, 0, &tb[s]);
                    if (need_x) tma_load_4d(stage + Cfg::OFF_X, &maps.xprev, x0, y0, o, n, &tb[s]);
                    CHEFSI_RING_ADVANCE();
                }
            }
        } else if (warp > Cfg::CONSUMER_WARPS) {
            /* ---- merge warps (the three spare warps of the producer warpgroup): same (item, plane) sequence as the
               producer lane; they copy the periodic-x strips of a landed stage into the zero-filled halo columns of
               its tile, so that the consumers read every x window with constant offsets ---- */
            const int ft = (int)threadIdx.x - (Cfg::CONSUMER_WARPS + 1) * 32; /* 0 .. 95 */
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int tile = item % (d.ntx * d.nty);
                const int x0 = tile_origin(tile % d.ntx, Cfg::TX, Nx);
                const bool need_l = xper && (x0 - R < 0), need_r = xper && (x0 + Cfg::TX + R > Nx);
                for (int p = -R; p < Nz + R; p++) {
                    const bool interior = (p >= 0 && p < Nz);
                    const int o = p - R;
                    const bool need_y = interior || zper;
                    const bool need_x = (o >= 0 && o < Nz) && HASX;
                    if (!need_y && !need_x) continue;
                    const int s = (int)rs;
                    unsigned char *stage = ring + (size_t)s * Cfg::STAGE_BYTES;
                    mbar_wait(&landed[s], rpar);
                    if (need_y && (need_l || need_r)) {
                        double *tile_d = reinterpret_cast<double *>(stage);
                        const double *sl = reinterpret_cast<const double *>(stage + Cfg::OFF_L);
                        const double *sr = reinterpret_cast<const double *>(stage + Cfg::OFF_R);
                        if (need_l)
                            for (int c = ft; c < 3 * Cfg::TY; c += 96) {
                                const int row = c / 3, j = c % 3, gi = x0 - R + 2 * j;
                                if (gi < 0)
                                    *reinterpret_cast<double2 *>(tile_d + (HT + row) * Cfg::YP + 2 * j) =
                                        *reinterpret_cast<const double2 *>(sl + row * SW + gi + SW);
                            }
                        if (need_r)
                            for (int c = ft; c < 4 * Cfg::TY; c += 96) {
                                const int row = c / 4, j = c % 4, col = Nx + 2 * j - (x0 - R);
                                if (col + 1 < Cfg::YP)
                                    *reinterpret_cast<double2 *>(tile_d + (HT + row) * Cfg::YP + col) =
                                        *reinterpret_cast<const double2 *>(sr + row * SW + 2 * j);
                            }
                    }
                    asm volatile("bar.sync 2, 96;" ::: "memory");
                    if (ft == 0) mbar_arrive(&full[s]);
                    CHEFSI_RING_ADVANCE();
                }
            }
        }
    } else {
        /* ================= consumer warps ================= */
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::CONSUMER_REGS));
        /* a thread owns a 2 x 2 patch (qx: x-pair index, ry: first of its two rows) */
        const int qx = lane & 15;
        const int ry = warp * 4 + 2 * (lane >> 4);
        /* the thread's offsets (doubles) into the haloed tile and into the Veff / xprev tiles; made opaque so that
           they live in a register instead of being re-derived from %tid (S2R + 6 ALU ops on the critical path of
           every plane step) */
        uint32_t yoff = (uint32_t)((ry + HT) * Cfg::YP + 2 * qx + R), voff = (uint32_t)(ry * Cfg::XP + 2 * qx);
        uint32_t lane0 = (lane == 0) ? 1u : 0u;
        asm volatile("" : "+r"(yoff), "+r"(voff), "+r"(lane0));
        double in[7][4], acc[7][4];
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int tile = item % (d.ntx * d.nty), n = item / (d.ntx * d.nty);
            const int tx = tile % d.ntx, ty = tile / d.ntx;
            const int x0 = tile_origin(tx, Cfg::TX, Nx), y0 = tile_origin(ty, Cfg::TY, Ny);
            const int gx = x0 + 2 * qx, gy = y0 + ry;
            /* a shifted last tile overlaps its neighbour: only the not yet covered points are computed */
            const bool act0 = (gx >= tx * Cfg::TX) && (gy >= ty * Cfg::TY);
            const bool act1 = (gx >= tx * Cfg::TX) && (gy + 1 >= ty * Cfg::TY); /* second row of the 2 x 2 patch */
            const bool active = act1;
            double *dst = reinterpret_cast<double *>(a.out) + (size_t)n * a.ld + (size_t)gy * Nx + gx; /* output plane 0 */
#pragma unroll
            for (int u = 0; u < 7; u++)
#pragma unroll
                for (int j = 0; j < 4; j++) { in[u][j] = 0.0; acc[u][j] = 0.0; }

#define CHEFSI_STEP(U, STEADY)                                                                           \
    if (STEADY || p + (U) < Nz + R) {                                                                    \
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
        consume_plane22<Cfg, (U), HASV, HASX, (STEADY)>(d, a, stage, pp, active, act0, act1, yoff, voff, dst, plane_elems, in, acc, zplane); \
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
            /* p runs over -6 .. Nz+5; groups start at p = -7 so that the phase U == (pp + 7) % 7 is
               compile-time inside the unrolled body (pp = -7 itself is skipped).  Groups whose seven planes all lie
               in [R, Nz) -- the plane is inside the grid and so is the plane it completes -- take a body without the
               boundary case distinctions */
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
/* 4-D view (x, y, z, column) of a block of dense columns; elements outside [0,N) read as zero */
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

template <int WX, int WY, bool HASV, bool HASX>
int launch_cfg(chefsi_ctx *ctx, const StepArgs &a)
{
    using Cfg = TileCfg<WX, WY>;
    const chefsi_grid_t &g = ctx->grid;
    const Layout &L = ctx->lay;
    DenseDesc d;
    d.Nx = g.Nx; d.Ny = g.Ny; d.Nz = g.Nz;
    d.bc[0] = g.BCx; d.bc[1] = g.BCy; d.bc[2] = g.BCz;
    d.ntx = (g.Nx + Cfg::TX - 1) / Cfg::TX;
    d.nty = (g.Ny + Cfg::TY - 1) / Cfg::TY;
    d.coef0 = ctx->desc.coef0;
    for (int r = 0; r <= R; r++) { d.wx[r] = ctx->desc.wx[r]; d.wy[r] = ctx->desc.wy[r]; d.wz[r] = ctx->desc.wz[r]; }
    /* s1 (and c) are applied through the weights: out = (s1 H') x - s2 xprev */
    d.coef0 = a.s1 * (d.coef0 + a.c);
    for (int r = 0; r <= R; r++) { d.wx[r] *= a.s1; d.wy[r] *= a.s1; d.wz[r] *= a.s1; }
    const long long nitems = (long long)a.ncol * d.ntx * d.nty;
    if (nitems > 0x7fffffffLL) { chefsi_fail(ctx, "stream kernel: too many work items"); return -1; }

    DenseMaps m;
    const void *xp = a.xprev ? a.xprev : a.x; /* never dereferenced when s2 == 0 */
    const int promo = ctx->tma_l2promo;
    if (!make_map(&m.y_full, a.x, L, a.ncol, Cfg::YP, Cfg::YROWS, promo) || !make_map(&m.y_top, a.x, L, a.ncol, Cfg::YP, HT, promo) ||
        !make_map(&m.y_body, a.x, L, a.ncol, Cfg::YP, Cfg::TY, promo) || !make_map(&m.y_bot, a.x, L, a.ncol, Cfg::YP, R, promo) ||
        !make_map(&m.y_strip, a.x, L, a.ncol, SW, Cfg::TY, promo) || !make_map(&m.xprev, xp, L, a.ncol, Cfg::XP, Cfg::TY, promo) ||
        !make_map(&m.veff, ctx->d_veff, L, 1, Cfg::XP, Cfg::TY, promo)) {
        chefsi_fail(ctx, "cuTensorMapEncodeTiled failed");
        return -1;
    }
    static_assert(Cfg::TX == 32 && Cfg::TY == 4 * Cfg::CONSUMER_WARPS && Cfg::CONSUMER_WARPS % 4 == 0, "2 x 2 mapping: 16 pairs x 4 rows per warp");
    auto kern = stream_dense_kernel<WX, WY, HASV, HASX>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) { chefsi_fail(ctx, "cudaFuncSetAttribute(stream): %s", cudaGetErrorString(e)); return -1; }
    const int grid = (int)((nitems < ctx->num_sms) ? nitems : ctx->num_sms);
    unsigned int *counter = nullptr;
    unsigned int base = 0;
    if (ctx->stream_gridsync && nitems > grid) {
        if (!ctx->d_sync) {
            if (cudaMalloc((void **)&ctx->d_sync, 256) != cudaSuccess || cudaMemset(ctx->d_sync, 0, 256) != cudaSuccess) {
                chefsi_fail(ctx, "stream kernel: cannot allocate the round-barrier counter");
                return -1;
            }
        }
        counter = ctx->d_sync;
        base = ctx->sync_arrivals;
        ctx->sync_arrivals += (unsigned int)nitems; /* every item arrives exactly once (wraps mod 2^32) */
    }
    int nit = (int)nitems;
    if (counter) { /* the barrier needs all CTAs co-resident: cooperative launch refuses otherwise */
        void *args[] = {(void *)&m, (void *)&d, (void *)&a, (void *)&nit, (void *)&counter, (void *)&base};
        e = cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(Cfg::THREADS), args, Cfg::SMEM, ctx->stream);
        if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
            /* the SMs are shared (MPS, another context, a smaller MIG slice): run without the round barrier -- the
               result is the same, only the xy-halos of neighbouring tiles hit L2 less often */
            cudaGetLastError();
            ctx->sync_arrivals = base; /* nothing will arrive for this launch */
            ctx->stream_gridsync = 0;
            counter = nullptr;
            kern<<<grid, Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(m, d, a, nit, counter, base);
        } else if (e != cudaSuccess) { chefsi_fail(ctx, "stream kernel cooperative launch: %s", cudaGetErrorString(e)); return -1; }
    } else {
        kern<<<grid, Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(m, d, a, nit, counter, base);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "stream kernel launch: %s", cudaGetErrorString(e)); return -1; }
    return 1;
}

}  // namespace

/* The dense streaming kernel needs: orthogonal cell, FD radius 6, real data, Nx even (a thread owns aligned x
 * pairs and TMA strides must be 16-byte multiples), at least one full 32 x 32 tile per plane, and -- for a periodic
 * y face -- a tile row count such that no 6-row halo box straddles the face (Ny mod 32 is 0 or >= 6).  Everything
 * else goes through the z-march kernels. */
bool stream_dense_supported(const chefsi_ctx *ctx, bool is_complex)
{
    using Cfg = TileCfg<2, 4>;
    const chefsi_grid_t &g = ctx->grid;
    if (ctx->force_general || is_complex) return false;
    if (g.cell_typ != 0 || g.FDn != R) return false;
    if (g.Nx % 2 != 0) return false;
    if (g.Nx < Cfg::TX || g.Ny < Cfg::TY || g.Nz < 2 * R) return false;
    if (g.BCy == 0 && g.Ny % Cfg::TY != 0 && g.Ny % Cfg::TY < R) return false;
    return true;
}

int launch_stencil_stream_dense(chefsi_ctx *ctx, const StepArgs &a)
{
    if (a.ncol <= 0) return 0;
    /* with / without the Veff tile and the xprev tile: compile-time, so the plane step carries no selects for them */
    const bool hv = a.veff != nullptr, hx = a.s2 != 0.0;
    if (hv) return hx ? launch_cfg<2, 4, true, true>(ctx, a) : launch_cfg<2, 4, true, false>(ctx, a);
    return hx ? launch_cfg<2, 4, false, true>(ctx, a) : launch_cfg<2, 4, false, false>(ctx, a);
}
