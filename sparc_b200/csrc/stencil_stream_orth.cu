/*
 * stencil_stream_orth.cu -- streaming fused Chebyshev step for orthogonal cells (sm_100a).
 *
 *   out = s1 * ( (-1/2 Lap + Veff + c) x ) - s2 * xprev          (FP64, radius-6 star stencil)
 *
 * This is the kernel the 160^3 x 4096 workload runs on; it replaces, per Chebyshev degree, the
 * reference's haloed copy + stencil_3axis_thread_radius6 (lapVecRoutines.c:185-227, called from
 * :586) + the three scale/axpy/swap passes of ChebyshevFiltering (eigenSolver.c:764-768,787-794)
 * with ONE pass that reads x and xprev once and writes out once (24 B per grid point).
 *
 * Design (2.5-D streaming, one persistent CTA per SM):
 *   - columns live in the halo-padded internal layout (chefsi_internal.h: Layout), so the haloed
 *     TX+12 x TY+12 tile of any xy-plane is one in-bounds box of a 4-D tensor map (x, y, z, column);
 *   - a work item is (orbital column, TX x TY tile of the xy-plane); the CTA marches the whole z
 *     extent of the item, so halos are re-read only in x/y, and neighbouring tiles of a column are
 *     in flight on other SMs at the same time, which turns those re-reads into L2 hits;
 *   - a producer warp streams, per z-plane, three TMA boxes (cp.async.bulk.tensor -> UTMALDG) into a
 *     kStages-deep shared-memory ring: the haloed x tile of plane p, the Veff tile of plane p and the
 *     xprev tile of plane p-6 (the plane whose result is completed by plane p); completion is counted
 *     on an mbarrier per stage, consumers release a stage through a second mbarrier;
 *   - 16x8-point warps: a thread owns 4 consecutive x of one row (16 B shared loads, 32 B global
 *     stores); box widths are chosen so that every row pitch is an odd number of 16-byte chunks, which
 *     makes every LDS.128 wavefront of this thread layout bank-conflict free;
 *   - the z direction never touches shared memory: each thread keeps the last 6 input planes and
 *     7 partial output accumulators of its 4 points in registers (scatter form: a plane adds its
 *     own x/y terms and the z terms of the 6 planes behind it when it arrives, and is added into
 *     the 6 accumulators behind it); the plane loop is unrolled by 7 so the register queues rotate
 *     by renaming instead of moves;
 *   - Veff, c, the recurrence scale s1 and the -s2*xprev term are applied in registers; the result
 *     plane (6 behind the one just loaded) is written with 256-bit stores, together with its periodic
 *     images in the halo pads (the next step's TMA boxes read them; Dirichlet pads stay zero).
 */
#include <cuda.h>

#include "chefsi_internal.h"
#include "tma_ring.cuh"

namespace {

using namespace tma_ring;

constexpr int R = 6;        /* FD radius this kernel is specialised for (FD_ORDER 12) */
constexpr int kStages = 5;  /* shared memory ring depth */

template <int WX, int WY> struct TileCfg {
    static constexpr int TX = 16 * WX;           /* tile width  (points) */
    static constexpr int TY = 8 * WY;            /* tile height (points) */
    static constexpr int YP = TX + 2 * R + 2;    /* haloed tile pitch (doubles): odd number of 16 B chunks */
    static constexpr int YROWS = TY + 2 * R;
    static constexpr int XP = TX + 2;            /* xprev / Veff tile pitch */
    static constexpr int Y_BYTES = ((YP * YROWS * 8 + 127) / 128) * 128;
    static constexpr int X_BYTES = ((XP * TY * 8 + 127) / 128) * 128;
    static constexpr int STAGE_BYTES = Y_BYTES + 2 * X_BYTES;
    static constexpr int CONSUMER_WARPS = WX * WY;
    static constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
    static constexpr size_t SMEM = (size_t)kStages * STAGE_BYTES + 2 * kStages * sizeof(unsigned long long);
    static_assert((YP / 2) % 2 == 1 && (XP / 2) % 2 == 1, "row pitches must be an odd number of 16-byte chunks");
};

struct StreamDesc {
    int Nx, Ny, Nz;
    int Nxp, Nyp, px, py;
    int bc[3];
    int ntx, nty;
    double coef0;
    double wx[R + 1], wy[R + 1], wz[R + 1];
};

/* ---- one plane step of a consumer thread -------------------------------------------------- */
/* U = (p + 7) mod 7 (compile time): register-queue rotation by renaming.                      */
template <class Cfg, int U>
__device__ __forceinline__ void consume_plane(const StreamDesc &d, const StepArgs &a, const unsigned char *stage, int p,
                                              bool active, int qx, int ry, double *__restrict__ out_row,
                                              size_t plane_elems, int img_x, int img_y, double (&in)[7][4],
                                              double (&acc)[7][4], bool plane_is_zero)
{
    const int Nz = d.Nz;
    const bool interior = (p >= 0) && (p < Nz);
    const int o = p - R;
    const bool emit = active && o >= 0 && o < Nz;
    const double *ytile = reinterpret_cast<const double *>(stage);
    const double *vtile = reinterpret_cast<const double *>(stage + Cfg::Y_BYTES);
    const double *xtile = reinterpret_cast<const double *>(stage + Cfg::Y_BYTES + Cfg::X_BYTES);

    double v[4] = {0, 0, 0, 0};
    if (active && !plane_is_zero) {
        const double *rowp = ytile + (ry + R) * Cfg::YP + 4 * qx; /* haloed row, element 0 = x0-6+4qx */
        if (interior) {
            double xr[16];
#pragma unroll
            for (int t = 0; t < 8; t++) {
                const double2 w = *reinterpret_cast<const double2 *>(rowp + 2 * t);
                xr[2 * t] = w.x;
                xr[2 * t + 1] = w.y;
            }
            double ve[4] = {0, 0, 0, 0};
            if (a.veff) {
                const double2 w0 = *reinterpret_cast<const double2 *>(vtile + ry * Cfg::XP + 4 * qx);
                const double2 w1 = *reinterpret_cast<const double2 *>(vtile + ry * Cfg::XP + 4 * qx + 2);
                ve[0] = w0.x; ve[1] = w0.y; ve[2] = w1.x; ve[3] = w1.y;
            }
            double t4[4], sx[4], sy[4], sz[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                v[j] = xr[R + j];
                t4[j] = (d.coef0 + a.c + ve[j]) * v[j];
                sx[j] = d.wx[1] * (xr[R + j - 1] + xr[R + j + 1]);
                sz[j] = d.wz[1] * in[(U - 1 + 7) % 7][j];
            }
#pragma unroll
            for (int r = 2; r <= R; r++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    sx[j] = fma(d.wx[r], xr[R + j - r] + xr[R + j + r], sx[j]);
                    sz[j] = fma(d.wz[r], in[(U - r + 7) % 7][j], sz[j]);
                }
#pragma unroll
            for (int r = 1; r <= R; r++) {
                const double2 u0 = *reinterpret_cast<const double2 *>(rowp - r * Cfg::YP + R);
                const double2 u1 = *reinterpret_cast<const double2 *>(rowp - r * Cfg::YP + R + 2);
                const double2 d0 = *reinterpret_cast<const double2 *>(rowp + r * Cfg::YP + R);
                const double2 d1 = *reinterpret_cast<const double2 *>(rowp + r * Cfg::YP + R + 2);
                if (r == 1) {
                    sy[0] = d.wy[1] * (u0.x + d0.x);
                    sy[1] = d.wy[1] * (u0.y + d0.y);
                    sy[2] = d.wy[1] * (u1.x + d1.x);
                    sy[3] = d.wy[1] * (u1.y + d1.y);
                } else {
                    sy[0] = fma(d.wy[r], u0.x + d0.x, sy[0]);
                    sy[1] = fma(d.wy[r], u0.y + d0.y, sy[1]);
                    sy[2] = fma(d.wy[r], u1.x + d1.x, sy[2]);
                    sy[3] = fma(d.wy[r], u1.y + d1.y, sy[3]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) acc[U][j] = (t4[j] + sx[j]) + (sy[j] + sz[j]);
        } else {
            const double2 w0 = *reinterpret_cast<const double2 *>(rowp + R);
            const double2 w1 = *reinterpret_cast<const double2 *>(rowp + R + 2);
            v[0] = w0.x; v[1] = w0.y; v[2] = w1.x; v[3] = w1.y;
        }
    }
    if (p >= 0) { /* scatter the z terms into the 6 accumulators behind this plane */
#pragma unroll
        for (int r = 1; r <= R; r++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[(U - r + 7) % 7][j] = fma(d.wz[r], v[j], acc[(U - r + 7) % 7][j]);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) in[U][j] = v[j];

    if (emit) {
        double res[4];
        if (a.s2 != 0.0) {
            const double2 w0 = *reinterpret_cast<const double2 *>(xtile + ry * Cfg::XP + 4 * qx);
            const double2 w1 = *reinterpret_cast<const double2 *>(xtile + ry * Cfg::XP + 4 * qx + 2);
            res[0] = fma(-a.s2, w0.x, a.s1 * acc[(U + 1) % 7][0]);
            res[1] = fma(-a.s2, w0.y, a.s1 * acc[(U + 1) % 7][1]);
            res[2] = fma(-a.s2, w1.x, a.s1 * acc[(U + 1) % 7][2]);
            res[3] = fma(-a.s2, w1.y, a.s1 * acc[(U + 1) % 7][3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) res[j] = a.s1 * acc[(U + 1) % 7][j];
        }
        double *dst = out_row + (size_t)o * plane_elems;
        stg256(dst, res);
        /* periodic images into the halo pads (img_x / img_y: element offsets, 0 = none) */
        if (img_x) stg256(dst + img_x, res);
        if (img_y) stg256(dst + img_y, res);
    }
}

template <int WX, int WY>
__global__ void __launch_bounds__(TileCfg<WX, WY>::THREADS, 1)
stream_orth_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_xprev,
                   const __grid_constant__ CUtensorMap map_veff, const __grid_constant__ StreamDesc d, const StepArgs a,
                   const int nitems, unsigned int *__restrict__ sync_counter, const unsigned int sync_base)
{
    using Cfg = TileCfg<WX, WY>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *ring = smem_raw;
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)kStages * Cfg::STAGE_BYTES);
    uint64_t *empty = full + kStages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], Cfg::CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int Nz = d.Nz;
    const bool zper = (d.bc[2] == 0);
    const size_t plane_elems = (size_t)d.Nxp * d.Nyp;
    uint32_t it = 0; /* ring position, continues across work items */

    if (warp == Cfg::CONSUMER_WARPS) {
        /* ================= producer warp (one elected lane issues the TMA boxes) ================= */
        if (lane == 0) {
            unsigned int round = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, round++) {
                const int tile = item % (d.ntx * d.nty), n = item / (d.ntx * d.nty);
                const int x0 = (tile % d.ntx) * Cfg::TX, y0 = (tile / d.ntx) * Cfg::TY;
                if (sync_counter) {
                    /* Round barrier between the producers of all (co-resident) CTAs: neighbouring tiles of a
                       column are marched by neighbouring CTAs in the same round; starting them together keeps
                       their z positions within the L2 residence time of a line (~20 us at 5.6 TB/s), so the
                       xy-halo of a tile is an L2 hit on the plane its neighbour fetched (a CTA that runs
                       ahead misses and is slowed down, one that lags hits: the skew is self-correcting).
                       Only the CTAs that have an item in this round take part; the spin is bounded. */
                    const unsigned int in_round = (unsigned int)min((long long)gridDim.x, (long long)nitems - (long long)round * gridDim.x);
                    const unsigned int done_before = round * gridDim.x; /* arrivals of the earlier (full) rounds */
                    __threadfence();
                    atomicAdd(sync_counter, 1u);
                    const unsigned int target = sync_base + done_before + in_round;
                    unsigned int spins = 0;
                    while ((int)(*(volatile unsigned int *)sync_counter - target) < 0 && ++spins < (1u << 22)) __nanosleep(64);
                }
                for (int p = -R; p < Nz + R; p++) {
                    int kz = p;
                    const bool interior = (p >= 0 && p < Nz);
                    if (p < 0) kz += Nz; else if (p >= Nz) kz -= Nz;
                    const int o = p - R;
                    const bool need_y = interior || zper;          /* Dirichlet z: planes outside are zero */
                    const bool need_v = interior && a.veff != nullptr;
                    const bool need_x = (o >= 0 && o < Nz) && a.s2 != 0.0;
                    if (!need_y && !need_x) continue;
                    const int s = it % kStages;
                    unsigned char *stage = ring + (size_t)s * Cfg::STAGE_BYTES;
                    mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
                    mbar_expect_tx(&full[s], (uint32_t)((need_y ? Cfg::YP * Cfg::YROWS * 8 : 0) +
                                                        (need_v ? Cfg::XP * Cfg::TY * 8 : 0) +
                                                        (need_x ? Cfg::XP * Cfg::TY * 8 : 0)));
                    if (need_y) tma_load_4d(stage, &map_x, d.px + x0 - R, d.py + y0 - R, kz, n, &full[s]);
                    if (need_v) tma_load_4d(stage + Cfg::Y_BYTES, &map_veff, d.px + x0, d.py + y0, p, 0, &full[s]);
                    if (need_x)
                        tma_load_4d(stage + Cfg::Y_BYTES + Cfg::X_BYTES, &map_xprev, d.px + x0, d.py + y0, o, n, &full[s]);
                    it++;
                }
            }
        }
    } else {
        /* ================= consumer warps ================= */
        const int wx = warp % WX, wy = warp / WX;
        const int qx = wx * 4 + (lane & 3); /* quad index along x inside the tile */
        const int ry = wy * 8 + (lane >> 2); /* row inside the tile                */
        double in[7][4], acc[7][4];
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int tile = item % (d.ntx * d.nty), n = item / (d.ntx * d.nty);
            const int x0 = (tile % d.ntx) * Cfg::TX, y0 = (tile / d.ntx) * Cfg::TY;
            const int gx = x0 + 4 * qx, gy = y0 + ry;
            const bool active = (gx < d.Nx) && (gy < d.Ny);
            double *out_row = reinterpret_cast<double *>(a.out) + (size_t)n * a.ld +
                              ((size_t)(gy + d.py)) * d.Nxp + (gx + d.px);
            /* where this thread's quad is mirrored in the halo pads (periodic faces only) */
            int img_x = 0, img_y = 0;
            if (d.bc[0] == 0) { if (gx < d.px) img_x = d.Nx; else if (gx >= d.Nx - d.px) img_x = -d.Nx; }
            if (d.bc[1] == 0) { if (gy < d.py) img_y = d.Ny * d.Nxp; else if (gy >= d.Ny - d.py) img_y = -d.Ny * d.Nxp; }
#pragma unroll
            for (int u = 0; u < 7; u++)
#pragma unroll
                for (int j = 0; j < 4; j++) { in[u][j] = 0.0; acc[u][j] = 0.0; }

#define CHEFSI_STEP(U)                                                                                   \
    if (p + (U) < Nz + R) {                                                                              \
        const int pp = p + (U);                                                                          \
        const bool zplane = !zper && (pp < 0 || pp >= Nz);   /* Dirichlet z: the plane is zero */        \
        const bool use_stage = !zplane || (pp - R >= 0 && pp - R < Nz && a.s2 != 0.0);                   \
        const unsigned char *stage = ring;                                                               \
        int s = 0;                                                                                       \
        if (use_stage) {                                                                                 \
            s = it % kStages;                                                                            \
            stage = ring + (size_t)s * Cfg::STAGE_BYTES;                                                 \
            mbar_wait(&full[s], (it / kStages) & 1);                                                     \
        }                                                                                                \
        consume_plane<Cfg, (U)>(d, a, stage, pp, active, qx, ry, out_row, plane_elems, img_x, img_y,     \
                                in, acc, zplane);                                                        \
        if (use_stage) {                                                                                 \
            __syncwarp();                                                                                \
            if (lane == 0) mbar_arrive(&empty[s]);                                                       \
            it++;                                                                                        \
        }                                                                                                \
    }
            /* p runs over -6 .. Nz+5; groups start at p = -7 so that the phase U == (pp + 7) % 7 is
               compile-time inside the unrolled body (pp = -7 itself is skipped) */
            for (int p = -R - 1; p < Nz + R; p += 7) {
                if (p + 0 >= -R) { CHEFSI_STEP(0) }
                CHEFSI_STEP(1)
                CHEFSI_STEP(2)
                CHEFSI_STEP(3)
                CHEFSI_STEP(4)
                CHEFSI_STEP(5)
                CHEFSI_STEP(6)
            }
#undef CHEFSI_STEP
        }
    }
}

/* ---- host side ---------------------------------------------------------------------------- */
/* 4-D view (x, y, z, column) of a block of columns in the internal layout */
bool make_map(CUtensorMap *map, const void *base, const Layout &L, int ncol, int box_x, int box_y, int promo)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[4] = {(cuuint64_t)L.Nxp, (cuuint64_t)L.Nyp, (cuuint64_t)L.Nz, (cuuint64_t)(ncol > 0 ? ncol : 1)};
    cuuint64_t strides[3] = {(cuuint64_t)L.Nxp * 8, (cuuint64_t)L.plane * 8, (cuuint64_t)L.ld * 8};
    cuuint32_t box[4] = {(cuuint32_t)box_x, (cuuint32_t)box_y, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void *>(base), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int WX, int WY>
int launch_cfg(chefsi_ctx *ctx, const StepArgs &a)
{
    using Cfg = TileCfg<WX, WY>;
    const chefsi_grid_t &g = ctx->grid;
    const Layout &L = ctx->lay;
    StreamDesc d;
    d.Nx = g.Nx; d.Ny = g.Ny; d.Nz = g.Nz;
    d.Nxp = L.Nxp; d.Nyp = L.Nyp; d.px = L.px; d.py = L.py;
    d.bc[0] = g.BCx; d.bc[1] = g.BCy; d.bc[2] = g.BCz;
    d.ntx = (g.Nx + Cfg::TX - 1) / Cfg::TX;
    d.nty = (g.Ny + Cfg::TY - 1) / Cfg::TY;
    d.coef0 = ctx->desc.coef0;
    for (int r = 0; r <= R; r++) { d.wx[r] = ctx->desc.wx[r]; d.wy[r] = ctx->desc.wy[r]; d.wz[r] = ctx->desc.wz[r]; }
    const long long nitems = (long long)a.ncol * d.ntx * d.nty;
    if (nitems > 0x7fffffffLL) { chefsi_fail(ctx, "stream kernel: too many work items"); return -1; }

    CUtensorMap mx, mp, mv;
    const void *xp = a.xprev ? a.xprev : a.x; /* never dereferenced when s2 == 0 */
    const int promo = ctx->tma_l2promo;
    if (!make_map(&mx, a.x, L, a.ncol, Cfg::YP, Cfg::YROWS, promo) || !make_map(&mp, xp, L, a.ncol, Cfg::XP, Cfg::TY, promo) ||
        !make_map(&mv, ctx->d_veff, L, 1, Cfg::XP, Cfg::TY, promo)) {
        chefsi_fail(ctx, "cuTensorMapEncodeTiled failed");
        return -1;
    }
    auto kern = stream_orth_kernel<WX, WY>;
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
        void *args[] = {(void *)&mx, (void *)&mp, (void *)&mv, (void *)&d, (void *)&a, (void *)&nit, (void *)&counter, (void *)&base};
        e = cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(Cfg::THREADS), args, Cfg::SMEM, ctx->stream);
        if (e != cudaSuccess) { chefsi_fail(ctx, "stream kernel cooperative launch: %s", cudaGetErrorString(e)); return -1; }
    } else {
        kern<<<grid, Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(mx, mp, mv, d, a, nit, counter, base);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "stream kernel launch: %s", cudaGetErrorString(e)); return -1; }
    return 1;
}

}  // namespace

/* The streaming kernel needs: orthogonal cell, FD radius 6, real data, Nx a multiple of 4 (each
 * thread owns an aligned quad; TMA strides must be 16-byte multiples), the halo-padded layout, and a
 * grid big enough that tiles are not mostly halo.  Everything else goes through the general kernel. */
bool stream_layout_wanted(const chefsi_grid_t &g)
{
    if (g.cell_typ != 0 || g.FDn != R) return false;
    if (g.Nx % 4 != 0) return false;
    if (g.Nx < 32 || g.Ny < 16 || g.Nz < 2 * R) return false;
    return true;
}

bool stream_orth_supported(const chefsi_ctx *ctx, bool is_complex)
{
    if (ctx->force_general || is_complex) return false;
    if (ctx->dense_stream) return ctx->lay.px == 0 && ctx->lay.py == 0 && stream_dense_wanted(ctx->grid, ctx->stream_variant);
    return ctx->lay.px == 8 && ctx->lay.py == R && stream_layout_wanted(ctx->grid);
}

int launch_stencil_stream_orth(chefsi_ctx *ctx, const StepArgs &a, bool is_complex)
{
    (void)is_complex;
    if (a.ncol <= 0) return 0;
    if (ctx->dense_stream) return launch_stencil_stream_dense(ctx, a);
    return launch_cfg<2, 4>(ctx, a);
}
