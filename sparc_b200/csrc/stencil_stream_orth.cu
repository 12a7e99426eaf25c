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
 *   - a work item is (orbital column, TX x TY tile of the xy-plane); the CTA marches the whole z
 *     extent of the item, so halos are re-read only in x/y, and neighbouring tiles of a column are
 *     in flight on other SMs at the same time, which turns those re-reads into L2 hits;
 *   - a dedicated producer warp stages each z-plane of the tile (+6-point halo) in a 4-deep shared
 *     memory ring with TMA bulk copies (cp.async.bulk -> UBLKCP), one per row segment, completion
 *     on an mbarrier; periodic wrap is done by the copy addresses (no materialised halo array),
 *     Dirichlet faces copy from a zero page;
 *   - 16x8-point warps: a thread owns 4 consecutive x of one row (vectorised 16 B smem loads,
 *     32 B global accesses); the row pitch is padded so every LDS.128 wavefront is conflict free;
 *   - the z direction never touches shared memory: each thread keeps the last 6 input planes and
 *     7 partial output accumulators of its 4 points in registers (scatter form: a plane adds its
 *     own x/y terms and the z terms of the 6 planes behind it when it arrives, and is added into
 *     the 6 accumulators behind it); the plane loop is unrolled by 7 so the register queues rotate
 *     by renaming instead of moves;
 *   - Veff, c, the recurrence scale s1 and the -s2*xprev term are applied in registers; the result
 *     plane (6 behind the one just loaded) is written with 256-bit stores.
 */
#include "chefsi_internal.h"

namespace {

constexpr int R = 6;        /* FD radius this kernel is specialised for (FD_ORDER 12) */
constexpr int kStages = 4;  /* shared memory ring depth */

template <int WX, int WY> struct TileCfg {
    static constexpr int TX = 16 * WX;           /* tile width  (points) */
    static constexpr int TY = 8 * WY;            /* tile height (points) */
    static constexpr int PITCH = TX + 2 * R + 2; /* doubles; (PITCH/2) odd -> conflict-free LDS.128 */
    static constexpr int ROWS = TY + 2 * R;
    static constexpr int PLANE = PITCH * ROWS;   /* doubles per ring slot */
    static constexpr int CONSUMER_WARPS = WX * WY;
    static constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
    static constexpr size_t SMEM = (size_t)kStages * PLANE * sizeof(double) + 2 * kStages * sizeof(unsigned long long);
    static_assert((PITCH / 2) % 2 == 1, "row pitch must be an odd number of 16-byte chunks");
};

struct StreamDesc {
    int Nx, Ny, Nz;
    int bc[3];
    int ntx, nty;
    double coef0;
    double wx[R + 1], wy[R + 1], wz[R + 1];
};

/* ---- PTX helpers ---------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
/* TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier */
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void ldg256(const double *p, double (&v)[4])
{
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3])
                 : "l"(p));
}
__device__ __forceinline__ void stg256(double *p, const double (&v)[4])
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

/* ---- producer: stage one haloed plane ---------------------------------------------------- */
template <class Cfg>
__device__ __forceinline__ void produce_plane(const StreamDesc &d, const double *__restrict__ col,
                                              const double *__restrict__ zero_page, int x0, int y0, int kz,
                                              int txe, double *slot, uint64_t *full, int lane)
{
    const int Lh = txe + 2 * R; /* haloed row length actually used */
    const int nrows = min(Cfg::TY, d.Ny - y0) + 2 * R;
    if (lane == 0) mbar_expect_tx(full, (uint32_t)(nrows * Lh * sizeof(double)));
    __syncwarp();
    for (int r = lane; r < nrows; r += 32) {
        int j = y0 - R + r;
        bool dead_row = false;
        if (j < 0) { j += d.Ny; dead_row = d.bc[1]; } else if (j >= d.Ny) { j -= d.Ny; dead_row = d.bc[1]; }
        const double *src_row = col + ((size_t)kz * d.Ny + j) * d.Nx;
        double *dst = slot + r * Cfg::PITCH;
        int t0 = 0;
        while (t0 < Lh) {
            int gx = x0 - R + t0;
            bool dead = dead_row;
            int len;
            if (gx < 0) { len = min(Lh - t0, -gx); gx += d.Nx; dead |= (d.bc[0] != 0); }
            else if (gx >= d.Nx) { len = Lh - t0; gx -= d.Nx; dead |= (d.bc[0] != 0); }
            else { len = min(Lh - t0, d.Nx - gx); }
            bulk_g2s(dst + t0, dead ? zero_page : src_row + gx, (uint32_t)(len * sizeof(double)), full);
            t0 += len;
        }
    }
}

/* ---- one plane step of a consumer thread -------------------------------------------------- */
/* U = p mod 7 (compile time): register-queue rotation by renaming.                            */
template <class Cfg, int U>
__device__ __forceinline__ void consume_plane(const StreamDesc &d, const StepArgs &a, const double *slot, int p,
                                              bool active, int qx, int ry, size_t gplane_off, size_t row_off,
                                              const double *__restrict__ veff, const double *__restrict__ xprev,
                                              double *__restrict__ out, double (&in)[7][4], double (&acc)[7][4],
                                              bool plane_is_zero)
{
    const int Nz = d.Nz;
    const bool interior = (p >= 0) && (p < Nz);
    /* issue the global loads of this step early */
    double ve[4] = {0, 0, 0, 0}, xp[4] = {0, 0, 0, 0};
    const int o = p - R;
    const bool emit = active && o >= 0 && o < Nz;
    if (active && interior && veff) ldg256(veff + (size_t)p * gplane_off + row_off, ve);
    if (emit && a.s2 != 0.0) ldg256(xprev + (size_t)o * gplane_off + row_off, xp);

    double v[4] = {0, 0, 0, 0};
    if (active && !plane_is_zero) {
        const double *rowp = slot + (ry + R) * Cfg::PITCH + 4 * qx; /* haloed row, element 0 = x0-6+4qx */
        if (interior) {
            double xr[16];
#pragma unroll
            for (int t = 0; t < 8; t++) {
                const double2 w = *reinterpret_cast<const double2 *>(rowp + 2 * t);
                xr[2 * t] = w.x;
                xr[2 * t + 1] = w.y;
            }
            double t4[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                v[j] = xr[R + j];
                t4[j] = (d.coef0 + a.c + ve[j]) * v[j];
            }
#pragma unroll
            for (int r = 1; r <= R; r++)
#pragma unroll
                for (int j = 0; j < 4; j++) t4[j] = fma(d.wx[r], xr[R + j - r] + xr[R + j + r], t4[j]);
#pragma unroll
            for (int r = 1; r <= R; r++) {
                const double2 u0 = *reinterpret_cast<const double2 *>(rowp - r * Cfg::PITCH + R);
                const double2 u1 = *reinterpret_cast<const double2 *>(rowp - r * Cfg::PITCH + R + 2);
                const double2 d0 = *reinterpret_cast<const double2 *>(rowp + r * Cfg::PITCH + R);
                const double2 d1 = *reinterpret_cast<const double2 *>(rowp + r * Cfg::PITCH + R + 2);
                t4[0] = fma(d.wy[r], u0.x + d0.x, t4[0]);
                t4[1] = fma(d.wy[r], u0.y + d0.y, t4[1]);
                t4[2] = fma(d.wy[r], u1.x + d1.x, t4[2]);
                t4[3] = fma(d.wy[r], u1.y + d1.y, t4[3]);
            }
#pragma unroll
            for (int r = 1; r <= R; r++)
#pragma unroll
                for (int j = 0; j < 4; j++) t4[j] = fma(d.wz[r], in[(U - r + 7) % 7][j], t4[j]);
#pragma unroll
            for (int j = 0; j < 4; j++) acc[U][j] = t4[j];
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
#pragma unroll
        for (int j = 0; j < 4; j++) res[j] = fma(-a.s2, xp[j], a.s1 * acc[(U + 1) % 7][j]);
        stg256(out + (size_t)o * gplane_off + row_off, res);
    }
}

template <int WX, int WY>
__global__ void __launch_bounds__(TileCfg<WX, WY>::THREADS, 1)
stream_orth_kernel(const __grid_constant__ StreamDesc d, const StepArgs a, const double *__restrict__ zero_page,
                   const int nitems)
{
    using Cfg = TileCfg<WX, WY>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *ring = reinterpret_cast<double *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)kStages * Cfg::PLANE);
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
    const size_t plane_elems = (size_t)d.Nx * d.Ny;
    uint32_t it = 0; /* ring position, continues across work items */

    if (warp == Cfg::CONSUMER_WARPS) {
        /* ================= producer warp ================= */
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int tile = item % (d.ntx * d.nty), n = item / (d.ntx * d.nty);
            const int x0 = (tile % d.ntx) * Cfg::TX, y0 = (tile / d.ntx) * Cfg::TY;
            const int txe = min(Cfg::TX, d.Nx - x0);
            const double *col = reinterpret_cast<const double *>(a.x) + (size_t)n * a.ld;
            for (int p = -R; p < Nz + R; p++) {
                int kz = p;
                if (p < 0) { if (!zper) continue; kz += Nz; }
                else if (p >= Nz) { if (!zper) continue; kz -= Nz; }
                const int s = it % kStages;
                mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
                produce_plane<Cfg>(d, col, zero_page, x0, y0, kz, txe, ring + (size_t)s * Cfg::PLANE, &full[s], lane);
                it++;
            }
        }
    } else {
        /* ================= consumer warps ================= */
        const int wx = warp % WX, wy = warp / WX;
        const int qx = wx * 4 + (lane & 3); /* quad index along x inside the tile */
        const int ry = wy * 8 + (lane >> 2); /* row inside the tile                */
        const double *veff = a.veff;
        double in[7][4], acc[7][4];
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int tile = item % (d.ntx * d.nty), n = item / (d.ntx * d.nty);
            const int x0 = (tile % d.ntx) * Cfg::TX, y0 = (tile / d.ntx) * Cfg::TY;
            const int gx = x0 + 4 * qx, gy = y0 + ry;
            const bool active = (gx < d.Nx) && (gy < d.Ny);
            const size_t row_off = (size_t)gy * d.Nx + gx;
            const double *xprev = reinterpret_cast<const double *>(a.xprev) + (size_t)n * a.ld;
            double *out = reinterpret_cast<double *>(a.out) + (size_t)n * a.ld;
#pragma unroll
            for (int u = 0; u < 7; u++)
#pragma unroll
                for (int j = 0; j < 4; j++) { in[u][j] = 0.0; acc[u][j] = 0.0; }

#define CHEFSI_STEP(U)                                                                                   \
    if (p + (U) < Nz + R) {                                                                              \
        const int pp = p + (U);                                                                          \
        const bool zplane = !zper && (pp < 0 || pp >= Nz);                                               \
        const double *slot = ring;                                                                       \
        int s = 0;                                                                                       \
        if (!zplane) {                                                                                   \
            s = it % kStages;                                                                            \
            slot = ring + (size_t)s * Cfg::PLANE;                                                        \
            mbar_wait(&full[s], (it / kStages) & 1);                                                     \
        }                                                                                                \
        consume_plane<Cfg, (U)>(d, a, slot, pp, active, qx, ry, plane_elems, row_off, veff, xprev, out,  \
                                in, acc, zplane);                                                        \
        if (!zplane) {                                                                                   \
            __syncwarp();                                                                                \
            if (lane == 0) mbar_arrive(&empty[s]);                                                       \
            it++;                                                                                        \
        }                                                                                                \
    }
            /* p runs over -6 .. Nz+5; the phase U = (p + 7) mod 7 is compile-time inside the unrolled body */
            for (int p = -R - 1; p < Nz + R; p += 7) {
                /* first group starts at p = -7 so that U == (pp + 7) % 7; pp = -7 itself is skipped */
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

template <int WX, int WY>
int launch_cfg(chefsi_ctx *ctx, const StepArgs &a, const double *zero_page)
{
    using Cfg = TileCfg<WX, WY>;
    const chefsi_grid_t &g = ctx->grid;
    StreamDesc d;
    d.Nx = g.Nx; d.Ny = g.Ny; d.Nz = g.Nz;
    d.bc[0] = g.BCx; d.bc[1] = g.BCy; d.bc[2] = g.BCz;
    d.ntx = (g.Nx + Cfg::TX - 1) / Cfg::TX;
    d.nty = (g.Ny + Cfg::TY - 1) / Cfg::TY;
    d.coef0 = ctx->desc.coef0;
    for (int r = 0; r <= R; r++) { d.wx[r] = ctx->desc.wx[r]; d.wy[r] = ctx->desc.wy[r]; d.wz[r] = ctx->desc.wz[r]; }
    const long long nitems = (long long)a.ncol * d.ntx * d.nty;
    if (nitems > 0x7fffffffLL) { chefsi_fail(ctx, "stream kernel: too many work items"); return -1; }
    auto kern = stream_orth_kernel<WX, WY>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) { chefsi_fail(ctx, "cudaFuncSetAttribute(stream): %s", cudaGetErrorString(e)); return -1; }
    const int grid = (int)((nitems < ctx->num_sms) ? nitems : ctx->num_sms);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, ctx->stream>>>(d, a, zero_page, (int)nitems);
    e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "stream kernel launch: %s", cudaGetErrorString(e)); return -1; }
    return 1;
}

}  // namespace

/* The streaming kernel needs: orthogonal cell, FD radius 6, real data, Nx a multiple of 4 (each
 * thread owns an aligned quad and all TMA row segments must be 16-byte multiples), a leading
 * dimension that keeps columns 32-byte aligned, and a grid big enough that tiles are not mostly
 * halo.  Everything else goes through the general kernel. */
bool stream_orth_supported(const chefsi_ctx *ctx, bool is_complex)
{
    const chefsi_grid_t &g = ctx->grid;
    if (ctx->force_general) return false;
    if (is_complex) return false;
    if (g.cell_typ != 0 || g.FDn != R) return false;
    if (g.Nx % 4 != 0 || ctx->ld % 4 != 0) return false;
    if (g.Nx < 32 || g.Ny < 16 || g.Nz < 2 * R) return false;
    return true;
}

int launch_stencil_stream_orth(chefsi_ctx *ctx, const StepArgs &a, bool is_complex)
{
    (void)is_complex;
    if (a.ncol <= 0) return 0;
    /* zero page for Dirichlet halos: the tail of the Veff allocation is kept zeroed by set_grid */
    const double *zero_page = ctx->d_veff + ctx->ld;
    return launch_cfg<2, 4>(ctx, a, zero_page);
}
