/*
 * stencil_zmarch.cu -- general z-marching fused Chebyshev/Hamiltonian step kernel (sm_100a).
 *
 *   out = s1 * ( (-1/2 Lap + Veff + c) x ) - s2 * xprev
 *
 * Same job and same arithmetic (operation for operation) as stencil_general.cu -- every cell type SPARC's
 * filter supports (cell_typ 0 and 11..17, lapVecRoutines.c:940-1331), periodic or Dirichlet faces, real
 * (Gamma) or complex (k-point, Bloch-phase halos, lapVecRoutinesKpt.c) data, any grid size, FD radius 6 --
 * but organised as a 2.5-D sweep instead of 3-D bricks: a CTA owns a TX x TY tile of the xy-plane of one
 * orbital column and marches the whole z extent, keeping a ring of the last 14 haloed planes in shared
 * memory.  A brick of the 3-D kernel re-reads an FDn-wide halo on all six faces (11-17 x the useful points
 * at FDn = 6); the sweep re-reads it only in x/y (3.4 x), which is what makes the k-point and non-orthogonal
 * paths usable on large grids (160^3: 20-30 x faster, profiles/r1_secondary_synthetic.log).
 *
 * Per output plane k:
 *   1. plane k+6 (haloed: tile + 6 points on each side incl. corners) is staged into the ring; points
 *      that leave the cell take the periodically wrapped value (times exp(i k.L) for complex data) or zero
 *      on Dirichlet faces -- the rule of the reference's x_ex copy at np = 1
 *      (lapVecRoutines.c:536-577,1183-1204; lapVecRoutinesKpt.c:370-462,647-865);
 *   2. for non-orthogonal cells the intermediate first-derivative fields of the reference's two-stage
 *      mixed-derivative composition (Calc_DX gradVecRoutines.c:318, Calc_DX1_DX2 lapVecRoutines.c:1429) are
 *      formed for this plane on the tile extended along x or y (in a per-plane buffer), or -- for the
 *      component that is extended along z (cell_typ 15) -- for plane k+6 on the tile itself (in a second
 *      ring of 14 planes);
 *   3. every thread evaluates the star stencil (stencil_3axis_thread_v2 lapVecRoutines.c:257, stencil_4comp
 *      :1481, stencil_5comp :1540) for its point of plane k, adds (Veff + c) x and applies the three-term
 *      recurrence scaling (eigenSolver.c:763-768,787-794) before the single store.
 * Global loads run PF = 4 planes ahead in registers: plane k+7+PF is issued when plane k+7 is committed to the
 * ring (the ring has one plane more than the stencil needs, so that slot is free while plane k is computed), which
 * keeps ~4 planes per CTA in flight and hides the DRAM/L2 latency that a load-then-sync loop exposes.  One __syncthreads per plane for orthogonal cells, two otherwise.
 */
#include "chefsi_internal.h"
#include "cplx.cuh"

namespace {

constexpr int F = 6;           /* FD radius this kernel is specialised for */
constexpr int RING = 2 * F + 2; /* planes kept in shared memory */
constexpr int PF = 4;           /* planes in flight in registers between global memory and the ring */

__device__ __forceinline__ int ring_slot(int kk)
{
    int s = (kk + 4 * RING) % RING; /* kk >= -F */
    return s;
}

template <typename T, int TX, int TY>
__global__ void __launch_bounds__(TX *TY)
stencil_zmarch_kernel(const __grid_constant__ StencilDesc d, const StepArgs a, const int ntx, const int dext_elems)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int EX = TX + 2 * F, EY = TY + 2 * F, PL = EX * EY, NT = TX * TY;
    T *win = reinterpret_cast<T *>(smem_raw); /* [RING][PL] haloed planes */
    T *dext = win + RING * PL;                /* [2][dext_elems] in-plane extended derivative fields of plane k */
    T *dzr = dext + 2 * dext_elems;           /* [RING][NT] z-extended derivative field (cell_typ 15 only) */

    const int tid = threadIdx.x;
    const int lx = tid % TX, ly = tid / TX;
    const int x0 = (blockIdx.x % ntx) * TX, y0 = (blockIdx.x / ntx) * TY;
    const size_t col = (size_t)blockIdx.y * a.ld;
    const T *__restrict__ x = reinterpret_cast<const T *>(a.x) + col;
    const T *__restrict__ xprev = reinterpret_cast<const T *>(a.xprev);
    T *__restrict__ out = reinterpret_cast<T *>(a.out) + col;
    const int Nx = d.Nx, Ny = d.Ny, Nz = d.Nz;

    /* which mixed component (if any) is extended along z */
    int qz = -1;
    for (int q = 0; q < d.nmix; q++)
        if (d.mix[q].ext == 2) qz = q;

    /* ---- 1. stage one haloed plane ----------------------------------------------------------- */
    /* The (x, y) part of the wrap / Dirichlet / Bloch-phase rule of a thread's elements does not depend on the
       plane: resolve it once.  el_off < 0: the element is zero (outside a Dirichlet face or the grid). */
    constexpr int NLD = (PL + NT - 1) / NT;
    int el_off[NLD], el_q[NLD];
#pragma unroll
    for (int e = 0; e < NLD; e++) {
        const int idx = tid + e * NT;
        const int ip = idx % EX, jp = idx / EX;
        int i = x0 - F + ip, j = y0 - F + jp;
        int ox = 0, oy = 0;
        if (i < 0) { i += Nx; ox = -1; } else if (i >= Nx) { i -= Nx; ox = 1; }
        if (j < 0) { j += Ny; oy = -1; } else if (j >= Ny) { j -= Ny; oy = 1; }
        const bool dead = idx >= PL || (ox && d.bc[0]) || (oy && d.bc[1]) || i < 0 || i >= Nx || j < 0 || j >= Ny;
        el_off[e] = dead ? -1 : (int)lay_pos(d.lay, i, j, 0);
        el_q[e] = (oy + 1) * 3 + (ox + 1);
    }
    const size_t plane_elems = d.lay.plane;
    /* issue the global loads of plane kk into registers (the caller overlaps them with the arithmetic of the
       plane it is working on) ... */
    auto fetch_plane = [&](int kk, T (&v)[NLD]) {
        int k = kk, oz = 0;
        if (k < 0) { k += Nz; oz = -1; } else if (k >= Nz) { k -= Nz; oz = 1; }
        const bool zdead = (oz && d.bc[2]) || k < 0 || k >= Nz;
        const T *xp = x + (size_t)(zdead ? 0 : k) * plane_elems;
#pragma unroll
        for (int e = 0; e < NLD; e++) {
            v[e] = cplx::zero<T>();
            if (!zdead && el_off[e] >= 0) v[e] = xp[el_off[e]];
        }
    };
    /* ... and put them into the ring slot of plane kk; the Bloch phase is applied here, not at the load, so that
       nothing waits on the loads until the plane is due */
    auto commit_plane = [&](int kk, const T (&v)[NLD]) {
        T *dst = win + ring_slot(kk) * PL;
        const int oz = (kk < 0) ? -1 : (kk >= Nz ? 1 : 0);
#pragma unroll
        for (int e = 0; e < NLD; e++) {
            const int idx = tid + e * NT;
            T w = v[e];
            if (cplx::is_complex<T>::value && (oz != 0 || el_q[e] != 4)) {
                const int q = (oz + 1) * 9 + el_q[e];
                w = cplx::mul_phase(w, d.ph_re[q], d.ph_im[q]);
            }
            if (idx < PL) dst[idx] = w;
        }
    };

    /* Per mixed component, resolved once: in-plane stride of the inner derivative axes (0 when the axis is z) */
    /* (scalars selected by q, not arrays: a runtime index would put them in local memory) */
    auto stride_of = [&](int ax) { return (ax == 0) ? 1 : (ax == 1 ? EX : 0); };
    const int in1_0 = d.nmix > 0 ? stride_of(d.mix[0].ax1) : 0, in2_0 = d.nmix > 0 ? stride_of(d.mix[0].ax2) : 0;
    const int in1_1 = d.nmix > 1 ? stride_of(d.mix[1].ax1) : 0, in2_1 = d.nmix > 1 ? stride_of(d.mix[1].ax2) : 0;
    const bool z1_0 = d.nmix > 0 && d.mix[0].ax1 == 2, z2_0 = d.nmix > 0 && d.mix[0].ax2 == 2, two_0 = d.nmix > 0 && d.mix[0].ax2 >= 0;
    const bool z1_1 = d.nmix > 1 && d.mix[1].ax1 == 2, z2_1 = d.nmix > 1 && d.mix[1].ax2 == 2, two_1 = d.nmix > 1 && d.mix[1].ax2 >= 0;
    /* first derivative(s) of component q at position pos of the haloed plane whose ring offsets are zo[F +- r] */
    auto inner = [&](int q, const int (&zo)[2 * F + 1], int pos) -> T {
        const MixedComp &mc = d.mix[q];
        const int i1 = q ? in1_1 : in1_0, i2 = q ? in2_1 : in2_0;
        const bool zz1 = q ? z1_1 : z1_0, zz2 = q ? z2_1 : z2_0, tw = q ? two_1 : two_0;
        T t1 = cplx::zero<T>(), t2 = cplx::zero<T>();
#pragma unroll
        for (int r = 1; r <= F; r++) {
            const T ap = win[(zz1 ? zo[F + r] : zo[F]) + pos + r * i1];
            const T am = win[(zz1 ? zo[F - r] : zo[F]) + pos - r * i1];
            t1 = cplx::fma(cplx::sub(ap, am), mc.c1[r], t1);
            if (tw) {
                const T bp = win[(zz2 ? zo[F + r] : zo[F]) + pos + r * i2];
                const T bm = win[(zz2 ? zo[F - r] : zo[F]) + pos - r * i2];
                t2 = cplx::fma(cplx::sub(bp, bm), mc.c2[r], t2);
            }
        }
        return tw ? cplx::add(t1, t2) : t1;
    };
    /* ring offsets (in elements of win) of planes kk-F .. kk+F */
    auto ring_offsets = [&](int kk, int (&zo)[2 * F + 1]) {
        const int s0 = ring_slot(kk);
        int sp = s0, sm = s0;
        zo[F] = s0 * PL;
#pragma unroll
        for (int r = 1; r <= F; r++) {
            sp = (sp + 1 == RING) ? 0 : sp + 1;
            sm = (sm == 0) ? RING - 1 : sm - 1;
            zo[F + r] = sp * PL;
            zo[F - r] = sm * PL;
        }
    };

    /* ---- 2a. z-extended derivative field of plane kk (its inner derivatives are in-plane: cell_typ 15) --- */
    auto make_dz = [&](int kk) {
        int zo[2 * F + 1];
#pragma unroll
        for (int r = 0; r <= 2 * F; r++) zo[r] = ring_slot(kk) * PL; /* only zo[F] is used: ax1, ax2 are x and y */
        dzr[ring_slot(kk) * NT + tid] = inner(qz, zo, (ly + F) * EX + lx + F);
    };

    /* ---- 2b. in-plane extended derivative fields of plane k ------------------------------------- */
    auto make_dext = [&](const int (&zo)[2 * F + 1]) {
        for (int q = 0; q < d.nmix; q++) {
            const int ext = d.mix[q].ext;
            if (ext == 2) continue;
            const int R0 = (ext == 0) ? EX : TX;            /* row length of the extended region */
            const int n = (ext == 0) ? EX * TY : TX * EY;
            T *D = dext + q * dext_elems;
            for (int e = tid; e < n; e += NT) {
                const int e0 = e % R0, e1 = e / R0;
                const int px = (ext == 0) ? e0 : e0 + F;
                const int py = (ext == 0) ? e1 + F : e1;
                D[e] = inner(q, zo, py * EX + px);
            }
        }
    };

    /* ---- prologue: planes -F .. F into the ring, planes F+1 .. F+PF on their way in registers ---------- */
    T stage[PF][NLD];
    for (int kk = -F; kk <= F; kk++) {
        fetch_plane(kk, stage[0]);
        commit_plane(kk, stage[0]);
    }
#pragma unroll
    for (int u = 0; u < PF; u++) fetch_plane(F + 1 + u, stage[u]);
    /* xprev and Veff of this thread's point travel the same way, PF planes ahead of their use */
    const int gi = x0 + lx, gj = y0 + ly;
    const bool inside = gi < Nx && gj < Ny;
    const size_t g0 = inside ? lay_pos(d.lay, gi, gj, 0) : 0;
    const bool use_xp = inside && a.s2 != 0.0, use_ve = inside && a.veff != nullptr;
    T xq[PF];
    double vq[PF];
#pragma unroll
    for (int u = 0; u < PF; u++) {
        xq[u] = cplx::zero<T>();
        vq[u] = 0.0;
        if (u < Nz) {
            if (use_xp) xq[u] = xprev[col + g0 + (size_t)u * plane_elems];
            if (use_ve) vq[u] = a.veff[g0 + (size_t)u * plane_elems];
        }
    }
    __syncthreads();
    if (qz >= 0) {
        for (int kk = -F; kk <= F; kk++) make_dz(kk);
    }

    const double diag0 = d.coef0 + a.c;
    const int p0 = (ly + F) * EX + lx + F;

    for (int kb = 0; kb < Nz; kb += PF) {
#pragma unroll
      for (int u = 0; u < PF; u++) { /* unrolled so that stage[u] stays in registers */
        const int k = kb + u;
        if (k >= Nz) break;
        /* planes k-F .. k+F are in the ring; planes k+F+1 .. k+F+PF are in flight in stage[] */
        int zo[2 * F + 1];
        ring_offsets(k, zo);
        if (d.nmix) {
            make_dext(zo);
            __syncthreads();
        }
        /* ---- 3. star stencil + potential + recurrence ---------------------------------------- */
        if (inside) {
            const T *f = win + zo[F];
            const size_t g = g0 + (size_t)k * plane_elems;
            const T xc = f[p0];
            T res = cplx::mul(xc, diag0);
#pragma unroll
            for (int r = 1; r <= F; r++) {
                T acc = cplx::mul(cplx::add(f[p0 - r], f[p0 + r]), d.wx[r]);
                acc = cplx::fma(cplx::add(f[p0 - r * EX], f[p0 + r * EX]), d.wy[r], acc);
                acc = cplx::fma(cplx::add(win[zo[F - r] + p0], win[zo[F + r] + p0]), d.wz[r], acc);
                for (int q = 0; q < d.nmix; q++) {
                    const MixedComp &mc = d.mix[q];
                    T dp, dm;
                    if (mc.ext == 2) { /* ring of tile-sized planes: slot = zo / PL */
                        dp = dzr[(zo[F + r] / PL) * NT + tid]; dm = dzr[(zo[F - r] / PL) * NT + tid];
                    } else {
                        const int st = (mc.ext == 0) ? 1 : TX;
                        const T *D = dext + q * dext_elems + ((mc.ext == 0) ? ly * EX + lx + F : (ly + F) * TX + lx);
                        dp = D[r * st]; dm = D[-r * st];
                    }
                    acc = cplx::fma(cplx::sub(dp, dm), mc.wm[r], acc);
                }
                res = cplx::add(res, acc);
            }
            if (a.veff) res = cplx::fma(xc, vq[u], res);
            T o = cplx::mul(res, a.s1);
            if (a.s2 != 0.0) o = cplx::fma(xq[u], -a.s2, o);
            out[g] = o;
            if (k + PF < Nz) {
                if (use_xp) xq[u] = xprev[col + g + (size_t)PF * plane_elems];
                if (use_ve) vq[u] = a.veff[g + (size_t)PF * plane_elems];
            }
        }
        /* plane k+F+1 (fetched PF planes ago) goes into the slot of plane k-F-1, which nobody reads any more;
           its registers take the loads of plane k+F+1+PF */
        if (k + 1 < Nz) {
            commit_plane(k + F + 1, stage[u]);
            if (k + 1 + PF < Nz) fetch_plane(k + F + 1 + PF, stage[u]);
            __syncthreads();
            if (qz >= 0) make_dz(k + F + 1); /* read by the next iteration after its barrier */
        }
      }
    }
}

template <typename T, int TX, int TY>
int launch_t(chefsi_ctx *ctx, const StepArgs &a)
{
    const StencilDesc &d = ctx->desc;
    constexpr int EX = TX + 2 * F, EY = TY + 2 * F, PL = EX * EY, NT = TX * TY;
    const int ntx = (d.Nx + TX - 1) / TX, nty = (d.Ny + TY - 1) / TY;
    bool has_z = false, has_xy = false;
    for (int q = 0; q < d.nmix; q++) {
        if (d.mix[q].ext == 2) has_z = true; else has_xy = true;
    }
    const int dext_elems = has_xy ? ((EX * TY > TX * EY) ? EX * TY : TX * EY) : 0;
    const size_t smem = sizeof(T) * ((size_t)RING * PL + 2 * (size_t)dext_elems + (has_z ? (size_t)RING * NT : 0));
    if (smem > ctx->max_smem_optin) return -2; /* complex cell_typ 15: the two rings do not fit; the brick kernel takes it */
    auto kern = stencil_zmarch_kernel<T, TX, TY>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { chefsi_fail(ctx, "cudaFuncSetAttribute(zmarch): %s", cudaGetErrorString(e)); return -1; }
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
        kern<<<dim3((unsigned)(ntx * nty), (unsigned)nc), NT, smem, ctx->stream>>>(d, b, ntx, dext_elems);
        launched++;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "z-march stencil launch: %s", cudaGetErrorString(e)); return -1; }
    return launched;
}

}  // namespace

bool stencil_zmarch_supported(const chefsi_ctx *ctx)
{
    return ctx->force_general < 2 && ctx->grid.FDn == F;
}

int launch_stencil_zmarch(chefsi_ctx *ctx, const StepArgs &a, bool is_complex)
{
    if (a.ncol <= 0) return 0;
    if (is_complex) return launch_t<double2, 16, 16>(ctx, a);
    return launch_t<double, 32, 8>(ctx, a);
}
