/*
 * chefsi_api.cu -- the C ABI of libchefsi_b200.so (include/chefsi_b200.h) and the host-side
 * orchestration of the Chebyshev filter.
 *
 * ChebyshevFiltering (eigenSolver.c:722-798) in this library is:
 *     e=(b-a)/2, c=(b+a)/2, sigma=sigma1=e/(a0-c), gamma=2/sigma1                      (:747-750)
 *     Y    = (sigma1/e) (H - c) X                                                      (:755-768)
 *     for j = 1..m-1: sigma2 = 1/(gamma-sigma)
 *                     Ynew = (2 sigma2/e)(H - c) Y - (sigma sigma2) X ; X<-Y ; Y<-Ynew  (:772-796)
 * where every "(H - c) .  scaled, minus xprev" is ONE fused stencil launch plus the nonlocal
 * projector launches, and X<-Y<-Ynew is a rotation of three device buffers (no copies).
 */
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "chefsi_internal.h"

/* ------------------------------------------------------------------------------------------ */
int chefsi_fail(chefsi_ctx *ctx, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) {
        strncpy(ctx->err, buf, sizeof(ctx->err) - 1);
        ctx->err[sizeof(ctx->err) - 1] = 0;
    }
    if (getenv("CHEFSI_B200_VERBOSE")) fprintf(stderr, "[chefsi_b200] error: %s\n", buf);
    return 1;
}

static char g_create_err[512] = "";

extern "C" const char *chefsi_version(void) { return "chefsi_b200 0.1 (sm_100a, FP64)"; }

extern "C" const char *chefsi_last_error(const chefsi_ctx_t *ctx) { return ctx ? ctx->err : g_create_err; }

extern "C" int chefsi_device_count(void)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return ndev;
}

extern "C" int chefsi_create(chefsi_ctx_t **out, int device)
{
    if (!out) return 1;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        snprintf(g_create_err, sizeof(g_create_err), "no CUDA device available (%s); this library has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return 1;
    }
    if (device < 0 || device >= ndev) {
        snprintf(g_create_err, sizeof(g_create_err), "device %d out of range (0..%d)", device, ndev - 1);
        return 1;
    }
    chefsi_ctx *ctx = new (std::nothrow) chefsi_ctx();
    if (!ctx) return 1;
    ctx->device = device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        snprintf(g_create_err, sizeof(g_create_err), "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
        delete ctx;
        return 1;
    }
    if (prop.major != 10) {
        snprintf(g_create_err, sizeof(g_create_err), "device %d is sm_%d%d; this library is built for sm_100a only",
                 device, prop.major, prop.minor);
        delete ctx;
        return 1;
    }
    ctx->num_sms = prop.multiProcessorCount;
    ctx->max_smem_optin = prop.sharedMemPerBlockOptin;
    cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 4; i++) cudaEventCreate(&ctx->ev[i]);
    ctx->force_general = getenv("CHEFSI_B200_FORCE_GENERAL") ? atoi(getenv("CHEFSI_B200_FORCE_GENERAL")) : 0;
    if (getenv("CHEFSI_B200_GRIDSYNC")) ctx->stream_gridsync = atoi(getenv("CHEFSI_B200_GRIDSYNC"));
    if (getenv("CHEFSI_B200_ALPHA_REDUCE_MIN")) ctx->alpha_reduce_min = atoi(getenv("CHEFSI_B200_ALPHA_REDUCE_MIN"));
    if (getenv("CHEFSI_B200_FAST_SMALL")) ctx->fast_small = atoi(getenv("CHEFSI_B200_FAST_SMALL"));
    if (getenv("CHEFSI_B200_GEMM_BIG_TILES")) ctx->gemm_big_tiles = atoi(getenv("CHEFSI_B200_GEMM_BIG_TILES"));
    if (getenv("CHEFSI_B200_GEMM_SYMMETRIC")) ctx->gemm_symmetric = atoi(getenv("CHEFSI_B200_GEMM_SYMMETRIC"));
    if (getenv("CHEFSI_B200_SMALL_BRICK")) ctx->small_brick = atoi(getenv("CHEFSI_B200_SMALL_BRICK"));
    if (getenv("CHEFSI_B200_NLOC_SORT")) ctx->nloc_sort = atoi(getenv("CHEFSI_B200_NLOC_SORT"));
    if (getenv("CHEFSI_B200_TMA_L2PROMO")) ctx->tma_l2promo = atoi(getenv("CHEFSI_B200_TMA_L2PROMO")) & 3;
    *out = ctx;
    return 0;
}

void chefsi_free_nloc(NlocDev &d)
{
    cudaFree(d.IP_displ); cudaFree(d.gamma); cudaFree(d.img_atom); cudaFree(d.img_ndc);
    cudaFree(d.pos_off); cudaFree(d.chiT_off); cudaFree(d.grid_pos); cudaFree(d.chiT); cudaFree(d.img_aoff);
    cudaFree(d.img_phase); cudaFree(d.atom_img_off); cudaFree(d.atom_img);
    free(d.h_img_coords);
    d = NlocDev();
}

extern "C" void chefsi_destroy(chefsi_ctx_t *ctx)
{
    if (!ctx) return;
    if (ctx->multi) { multi_destroy(ctx); return; }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    chefsi_free_nloc(ctx->nl);
    cudaFree(ctx->d_veff);
    for (int i = 0; i < 3; i++) { cudaFree(ctx->d_buf[i]); cudaFree(ctx->d_buf2[i]); cudaFree(ctx->d_buf3[i]); }
    cudaFree(ctx->d_alpha[0]);
    cudaFree(ctx->d_alpha[1]);
    cudaFree(ctx->d_alpha_sum);
    for (int i = 0; i < 3; i++) if (ctx->h_pin[i]) cudaFreeHost(ctx->h_pin[i]);
    cudaFree(ctx->d_res_Y); cudaFree(ctx->d_res_W); cudaFree(ctx->d_res_T); cudaFree(ctx->d_gemm_ws); cudaFree(ctx->d_lanczos); cudaFree(ctx->d_aar);
    for (int i = 0; i < 3; i++) cudaFree(ctx->d_small[i]);
    rayleigh_ritz_destroy(ctx);
    rank_state_destroy(ctx);
    for (int i = 0; i < 12; i++) if (ctx->pipe_ev[i]) cudaEventDestroy(ctx->pipe_ev[i]);
    cudaFree(ctx->d_sync);
    for (int i = 0; i < 4; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    delete ctx;
}

/* ---- grid ---------------------------------------------------------------------------------- */
static void update_phases(chefsi_ctx *ctx)
{
    const chefsi_grid_t &g = ctx->grid;
    for (int oz = -1; oz <= 1; oz++)
        for (int oy = -1; oy <= 1; oy++)
            for (int ox = -1; ox <= 1; ox++) {
                /* exp(i k.(ox Lx, oy Ly, oz Lz)), lapVecRoutinesKpt.c:370-382,647-660 */
                const double th = ctx->kvec[0] * (ox * g.range_x) + ctx->kvec[1] * (oy * g.range_y) +
                                  ctx->kvec[2] * (oz * g.range_z);
                const int q = (oz + 1) * 9 + (oy + 1) * 3 + (ox + 1);
                ctx->desc.ph_re[q] = cos(th);
                ctx->desc.ph_im[q] = sin(th);
            }
}

static int update_nloc_phases(chefsi_ctx *ctx)
{
    NlocDev &d = ctx->nl;
    if (d.n_img == 0) return 0;
    const chefsi_grid_t &g = ctx->grid;
    std::vector<double2> ph(d.n_img);
    for (int J = 0; J < d.n_img; J++) {
        /* nlocVecRoutines.c:911-921 */
        const double *co = d.h_img_coords + 3 * J;
        const double th = -ctx->kvec[0] * (floor(co[0] / g.range_x) * g.range_x) -
                          ctx->kvec[1] * (floor(co[1] / g.range_y) * g.range_y) -
                          ctx->kvec[2] * (floor(co[2] / g.range_z) * g.range_z);
        ph[J] = make_double2(cos(th), sin(th));
    }
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(d.img_phase, ph.data(), sizeof(double2) * d.n_img, cudaMemcpyHostToDevice, ctx->stream));
    CHEFSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int chefsi_set_grid(chefsi_ctx_t *ctx, const chefsi_grid_t *g)
{
    if (!ctx || !g) return 1;
    if (ctx->multi) return multi_set_grid(ctx, g);
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    if (g->FDn < 1 || g->FDn > CHEFSI_MAX_FDN) return chefsi_fail(ctx, "FDn %d out of range", g->FDn);
    if (!(g->cell_typ == 0 || (g->cell_typ >= 11 && g->cell_typ <= 17)))
        return chefsi_fail(ctx, "cell_typ %d not supported (0, 11..17)", g->cell_typ);
    if (g->Nx < 1 || g->Ny < 1 || g->Nz < 1) return chefsi_fail(ctx, "bad grid size");
    const int per[3] = {!g->BCx, !g->BCy, !g->BCz}, N[3] = {g->Nx, g->Ny, g->Nz};
    for (int d = 0; d < 3; d++)
        if (per[d] && N[d] < g->FDn) return chefsi_fail(ctx, "periodic axis %d has fewer points than the FD radius", d);
    ctx->grid = *g;
    ctx->Nd = (size_t)g->Nx * g->Ny * g->Nz;
    band_store_clear(ctx); /* rayleigh_ritz.cu: blocks kept for the density have the old shape */
    ctx->res_ncol = ctx->small_ncol = ctx->q_ncol = 0;
    {   /* internal layout = the reference's dense layout, ld rounded up to 16 doubles (see Layout) */
        Layout &L = ctx->lay;
        L.Nx = g->Nx; L.Ny = g->Ny; L.Nz = g->Nz;
        L.plane = (size_t)L.Nx * L.Ny;
        L.ld = (L.plane * g->Nz + 15) / 16 * 16;
        ctx->ld = L.ld;
        if (L.ld > 0x7fffffffULL) return chefsi_fail(ctx, "grid too large for 32-bit sphere indices");
    }

    StencilDesc &d = ctx->desc;
    memset(&d, 0, sizeof(d));
    d.lay = ctx->lay;
    d.Nx = g->Nx; d.Ny = g->Ny; d.Nz = g->Nz;
    d.bc[0] = g->BCx; d.bc[1] = g->BCy; d.bc[2] = g->BCz;
    d.F = g->FDn;
    const double a = -0.5; /* hamiltonianVecRoutines.c:63-83 */
    d.coef0 = (g->D2_x[0] + g->D2_y[0] + g->D2_z[0]) * a;
    for (int r = 0; r <= g->FDn; r++) {
        d.wx[r] = g->D2_x[r] * a;
        d.wy[r] = g->D2_y[r] * a;
        d.wz[r] = g->D2_z[r] * a;
    }
    /* the reference's two-stage mixed-derivative composition, lapVecRoutines.c:1210-1297 and the
       compact weight table :1337-1423 */
    auto set_mix = [&](int q, int ext, int ax1, const double *c1, int ax2, const double *c2, const double *wm) {
        MixedComp &m = d.mix[q];
        m.ext = ext; m.ax1 = ax1; m.ax2 = ax2;
        for (int r = 0; r <= g->FDn; r++) {
            m.c1[r] = c1[r];
            m.c2[r] = c2 ? c2[r] : 0.0;
            m.wm[r] = wm[r] * a;
        }
    };
    switch (g->cell_typ) {
    case 0: d.nmix = 0; break;
    case 11: d.nmix = 1; set_mix(0, 0, 1, g->D1_y, -1, nullptr, g->D2_xy); break;
    case 12: d.nmix = 1; set_mix(0, 0, 2, g->D1_z, -1, nullptr, g->D2_xz); break;
    case 13: d.nmix = 1; set_mix(0, 1, 2, g->D1_z, -1, nullptr, g->D2_yz); break;
    case 14: d.nmix = 1; set_mix(0, 0, 1, g->D1_xy, 2, g->D1_xz, g->D1_x); break;
    case 15: d.nmix = 1; set_mix(0, 2, 0, g->D1_zx, 1, g->D1_zy, g->D1_z); break;
    case 16: d.nmix = 1; set_mix(0, 1, 0, g->D1_yx, 2, g->D1_yz, g->D1_y); break;
    case 17:
        d.nmix = 2;
        set_mix(0, 0, 1, g->D1_xy, 2, g->D1_xz, g->D1_x);
        set_mix(1, 1, 2, g->D1_z, -1, nullptr, g->D2_yz);
        break;
    }
    update_phases(ctx);

    /* Veff buffer, internal layout */
    cudaFree(ctx->d_veff);
    ctx->d_veff = nullptr;
    const size_t nv = ctx->ld;
    CHEFSI_CUDA(ctx, cudaMalloc(&ctx->d_veff, nv * sizeof(double)));
    CHEFSI_CUDA(ctx, cudaMemsetAsync(ctx->d_veff, 0, nv * sizeof(double), ctx->stream));
    CHEFSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->have_veff = false;
    ctx->have_grid = true;
    ctx->res_ncol = 0;
    /* projector tables refer to grid indices: drop them */
    chefsi_free_nloc(ctx->nl);
    return 0;
}

extern "C" int chefsi_set_kpoint(chefsi_ctx_t *ctx, double k1, double k2, double k3)
{
    if (!ctx) return 1;
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (ctx->multi) return multi_set_kpoint(ctx, k1, k2, k3);
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->kvec[0] = k1; ctx->kvec[1] = k2; ctx->kvec[2] = k3;
    update_phases(ctx);
    return update_nloc_phases(ctx);
}

extern "C" int chefsi_set_veff(chefsi_ctx_t *ctx, const double *veff_host)
{
    if (!ctx) return 1;
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (ctx->multi) return multi_set_veff(ctx, veff_host);
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!veff_host) { ctx->have_veff = false; return 0; }
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(ctx->d_veff, veff_host, ctx->Nd * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CHEFSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); /* the caller may reuse veff_host right away */
    ctx->have_veff = true;
    return 0;
}

/* ---- projectors ------------------------------------------------------------------------------ */
template <typename T>
static int upload(chefsi_ctx *ctx, T **dst, const T *src, size_t n)
{
    *dst = nullptr;
    CHEFSI_CUDA(ctx, cudaMalloc((void **)dst, (n ? n : 1) * sizeof(T)));
    if (n) CHEFSI_CUDA(ctx, cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int chefsi_set_projectors(chefsi_ctx_t *ctx, const chefsi_nloc_t *nl)
{
    if (!ctx) return 1;
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (ctx->multi) return multi_set_projectors(ctx, nl);
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    chefsi_free_nloc(ctx->nl);
    if (!nl || nl->n_img == 0 || nl->n_atom == 0) return 0;
    NlocDev &d = ctx->nl;
    d.n_atom = nl->n_atom;
    d.n_img = nl->n_img;
    d.ntot = nl->IP_displ[nl->n_atom];
    d.max_nproj = 0;
    for (int a = 0; a < nl->n_atom; a++) {
        const int np = nl->IP_displ[a + 1] - nl->IP_displ[a];
        if (np > d.max_nproj) d.max_nproj = np;
    }
    const long long npos = nl->pos_off[nl->n_img];
    d.np_pad = nloc_padded_nproj(d.max_nproj);
    if (d.np_pad < 0) return chefsi_fail(ctx, "projectors: more than 32 projectors per atom not supported");
    d.total_pts = npos;
    /* validate + overlap detection (decides whether the scatter needs atomics) */
    std::vector<unsigned char> seen(ctx->Nd, 0);
    d.overlap = 0;
    for (int J = 0; J < nl->n_img; J++) {
        if (nl->img_atom[J] < 0 || nl->img_atom[J] >= nl->n_atom) return chefsi_fail(ctx, "projectors: bad atom index");
        if (nl->pos_off[J + 1] - nl->pos_off[J] != nl->img_ndc[J]) return chefsi_fail(ctx, "projectors: pos_off/ndc mismatch");
        const int np = nl->IP_displ[nl->img_atom[J] + 1] - nl->IP_displ[nl->img_atom[J]];
        if (nl->chi_off[J + 1] - nl->chi_off[J] != (long long)nl->img_ndc[J] * np)
            return chefsi_fail(ctx, "projectors: chi_off mismatch");
        for (long long i = nl->pos_off[J]; i < nl->pos_off[J + 1]; i++) {
            const int p = nl->grid_pos[i];
            if (p < 0 || (size_t)p >= ctx->Nd) return chefsi_fail(ctx, "projectors: grid_pos out of range");
            if (seen[p]) d.overlap = 1;
            seen[p] = 1;
        }
    }
    /* Segments: the projector kernels run one CTA per (sphere image, 32 columns).  A system with few atoms
       (the SCF test systems: 12-18 images) would occupy a tenth of the SMs, so its images are cut into
       segments of kSeg consecutive sphere points, each of which the kernels treat as an image of its own
       (own alpha partial; the consumer sums the partials of all segments of the atom in a fixed order).
       With >= one image per SM the images stay whole. */
    const int kSeg = (nl->n_img >= ctx->num_sms) ? 0x7fffffff : 256;
    std::vector<int> seg_img, seg_atom, seg_ndc, seg_pt0;
    for (int J = 0; J < nl->n_img; J++)
        for (int pt0 = 0; pt0 < nl->img_ndc[J]; pt0 += kSeg) {
            seg_img.push_back(J);
            seg_atom.push_back(nl->img_atom[J]);
            seg_pt0.push_back(pt0);
            seg_ndc.push_back(std::min(kSeg, nl->img_ndc[J] - pt0));
        }
    int n_seg = (int)seg_img.size();
    if (ctx->nloc_sort) {
        /* longest first: the projector kernels run one CTA per (segment, 32 columns) in list order, ~15 waves of CTAs on the
           bench workload whose periodic images come in all sizes; with the big ones in front the last wave is made of the
           small ones.  (The alpha partials of an atom are then summed in this order: fixed, so still deterministic.) */
        std::vector<int> order(n_seg);
        for (int i = 0; i < n_seg; i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return seg_ndc[x] > seg_ndc[y]; });
        std::vector<int> t_img(n_seg), t_atom(n_seg), t_ndc(n_seg), t_pt0(n_seg);
        for (int i = 0; i < n_seg; i++) {
            t_img[i] = seg_img[order[i]]; t_atom[i] = seg_atom[order[i]]; t_ndc[i] = seg_ndc[order[i]]; t_pt0[i] = seg_pt0[order[i]];
        }
        seg_img.swap(t_img); seg_atom.swap(t_atom); seg_ndc.swap(t_ndc); seg_pt0.swap(t_pt0);
    }
    d.n_img = n_seg;
    /* CSR atom -> segments */
    std::vector<int> off(nl->n_atom + 1, 0), lst(n_seg);
    for (int s2 = 0; s2 < n_seg; s2++) off[seg_atom[s2] + 1]++;
    d.max_parts = 0;
    for (int a = 0; a < nl->n_atom; a++) d.max_parts = std::max(d.max_parts, off[a + 1]);
    for (int a = 0; a < nl->n_atom; a++) off[a + 1] += off[a];
    {
        std::vector<int> cur(off.begin(), off.end() - 1);
        for (int s2 = 0; s2 < n_seg; s2++) lst[cur[seg_atom[s2]]++] = s2;
    }
    if (upload(ctx, &d.IP_displ, nl->IP_displ, (size_t)nl->n_atom + 1)) return 1;
    if (upload(ctx, &d.gamma, nl->gamma, (size_t)d.ntot)) return 1;
    if (upload(ctx, &d.img_atom, seg_atom.data(), (size_t)n_seg)) return 1;
    if (upload(ctx, &d.img_ndc, seg_ndc.data(), (size_t)n_seg)) return 1;
    {
        std::vector<long long> poff((size_t)n_seg + 1, 0);
        for (int s2 = 0; s2 < n_seg; s2++) poff[s2] = nl->pos_off[seg_img[s2]] + seg_pt0[s2];
        poff[n_seg] = npos;
        if (upload(ctx, &d.pos_off, poff.data(), poff.size())) return 1;
    }
    if (upload(ctx, &d.grid_pos, nl->grid_pos, (size_t)npos)) return 1; /* internal layout == the reference's: indices as they are */
    {   /* Chi per image transposed to point-major and zero-padded to np_pad projectors (what the nloc
           kernel streams with 16-byte cp.async); per-segment offsets into it and of the alpha partials */
        std::vector<long long> toff((size_t)nl->n_img + 1, 0);
        for (int J = 0; J < nl->n_img; J++) toff[J + 1] = toff[J] + (long long)nl->img_ndc[J] * d.np_pad;
        std::vector<double> chiT((size_t)toff[nl->n_img], 0.0);
        for (int J = 0; J < nl->n_img; J++) {
            const int np = nl->IP_displ[nl->img_atom[J] + 1] - nl->IP_displ[nl->img_atom[J]];
            const int ndc = nl->img_ndc[J];
            const double *src = nl->chi + nl->chi_off[J];
            double *dst = chiT.data() + toff[J];
            for (int p = 0; p < np; p++)
                for (int i = 0; i < ndc; i++) dst[(size_t)i * d.np_pad + p] = src[(size_t)p * ndc + i];
        }
        std::vector<long long> soff((size_t)n_seg + 1, 0);
        std::vector<int> aoff((size_t)n_seg + 1, 0);
        for (int s2 = 0; s2 < n_seg; s2++) {
            const int np = nl->IP_displ[seg_atom[s2] + 1] - nl->IP_displ[seg_atom[s2]];
            soff[s2] = toff[seg_img[s2]] + (long long)seg_pt0[s2] * d.np_pad;
            aoff[s2 + 1] = aoff[s2] + np;
        }
        soff[n_seg] = toff[nl->n_img];
        d.img_proj_total = aoff[n_seg];
        if (upload(ctx, &d.chiT_off, soff.data(), soff.size())) return 1;
        if (upload(ctx, &d.img_aoff, aoff.data(), aoff.size())) return 1;
        if (upload(ctx, &d.chiT, chiT.data(), chiT.size())) return 1;
        d.n_chiT = (long long)chiT.size();
    }
    if (upload(ctx, &d.atom_img_off, off.data(), off.size())) return 1;
    if (upload(ctx, &d.atom_img, lst.data(), lst.size())) return 1;
    CHEFSI_CUDA(ctx, cudaMalloc((void **)&d.img_phase, sizeof(double2) * n_seg));
    d.h_img_coords = (double *)malloc(sizeof(double) * 3 * n_seg);
    for (int s2 = 0; s2 < n_seg; s2++) memcpy(d.h_img_coords + 3 * s2, nl->img_coords + 3 * seg_img[s2], sizeof(double) * 3);
    return update_nloc_phases(ctx);
}

/* ---- one fused step ---------------------------------------------------------------------------- */
struct EvPair { cudaEvent_t a, b; int kind; };

struct Profiler {
    chefsi_ctx *ctx;
    std::vector<EvPair> evs;
    explicit Profiler(chefsi_ctx *c) : ctx(c) {}
    void begin(int kind)
    {
        if (!ctx->profiling) return;
        EvPair p;
        cudaEventCreate(&p.a);
        cudaEventCreate(&p.b);
        p.kind = kind;
        cudaEventRecord(p.a, ctx->stream);
        evs.push_back(p);
    }
    void end()
    {
        if (!ctx->profiling) return;
        cudaEventRecord(evs.back().b, ctx->stream);
    }
    void finish()
    {
        if (!ctx->profiling) return;
        cudaStreamSynchronize(ctx->stream);
        double st = 0, nl = 0;
        int ns = 0;
        for (auto &p : evs) {
            float ms = 0;
            cudaEventElapsedTime(&ms, p.a, p.b);
            if (p.kind == 0) { st += ms; ns++; } else nl += ms;
            cudaEventDestroy(p.a);
            cudaEventDestroy(p.b);
        }
        ctx->stats.last_stencil_ms = st;
        ctx->stats.last_nloc_ms = nl;
        ctx->stats.last_stencil_launches = ns;
        evs.clear();
    }
};

/* One fused step  out = s1 * ((H_local + c) x + Vnl x) - s2 * xprev.
 * nl_in : the alpha partials of x are already in ctx->d_alpha[cur] (left there by the previous step's
 *         fused projector kernel); otherwise they are computed here from x.
 * nl_out: after adding Vnl x, also project `out` for the next step (only legal when spheres are disjoint). */
static int apply_step(chefsi_ctx *ctx, Profiler &prof, const void *x, const void *xprev, void *out, int ncol, double c,
                      double s1, double s2, bool is_complex, bool nl_in, bool nl_out, bool with_nl = true, bool with_veff = true)
{
    StepArgs a;
    a.x = x; a.xprev = xprev; a.out = out;
    a.veff = (with_veff && ctx->have_veff) ? ctx->d_veff : nullptr;
    a.ld = ctx->ld; a.ncol = ncol; a.c = c; a.s1 = s1; a.s2 = s2;
    int n;
    const bool have_nl = with_nl && ctx->nl.n_img > 0 && ctx->nl.ntot > 0;
    if (have_nl && !nl_in) {
        prof.begin(1);
        n = launch_nloc(ctx, NLOC_PROJECT, const_cast<void *>(x), ctx->ld, ncol, 0.0, is_complex);
        prof.end();
        if (n < 0) return 1;
        ctx->stats.kernel_launches += n;
    }
    prof.begin(0);
    if (stream_dense_supported(ctx, is_complex)) {
        n = launch_stencil_stream_dense(ctx, a);
        ctx->stats.last_path = 1;
    } else if (is_complex && stream_kpt_supported(ctx)) {
        n = launch_stencil_stream_kpt(ctx, a);
        ctx->stats.last_path = 1;
    } else {
        n = -2;
        if (stream_mixed_supported(ctx, is_complex)) {
            n = launch_stencil_stream_mixed(ctx, a); /* -2: tables without the structure the kernel folds */
            ctx->stats.last_path = 3;
        }
        /* a launch with fewer z-marching CTAs than SMs (single columns of Lanczos / the AAR iteration on the SCF test
           systems: 4-15 CTAs that walk the whole z extent) is latency-bound; the 3-D brick kernel cuts z as well
           (Au_fcc211: 125 bricks per column) */
        if (n == -2 && ctx->small_brick && ctx->force_general < 2) {
            const int tx = is_complex ? 16 : 32, ty = is_complex ? 16 : 8;
            const long long ctas = (long long)ncol * ((ctx->grid.Nx + tx - 1) / tx) * ((ctx->grid.Ny + ty - 1) / ty);
            if (ctas * 2 < ctx->num_sms) {
                n = launch_stencil_general(ctx, a, is_complex);
                ctx->stats.last_path = 0;
            }
        }
        if (n == -2 && stencil_zmarch_supported(ctx)) {
            n = launch_stencil_zmarch(ctx, a, is_complex); /* -2: this case does not fit (shared memory) */
            ctx->stats.last_path = 2;
        }
        if (n == -2) {
            n = launch_stencil_general(ctx, a, is_complex);
            ctx->stats.last_path = 0;
        }
    }
    prof.end();
    if (n < 0) return 1;
    ctx->stats.kernel_launches += n;
    if (have_nl) {
        prof.begin(1);
        n = launch_nloc(ctx, nl_out ? NLOC_FUSED : NLOC_EXPAND, out, ctx->ld, ncol, s1, is_complex);
        prof.end();
        if (n < 0) return 1;
        ctx->stats.kernel_launches += n;
    }
    return 0;
}

static const char kMultiDeviceApi[] = "device-resident entry points take a single-device context (a multi-device context splits HOST blocks)";

static int filter_device(chefsi_ctx *ctx, void *bufs[3], int ncol, int m, double a, double b, double a0, bool is_complex,
                         int *y_slot, int *x_slot)
{
    if (ctx->multi) return chefsi_fail(ctx, kMultiDeviceApi);
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (m < 1) return chefsi_fail(ctx, "Chebyshev degree must be >= 1");
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    Profiler prof(ctx);
    const double e = 0.5 * (b - a);
    const double c = 0.5 * (b + a);
    double sigma = e / (a0 - c);
    const double sigma1 = sigma;
    const double gamma = 2.0 / sigma1;
    int X = 0, Y = 1, W = 2;
    cudaEventRecord(ctx->ev[0], ctx->stream);
    /* with disjoint spheres every step's projector kernel also projects its (final) output, so the
       next step starts with alpha in hand: one stencil launch + one projector launch per degree */
    const bool chain = !ctx->nl.overlap;
    if (apply_step(ctx, prof, bufs[X], nullptr, bufs[Y], ncol, -c, sigma1 / e, 0.0, is_complex, false, chain && m > 1)) return 1;
    for (int j = 1; j < m; j++) {
        const double sigma2 = 1.0 / (gamma - sigma);
        if (apply_step(ctx, prof, bufs[Y], bufs[X], bufs[W], ncol, -c, 2.0 * sigma2 / e, sigma * sigma2, is_complex,
                       chain, chain && j < m - 1)) return 1;
        const int t = X; X = Y; Y = W; W = t;
        sigma = sigma2;
    }
    cudaEventRecord(ctx->ev[1], ctx->stream);
    prof.finish();
    if (y_slot) *y_slot = Y;
    if (x_slot) *x_slot = X;
    return 0;
}

extern "C" size_t chefsi_device_ld(const chefsi_ctx_t *ctx) { return ctx ? ctx->ld : 0; }

extern "C" int chefsi_chebyshev_filter_device(chefsi_ctx_t *ctx, double *bufA, double *bufB, double *bufC, int ncol, int m,
                                              double a, double b, double a0, int *y_slot, int *x_slot)
{
    if (!ctx) return 1;
    void *bufs[3] = {bufA, bufB, bufC};
    return filter_device(ctx, bufs, ncol, m, a, b, a0, false, y_slot, x_slot);
}
extern "C" int chefsi_chebyshev_filter_kpt_device(chefsi_ctx_t *ctx, void *bufA, void *bufB, void *bufC, int ncol, int m,
                                                  double a, double b, double a0, int *y_slot, int *x_slot)
{
    if (!ctx) return 1;
    void *bufs[3] = {bufA, bufB, bufC};
    return filter_device(ctx, bufs, ncol, m, a, b, a0, true, y_slot, x_slot);
}

static int hmult_device(chefsi_ctx *ctx, int ncol, double c, const void *x, void *Hx, bool is_complex)
{
    if (ctx->multi) return chefsi_fail(ctx, kMultiDeviceApi);
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    Profiler prof(ctx);
    const int rc = apply_step(ctx, prof, x, nullptr, Hx, ncol, c, 1.0, 0.0, is_complex, false, false);
    prof.finish();
    return rc;
}
/* r = b + (Lap + c) x for one resident real column: the residual of aar.cu (poisson_residual, lapVecRoutines.c:61-79).
   The stencil kernels evaluate s1 ((-1/2 Lap + c') x) - s2 xprev: s1 = -2, c' = c / s1, xprev = b, s2 = -1. */
int lap_residual_device(chefsi_ctx *ctx, double c, const void *x, const void *b, void *r)
{
    Profiler prof(ctx);
    const int rc = apply_step(ctx, prof, x, b, r, 1, c / -2.0, -2.0, -1.0, false, false, false, /*with_nl=*/false, /*with_veff=*/false);
    prof.finish();
    return rc;
}

/* one H apply (c = 0) of a single resident column: the operator of lanczos.cu */
int apply_h_device(chefsi_ctx *ctx, const void *x, void *Hx, bool is_complex) { return hmult_device(ctx, 1, 0.0, x, Hx, is_complex); }

extern "C" int chefsi_hamiltonian_mult_device(chefsi_ctx_t *ctx, int ncol, double c, const double *x, double *Hx)
{
    return ctx ? hmult_device(ctx, ncol, c, x, Hx, false) : 1;
}
extern "C" int chefsi_hamiltonian_mult_kpt_device(chefsi_ctx_t *ctx, int ncol, double c, const void *x, void *Hx)
{
    return ctx ? hmult_device(ctx, ncol, c, x, Hx, true) : 1;
}

extern "C" int chefsi_synchronize(chefsi_ctx_t *ctx)
{
    if (!ctx) return 1;
    if (ctx->multi) return multi_synchronize(ctx);
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    CHEFSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess) ctx->stats.last_filter_ms = ms;
    else cudaGetLastError();
    return 0;
}

/* ---- host-buffer entry points ----------------------------------------------------------------
 * What the reference's ChebyshevFiltering / Hamiltonian_vectors_mult bind to: X and Y live in host
 * memory (SPARC's Xorb / Yorb).  The block is cut into column chunks that flow through a three-stage
 * pipeline on three streams -- H2D copy of chunk k+1, filter of chunk k, D2H copy of chunk k-1 -- so
 * the call costs about max(PCIe in, compute, PCIe out) instead of their sum.  The device block IS the
 * reference's layout (columns ld apart), so the copies go straight into / out of the recurrence buffers;
 * chunks rotate through three buffer trios (with two, the chain D2H(k) -> H2D(k+2) -> filter(k+2) would
 * expose a copy per chunk). */
static int ensure_bufs(chefsi_ctx *ctx, size_t bytes_each)
{
    if (bytes_each <= ctx->buf_bytes) return 0;
    for (int i = 0; i < 3; i++) { cudaFree(ctx->d_buf[i]); ctx->d_buf[i] = nullptr; }
    ctx->buf_bytes = 0;
    for (int i = 0; i < 3; i++) CHEFSI_CUDA(ctx, cudaMalloc(&ctx->d_buf[i], bytes_each));
    ctx->buf_bytes = bytes_each;
    return 0;
}

static int ensure_bufs2(chefsi_ctx *ctx, size_t bytes_each)
{
    if (bytes_each <= ctx->buf2_bytes) return 0;
    for (int i = 0; i < 3; i++) {
        cudaFree(ctx->d_buf2[i]); ctx->d_buf2[i] = nullptr;
        cudaFree(ctx->d_buf3[i]); ctx->d_buf3[i] = nullptr;
    }
    ctx->buf2_bytes = 0;
    for (int i = 0; i < 3; i++) {
        CHEFSI_CUDA(ctx, cudaMalloc(&ctx->d_buf2[i], bytes_each));
        CHEFSI_CUDA(ctx, cudaMalloc(&ctx->d_buf3[i], bytes_each));
    }
    ctx->buf2_bytes = bytes_each;
    return 0;
}

/* columns per chunk: small enough that the pipeline has ~8 chunks to overlap and that three buffer trios and
 * alpha fit in the free device memory; large enough to fill the GPU */
static int chunk_columns(chefsi_ctx *ctx, int ncol, size_t esz)
{
    /* a small block that fits the buffers this context already holds: no driver query (cudaMemGetInfo costs more
       than a single-column H apply on an SCF-test-sized grid) */
    if (ctx->fast_small && (size_t)ncol * ctx->Nd * esz <= ((size_t)64 << 20) && (size_t)ncol * ctx->ld * esz <= ctx->buf_bytes)
        return ncol;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = (size_t)8 << 30; }
    free_b += 3 * ctx->buf_bytes + 6 * ctx->buf2_bytes + 2 * ctx->alpha_bytes; /* what we already hold can be reused */
    const size_t per_col = 9 * ctx->ld * esz + 2 * (size_t)ctx->nl.img_proj_total * esz;
    size_t budget = (size_t)(0.85 * (double)free_b);
    const char *env = getenv("CHEFSI_B200_MAX_CHUNK_BYTES");
    if (env) { size_t v = strtoull(env, nullptr, 10); if (v && v < budget) budget = v; }
    size_t n = budget / (per_col ? per_col : 1);
    if (n < 1) n = 1;
    if (n > (size_t)ncol) n = (size_t)ncol;
    /* pipeline granularity: only worth it when the block is big enough for the copies to matter */
    const size_t block_bytes = (size_t)ncol * ctx->Nd * esz;
    if (block_bytes > ((size_t)64 << 20)) {
        size_t want = ((size_t)ncol + 7) / 8;
        if (want < 32) want = 32;
        if (want > 128) want = 128;
        const char *e2 = getenv("CHEFSI_B200_HOST_CHUNK");
        if (e2 && atoi(e2) > 0) want = (size_t)atoi(e2);
        if (want < n) n = want;
    }
    return (int)n;
}

static int ensure_pipe_events(chefsi_ctx *ctx)
{
    if (ctx->pipe_ev[0]) return 0;
    for (int i = 0; i < 12; i++) CHEFSI_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pipe_ev[i], cudaEventDisableTiming));
    return 0;
}

/* ---- pinned staging of small pageable blocks ---------------------------------------------------------------- */
static const size_t kStageMax = (size_t)16 << 20;

static bool host_is_pageable(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}
/* true when the call should stage: small block, pageable caller memory */
static bool want_staging(chefsi_ctx *ctx, const void *h_in, const void *h_out, size_t bytes)
{
    if (!ctx->fast_small || bytes == 0 || bytes > kStageMax) return false;
    if (!host_is_pageable(h_in) && !host_is_pageable(h_out)) return false;
    if (ctx->h_pin_bytes < kStageMax) {
        for (int i = 0; i < 3; i++)
            if (cudaHostAlloc(&ctx->h_pin[i], kStageMax, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); ctx->h_pin[i] = nullptr; return false; }
        ctx->h_pin_bytes = kStageMax;
    }
    return true;
}
/* host block (columns ldh_bytes apart) -> device block (columns pitch apart) through pinned slot `slot`, on ctx->stream */
static int staged_h2d(chefsi_ctx *ctx, int slot, void *dev, size_t pitch, const void *host, size_t ldh_bytes, size_t row, int ncol)
{
    char *pin = (char *)ctx->h_pin[slot];
    for (int n = 0; n < ncol; n++) memcpy(pin + (size_t)n * row, (const char *)host + (size_t)n * ldh_bytes, row);
    if (pitch == row || ncol == 1) CHEFSI_CUDA(ctx, cudaMemcpyAsync(dev, pin, row * ncol, cudaMemcpyHostToDevice, ctx->stream));
    else CHEFSI_CUDA(ctx, cudaMemcpy2DAsync(dev, pitch, pin, row, row, ncol, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
static int staged_d2h_begin(chefsi_ctx *ctx, int slot, const void *dev, size_t pitch, size_t row, int ncol)
{
    char *pin = (char *)ctx->h_pin[slot];
    if (pitch == row || ncol == 1) CHEFSI_CUDA(ctx, cudaMemcpyAsync(pin, dev, row * ncol, cudaMemcpyDeviceToHost, ctx->stream));
    else CHEFSI_CUDA(ctx, cudaMemcpy2DAsync(pin, row, dev, pitch, row, ncol, cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}
static void staged_d2h_finish(chefsi_ctx *ctx, int slot, void *host, size_t ldh_bytes, size_t row, int ncol)
{
    const char *pin = (const char *)ctx->h_pin[slot];
    for (int n = 0; n < ncol; n++) memcpy((char *)host + (size_t)n * ldh_bytes, pin + (size_t)n * row, row);
}

/* after a failure in the middle of a pipelined call: no async copy may still be reading or writing the caller's
 * buffers when the call returns (the caller may free them), and the streams must be joinable for the next call */
static int drain_streams(chefsi_ctx *ctx, int rc)
{
    cudaStreamSynchronize(ctx->h2d_stream);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->d2h_stream);
    cudaGetLastError();
    return rc;
}
#define CHEFSI_CUDA_DRAIN(ctx, call)                                                                      \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return drain_streams((ctx), chefsi_fail((ctx), "%s:%d: %s -> %s", __FILE__, __LINE__, #call,   \
                                                    cudaGetErrorString(e_)));                             \
    } while (0)

static int filter_host(chefsi_ctx *ctx, void *X, size_t ldi, void *Y, size_t ldo, int ncol, int m, double a, double b,
                       double a0, int flags, bool is_complex)
{
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (ncol <= 0) return 0;
    if (ldi < ctx->Nd || ldo < ctx->Nd) return chefsi_fail(ctx, "leading dimension smaller than the grid");
    if (ctx->multi) return multi_filter_host(ctx, X, ldi, Y, ldo, ncol, m, a, b, a0, flags, is_complex);
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    band_store_invalidate(ctx, X); /* the host blocks change: their device copies for the density are stale */
    band_store_invalidate(ctx, Y);
    const size_t esz = is_complex ? 2 * sizeof(double) : sizeof(double);
    const int chunk = chunk_columns(ctx, ncol, esz);
    if (ensure_bufs(ctx, (size_t)chunk * ctx->ld * esz)) return 1;
    /* KEEP_Y: the result also goes into the resident block of the subspace routines (subspace.cu) */
    const bool keep_y = (flags & CHEFSI_FLAG_KEEP_Y) && ctx->d_res_Y && (size_t)ncol * ctx->ld * esz <= ctx->res_bytes;
    const bool skip_y = keep_y && (flags & CHEFSI_FLAG_NO_Y_COPYBACK);
    ctx->res_ncol = 0;
    ctx->res_unwritten_host = skip_y ? Y : nullptr;
    if (chunk == ncol && want_staging(ctx, X, Y, (size_t)ncol * ctx->Nd * esz)) {
        /* one small chunk in pageable memory: pinned staging, one stream, one synchronisation */
        const size_t row = ctx->Nd * esz, pitch = ctx->ld * esz;
        CHEFSI_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
        if (staged_h2d(ctx, 0, ctx->d_buf[0], pitch, X, ldi * esz, row, ncol)) return drain_streams(ctx, 1);
        int ys = 1, xs = 0;
        if (filter_device(ctx, ctx->d_buf, ncol, m, a, b, a0, is_complex, &ys, &xs)) return drain_streams(ctx, 1);
        if (keep_y) CHEFSI_CUDA_DRAIN(ctx, cudaMemcpyAsync(ctx->d_res_Y, ctx->d_buf[ys], (size_t)ncol * pitch, cudaMemcpyDeviceToDevice, ctx->stream));
        if (!skip_y && staged_d2h_begin(ctx, 1, ctx->d_buf[ys], pitch, row, ncol)) return drain_streams(ctx, 1);
        if (!(flags & CHEFSI_FLAG_NO_X_COPYBACK) && staged_d2h_begin(ctx, 2, ctx->d_buf[xs], pitch, row, ncol)) return drain_streams(ctx, 1);
        CHEFSI_CUDA_DRAIN(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
        CHEFSI_CUDA_DRAIN(ctx, cudaStreamSynchronize(ctx->stream));
        if (!skip_y) staged_d2h_finish(ctx, 1, Y, ldo * esz, row, ncol);
        if (!(flags & CHEFSI_FLAG_NO_X_COPYBACK)) staged_d2h_finish(ctx, 2, X, ldi * esz, row, ncol);
        if (keep_y) { ctx->res_ncol = ncol; ctx->res_host = Y; ctx->res_complex = is_complex; }
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]) == cudaSuccess) ctx->stats.last_filter_ms = ms; else cudaGetLastError();
        return 0;
    }
    if (chunk < ncol && ensure_bufs2(ctx, (size_t)chunk * ctx->ld * esz)) return 1;
    if (ensure_pipe_events(ctx)) return 1;
    cudaEvent_t *ev_h2d = ctx->pipe_ev, *ev_out = ctx->pipe_ev + 6, *ev_d2h = ctx->pipe_ev + 9;
    const bool copy_x = !(flags & CHEFSI_FLAG_NO_X_COPYBACK);
    const size_t row = ctx->Nd * esz, pitch = ctx->ld * esz;
    cudaEvent_t t0 = ctx->ev[2], t1 = ctx->ev[3];
    CHEFSI_CUDA(ctx, cudaEventRecord(t0, ctx->stream));
    CHEFSI_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d_stream, t0, 0)); /* order after earlier work of this context */
    /* Chunk schedule: the first H2D and the last D2H are the only copies nothing overlaps with, so the
       pipeline ramps up and down through quarter- and half-size chunks (c/4, c/2, c, ..., c, c/2, c/4). */
    std::vector<int> sched;
    {
        const int q = chunk / 4, h = chunk / 2;
        /* only for long pipelines: a chunk of < 32 columns costs the projector kernel as much as 32 (it works on
           groups of 32 columns), so on a short block the ramp costs more than the copies it hides (measured at
           128 columns per rank: 8.1e9 flat vs 5.7e9 ramped at N = 2) */
        const char *sched_env = getenv("CHEFSI_B200_CHUNK_SCHED"); /* experiments: explicit "c0,c1,..." (the last entry repeats) */
        if (sched_env && *sched_env) {
            std::vector<int> pat;
            for (const char *p2 = sched_env; *p2;) {
                const int v = atoi(p2);
                if (v > 0) pat.push_back(v < chunk ? v : chunk);
                while (*p2 && *p2 != ',') p2++;
                if (*p2 == ',') p2++;
            }
            int rest = ncol;
            for (size_t i = 0; rest > 0 && !pat.empty(); i++) {
                const int c1 = pat[i < pat.size() ? i : pat.size() - 1];
                sched.push_back(rest < c1 ? rest : c1);
                rest -= sched.back();
            }
        }
        if (!sched.empty()) {
        } else if (ncol >= 6 * chunk && q >= 8 && getenv("CHEFSI_B200_FLAT_CHUNKS") == nullptr) {
            int rest = ncol - 2 * (q + h);
            sched.push_back(q); sched.push_back(h);
            while (rest > 0) { const int c1 = rest < chunk ? rest : chunk; sched.push_back(c1); rest -= c1; }
            sched.push_back(h); sched.push_back(q);
        } else {
            for (int c0 = 0; c0 < ncol; c0 += chunk) sched.push_back((ncol - c0 < chunk) ? ncol - c0 : chunk);
        }
    }
    int c0 = 0, k = 0;
    for (size_t kk = 0; kk < sched.size(); c0 += sched[kk], kk++, k++) {
        const int nc = sched[kk];
        const int s = k % 3;
        void **trio = s == 0 ? ctx->d_buf : (s == 1 ? ctx->d_buf2 : ctx->d_buf3);
        /* trio s is free again when the copies of chunk k-3 out of it have finished */
        if (k >= 3) CHEFSI_CUDA_DRAIN(ctx, cudaStreamWaitEvent(ctx->h2d_stream, ev_d2h[s], 0));
        CHEFSI_CUDA_DRAIN(ctx, cudaMemcpy2DAsync(trio[0], pitch, (const char *)X + (size_t)c0 * ldi * esz, ldi * esz, row, nc,
                                                 cudaMemcpyHostToDevice, ctx->h2d_stream));
        CHEFSI_CUDA_DRAIN(ctx, cudaEventRecord(ev_h2d[s], ctx->h2d_stream));
        CHEFSI_CUDA_DRAIN(ctx, cudaStreamWaitEvent(ctx->stream, ev_h2d[s], 0));
        int ys = 1, xs = 0;
        if (filter_device(ctx, trio, nc, m, a, b, a0, is_complex, &ys, &xs)) return drain_streams(ctx, 1);
        if (keep_y)
            CHEFSI_CUDA_DRAIN(ctx, cudaMemcpyAsync((char *)ctx->d_res_Y + (size_t)c0 * pitch, trio[ys], (size_t)nc * pitch,
                                                   cudaMemcpyDeviceToDevice, ctx->stream));
        CHEFSI_CUDA_DRAIN(ctx, cudaEventRecord(ev_out[s], ctx->stream));
        CHEFSI_CUDA_DRAIN(ctx, cudaStreamWaitEvent(ctx->d2h_stream, ev_out[s], 0));
        if (!skip_y)
            CHEFSI_CUDA_DRAIN(ctx, cudaMemcpy2DAsync((char *)Y + (size_t)c0 * ldo * esz, ldo * esz, trio[ys], pitch, row, nc,
                                                     cudaMemcpyDeviceToHost, ctx->d2h_stream));
        if (copy_x)
            CHEFSI_CUDA_DRAIN(ctx, cudaMemcpy2DAsync((char *)X + (size_t)c0 * ldi * esz, ldi * esz, trio[xs], pitch, row, nc,
                                                     cudaMemcpyDeviceToHost, ctx->d2h_stream));
        CHEFSI_CUDA_DRAIN(ctx, cudaEventRecord(ev_d2h[s], ctx->d2h_stream));
    }
    /* join: the compute stream waits for the last copies, then the host waits for it */
    for (int s = 0; s < 3 && s < k; s++) CHEFSI_CUDA_DRAIN(ctx, cudaStreamWaitEvent(ctx->stream, ev_d2h[s], 0));
    CHEFSI_CUDA_DRAIN(ctx, cudaEventRecord(t1, ctx->stream));
    CHEFSI_CUDA_DRAIN(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, t0, t1) == cudaSuccess) ctx->stats.last_filter_ms = ms; else cudaGetLastError();
    if (keep_y) { ctx->res_ncol = ncol; ctx->res_host = Y; ctx->res_complex = is_complex; }
    return 0;
}

extern "C" int chefsi_chebyshev_filter(chefsi_ctx_t *ctx, double *X, size_t ldi, double *Y, size_t ldo, int ncol, int m,
                                       double a, double b, double a0, int flags)
{
    return ctx ? filter_host(ctx, X, ldi, Y, ldo, ncol, m, a, b, a0, flags, false) : 1;
}
extern "C" int chefsi_chebyshev_filter_kpt(chefsi_ctx_t *ctx, void *X, size_t ldi, void *Y, size_t ldo, int ncol, int m,
                                           double a, double b, double a0, int flags)
{
    return ctx ? filter_host(ctx, X, ldi, Y, ldo, ncol, m, a, b, a0, flags, true) : 1;
}

static int hmult_host(chefsi_ctx *ctx, int ncol, double c, const void *x, size_t ldi, void *Hx, size_t ldo, bool is_complex)
{
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (ncol <= 0) return 0;
    if (ldi < ctx->Nd || ldo < ctx->Nd) return chefsi_fail(ctx, "leading dimension smaller than the grid");
    if (ctx->multi) return multi_hmult_host(ctx, ncol, c, x, ldi, Hx, ldo, is_complex);
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t esz = is_complex ? 2 * sizeof(double) : sizeof(double);
    const int chunk = chunk_columns(ctx, ncol, esz);
    if (ensure_bufs(ctx, (size_t)chunk * ctx->ld * esz)) return 1;
    const size_t row = ctx->Nd * esz, pitch = ctx->ld * esz;
    if (chunk == ncol && want_staging(ctx, x, Hx, (size_t)ncol * row)) {
        if (staged_h2d(ctx, 0, ctx->d_buf[0], pitch, x, ldi * esz, row, ncol)) return drain_streams(ctx, 1);
        if (hmult_device(ctx, ncol, c, ctx->d_buf[0], ctx->d_buf[1], is_complex)) return drain_streams(ctx, 1);
        if (staged_d2h_begin(ctx, 1, ctx->d_buf[1], pitch, row, ncol)) return drain_streams(ctx, 1);
        CHEFSI_CUDA_DRAIN(ctx, cudaStreamSynchronize(ctx->stream));
        staged_d2h_finish(ctx, 1, Hx, ldo * esz, row, ncol);
        return 0;
    }
    for (int c0 = 0; c0 < ncol; c0 += chunk) {
        const int nc = (ncol - c0 < chunk) ? ncol - c0 : chunk;
        CHEFSI_CUDA_DRAIN(ctx, cudaMemcpy2DAsync(ctx->d_buf[0], pitch, (const char *)x + (size_t)c0 * ldi * esz, ldi * esz, row, nc,
                                                 cudaMemcpyHostToDevice, ctx->stream));
        if (hmult_device(ctx, nc, c, ctx->d_buf[0], ctx->d_buf[1], is_complex)) return drain_streams(ctx, 1);
        CHEFSI_CUDA_DRAIN(ctx, cudaMemcpy2DAsync((char *)Hx + (size_t)c0 * ldo * esz, ldo * esz, ctx->d_buf[1], pitch, row, nc,
                                                 cudaMemcpyDeviceToHost, ctx->stream));
    }
    /* one synchronisation per call: the chunks reuse d_buf in stream order */
    CHEFSI_CUDA_DRAIN(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int chefsi_hamiltonian_mult(chefsi_ctx_t *ctx, int ncol, double c, const double *x, size_t ldi, double *Hx,
                                       size_t ldo)
{
    return ctx ? hmult_host(ctx, ncol, c, x, ldi, Hx, ldo, false) : 1;
}
extern "C" int chefsi_hamiltonian_mult_kpt(chefsi_ctx_t *ctx, int ncol, double c, const void *x, size_t ldi, void *Hx,
                                           size_t ldo)
{
    return ctx ? hmult_host(ctx, ncol, c, x, ldi, Hx, ldo, true) : 1;
}

/* ---- Rayleigh-Ritz projection / subspace rotation on the resident block (kernels: subspace.cu) ------------------- */
static int subspace_reserve(chefsi_ctx *ctx, int ncol, bool is_complex)
{
    if (ctx->multi) return multi_subspace_reserve(ctx, ncol, is_complex);
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (ncol <= 0) return chefsi_fail(ctx, "subspace_reserve: ncol must be positive");
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t need = (size_t)ncol * ctx->ld * sizeof(double) * (is_complex ? 2 : 1);
    if (need > ctx->res_bytes) {
        cudaFree(ctx->d_res_Y); cudaFree(ctx->d_res_W);
        ctx->d_res_Y = ctx->d_res_W = nullptr;
        ctx->res_bytes = 0;
        ctx->res_ncol = 0;
        if (cudaMalloc(&ctx->d_res_Y, need) != cudaSuccess || cudaMalloc(&ctx->d_res_W, need) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(ctx->d_res_Y); cudaFree(ctx->d_res_W);
            ctx->d_res_Y = ctx->d_res_W = nullptr;
            return chefsi_fail(ctx, "subspace_reserve: two blocks of %d columns (%zu bytes each) do not fit on the device", ncol, need);
        }
        ctx->res_bytes = need;
    }
    if (is_complex && need > ctx->res_t_bytes) { /* third block: -i H Y, then i Y */
        cudaFree(ctx->d_res_T);
        ctx->d_res_T = nullptr;
        ctx->res_t_bytes = 0;
        if (cudaMalloc(&ctx->d_res_T, need) != cudaSuccess) {
            cudaGetLastError();
            return chefsi_fail(ctx, "subspace_reserve: the third block of %d complex columns does not fit on the device", ncol);
        }
        ctx->res_t_bytes = need;
    }
    /* Hp, Mp, Q: ncol x ncol (complex: twice that; Q split into real and imaginary parts) */
    const size_t sm = (size_t)ncol * ncol * sizeof(double) * (is_complex ? 2 : 1);
    if (sm > ctx->small_bytes) {
        for (int i = 0; i < 3; i++) { cudaFree(ctx->d_small[i]); ctx->d_small[i] = nullptr; }
        ctx->small_bytes = 0;
        ctx->small_ncol = ctx->q_ncol = 0;
        for (int i = 0; i < 3; i++) CHEFSI_CUDA(ctx, cudaMalloc(&ctx->d_small[i], sm));
        ctx->small_bytes = sm;
    }
    return 0;
}

extern "C" int chefsi_subspace_reserve(chefsi_ctx_t *ctx, int ncol) { return ctx ? subspace_reserve(ctx, ncol, false) : 1; }
extern "C" int chefsi_subspace_reserve_kpt(chefsi_ctx_t *ctx, int ncol) { return ctx ? subspace_reserve(ctx, ncol, true) : 1; }

/* Hp = Y^H (H Y), Mp = Y^H Y.  Complex columns are handled as real columns of twice the length (subspace.cu). */
static int subspace_project(chefsi_ctx *ctx, const void *Y, size_t ldy, int ncol, void *Hp, void *Mp, size_t ldp, bool is_complex)
{
    if (ctx->multi) return multi_subspace_project(ctx, Y, ldy, ncol, Hp, Mp, ldp, is_complex);
    if (ncol <= 0 || ldp < (size_t)ncol || ldy < ctx->Nd) return chefsi_fail(ctx, "subspace_project: bad dimensions");
    if (subspace_reserve(ctx, ncol, is_complex)) return 1;
    ctx->small_ncol = ctx->q_ncol = 0;
    const int words = is_complex ? 2 : 1;
    const size_t esz = sizeof(double) * words, row = ctx->Nd * esz, pitch = ctx->ld * esz;
    if (!(ctx->res_ncol == ncol && ctx->res_host == Y && ctx->res_complex == (int)is_complex)) {
        if (ctx->res_unwritten_host == Y)
            return chefsi_fail(ctx, "subspace_project: this Y was filtered with NO_Y_COPYBACK (its host copy was never written) and the "
                                    "device copy does not match the call (%d columns, %s)", ncol, is_complex ? "complex" : "real");
        /* Y is not resident: upload it (pinned or pageable, the driver stages the latter) */
        CHEFSI_CUDA_DRAIN(ctx, cudaMemcpy2DAsync(ctx->d_res_Y, pitch, Y, ldy * esz, row, ncol, cudaMemcpyHostToDevice, ctx->stream));
        ctx->res_ncol = ncol;
        ctx->res_host = Y;
        ctx->res_complex = is_complex;
    }
    Profiler prof(ctx);
    /* H Y with c = 0 (eigenSolver.c:960-967, eigenSolverKpt.c:699-704), in groups of columns */
    const int grp = 256;
    for (int c0 = 0; c0 < ncol; c0 += grp) {
        const int nc = ncol - c0 < grp ? ncol - c0 : grp;
        if (apply_step(ctx, prof, (const char *)ctx->d_res_Y + (size_t)c0 * pitch, nullptr, (char *)ctx->d_res_W + (size_t)c0 * pitch, nc, 0.0,
                       1.0, 0.0, is_complex, false, false))
            return drain_streams(ctx, 1);
    }
    double *dHp = (double *)ctx->d_small[0], *dMp = (double *)ctx->d_small[1];
    const double *Yv = (const double *)ctx->d_res_Y, *Wv = (const double *)ctx->d_res_W;
    const size_t K = ctx->Nd * words, ldv = ctx->ld * words; /* real view of the columns */
    int n = launch_gemm_tn(ctx, Yv, ldv, Yv, ldv, ncol, ncol, K, 1.0, dMp, ncol, words, +1);
    if (n >= 0) { ctx->stats.kernel_launches += n; n = launch_gemm_tn(ctx, Yv, ldv, Wv, ldv, ncol, ncol, K, 1.0, dHp, ncol, words, +1); }
    if (n >= 0 && is_complex) {
        /* imaginary parts: Im(A^H B) = A_view^T (-i B)_view */
        ctx->stats.kernel_launches += n;
        const double *Tv = (const double *)ctx->d_res_T;
        n = launch_rot90(ctx, ctx->d_res_Y, ctx->d_res_T, ctx->Nd, ctx->ld, ncol, -1.0);
        if (n >= 0) n = launch_gemm_tn(ctx, Yv, ldv, Tv, ldv, ncol, ncol, K, 1.0, dMp + 1, ncol, 2, -1);
        if (n >= 0) n = launch_rot90(ctx, ctx->d_res_W, ctx->d_res_T, ctx->Nd, ctx->ld, ncol, -1.0);
        if (n >= 0) n = launch_gemm_tn(ctx, Yv, ldv, Tv, ldv, ncol, ncol, K, 1.0, dHp + 1, ncol, 2, -1);
        if (n >= 0) ctx->stats.kernel_launches += 6;
    }
    if (n < 0) return drain_streams(ctx, 1);
    const size_t w = (size_t)ncol * esz;
    CHEFSI_CUDA_DRAIN(ctx, cudaMemcpy2DAsync(Mp, ldp * esz, dMp, w, w, ncol, cudaMemcpyDeviceToHost, ctx->stream));
    CHEFSI_CUDA_DRAIN(ctx, cudaMemcpy2DAsync(Hp, ldp * esz, dHp, w, w, ncol, cudaMemcpyDeviceToHost, ctx->stream));
    CHEFSI_CUDA_DRAIN(ctx, cudaStreamSynchronize(ctx->stream));
    prof.finish();
    ctx->small_ncol = ncol; /* Hp, Mp stay on the device for chefsi_subspace_eig (rayleigh_ritz.cu) */
    ctx->small_complex = is_complex;
    return 0;
}

extern "C" int chefsi_subspace_project(chefsi_ctx_t *ctx, const double *Y, size_t ldy, int ncol, double *Hp, double *Mp, size_t ldp)
{
    if (!ctx || !Y || !Hp || !Mp) return 1;
    return subspace_project(ctx, Y, ldy, ncol, Hp, Mp, ldp, false);
}
extern "C" int chefsi_subspace_project_kpt(chefsi_ctx_t *ctx, const void *Y, size_t ldy, int ncol, void *Hp, void *Mp, size_t ldp)
{
    if (!ctx || !Y || !Hp || !Mp) return 1;
    return subspace_project(ctx, Y, ldy, ncol, Hp, Mp, ldp, true);
}

static int subspace_rotate(chefsi_ctx *ctx, const void *Q, size_t ldq, int ncol, void *X, size_t ldx, bool is_complex)
{
    if (ctx->multi) return multi_subspace_rotate(ctx, Q, ldq, ncol, X, ldx, is_complex);
    if (ctx->res_ncol != ncol || !ctx->d_res_Y || ctx->res_complex != (int)is_complex)
        return chefsi_fail(ctx, "subspace_rotate: no resident block of %d columns (call chefsi_subspace_project first)", ncol);
    if ((Q && ldq < (size_t)ncol) || ldx < ctx->Nd) return chefsi_fail(ctx, "subspace_rotate: bad dimensions");
    if (!Q && (ctx->q_ncol != ncol || ctx->q_complex != (int)is_complex))
        return chefsi_fail(ctx, "subspace_rotate: Q == NULL but no eigenvectors of %d columns are on the device (call chefsi_subspace_eig first)", ncol);
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    const int words = is_complex ? 2 : 1;
    const size_t esz = sizeof(double) * words, w = (size_t)ncol * esz;
    double *dQ = (double *)ctx->d_small[2];
    if (Q) CHEFSI_CUDA_DRAIN(ctx, cudaMemcpy2DAsync(dQ, w, Q, ldq * esz, w, ncol, cudaMemcpyHostToDevice, ctx->stream));
    ctx->small_ncol = 0; /* the complex branch splits Q into the Hp / Mp slots */
    const size_t K = ctx->Nd * words, ldv = ctx->ld * words;
    int n;
    if (!is_complex) {
        n = launch_gemm_nn(ctx, (const double *)ctx->d_res_Y, ldv, dQ, ncol, K, ncol, ncol, (double *)ctx->d_res_W, ldv, 0);
    } else {
        /* (Y Q)_view = Y_view Q_r + (i Y)_view Q_i */
        double *Qr = (double *)ctx->d_small[0], *Qi = (double *)ctx->d_small[1];
        n = launch_split_complex(ctx, dQ, ncol, ncol, ncol, Qr, Qi);
        if (n >= 0) n = launch_rot90(ctx, ctx->d_res_Y, ctx->d_res_T, ctx->Nd, ctx->ld, ncol, 1.0);
        if (n >= 0) n = launch_gemm_nn(ctx, (const double *)ctx->d_res_Y, ldv, Qr, ncol, K, ncol, ncol, (double *)ctx->d_res_W, ldv, 0);
        if (n >= 0) n = launch_gemm_nn(ctx, (const double *)ctx->d_res_T, ldv, Qi, ncol, K, ncol, ncol, (double *)ctx->d_res_W, ldv, 1);
        if (n >= 0) n = 4;
    }
    if (n < 0) return drain_streams(ctx, 1);
    ctx->stats.kernel_launches += n;
    if (band_store_put(ctx, X, ncol, is_complex, ctx->d_res_W)) return drain_streams(ctx, 1); /* device copy for the density (rayleigh_ritz.cu) */
    CHEFSI_CUDA_DRAIN(ctx, cudaMemcpy2DAsync(X, ldx * esz, ctx->d_res_W, ctx->ld * esz, ctx->Nd * esz, ncol, cudaMemcpyDeviceToHost, ctx->stream));
    CHEFSI_CUDA_DRAIN(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->res_ncol = 0; /* the block has been consumed */
    ctx->q_ncol = 0;
    return 0;
}

extern "C" int chefsi_subspace_rotate(chefsi_ctx_t *ctx, const double *Q, size_t ldq, int ncol, double *X, size_t ldx)
{
    if (!ctx || !X) return 1;
    return subspace_rotate(ctx, Q, ldq, ncol, X, ldx, false);
}
extern "C" int chefsi_subspace_rotate_kpt(chefsi_ctx_t *ctx, const void *Q, size_t ldq, int ncol, void *X, size_t ldx)
{
    if (!ctx || !X) return 1;
    return subspace_rotate(ctx, Q, ldq, ncol, X, ldx, true);
}

/* ---- (a Lap + c) x: the operator of the Poisson residual ------------------------------------------------------
 * Lap_vec_mult (src/lapVecRoutines.c:37-58) = Lap_plus_diag_vec_mult_{orth,nonorth} with b = 0, v = NULL
 * (:321-322): no potential, no projectors.  The kernels evaluate s1 ((-1/2 Lap + c') x), so s1 = -2 a, c' = c / s1. */
static int lapmult_host(chefsi_ctx *ctx, int ncol, double a, double c, const void *x, size_t ldi, void *y, size_t ldo,
                        bool is_complex)
{
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (ncol <= 0) return 0;
    if (a == 0.0) return chefsi_fail(ctx, "laplacian_mult: a must be non-zero");
    if (ldi < ctx->Nd || ldo < ctx->Nd) return chefsi_fail(ctx, "leading dimension smaller than the grid");
    if (ctx->multi) return multi_lapmult_host(ctx, ncol, a, c, x, ldi, y, ldo, is_complex);
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t esz = is_complex ? 2 * sizeof(double) : sizeof(double);
    const int chunk = chunk_columns(ctx, ncol, esz);
    if (ensure_bufs(ctx, (size_t)chunk * ctx->ld * esz)) return 1;
    const size_t row = ctx->Nd * esz, pitch = ctx->ld * esz;
    const double s1 = -2.0 * a;
    Profiler prof(ctx);
    if (chunk == ncol && want_staging(ctx, x, y, (size_t)ncol * row)) {
        if (staged_h2d(ctx, 0, ctx->d_buf[0], pitch, x, ldi * esz, row, ncol)) return drain_streams(ctx, 1);
        if (apply_step(ctx, prof, ctx->d_buf[0], nullptr, ctx->d_buf[1], ncol, c / s1, s1, 0.0, is_complex, false, false,
                       /*with_nl=*/false, /*with_veff=*/false))
            return drain_streams(ctx, 1);
        if (staged_d2h_begin(ctx, 1, ctx->d_buf[1], pitch, row, ncol)) return drain_streams(ctx, 1);
        CHEFSI_CUDA_DRAIN(ctx, cudaStreamSynchronize(ctx->stream));
        staged_d2h_finish(ctx, 1, y, ldo * esz, row, ncol);
        prof.finish();
        return 0;
    }
    for (int c0 = 0; c0 < ncol; c0 += chunk) {
        const int nc = (ncol - c0 < chunk) ? ncol - c0 : chunk;
        CHEFSI_CUDA_DRAIN(ctx, cudaMemcpy2DAsync(ctx->d_buf[0], pitch, (const char *)x + (size_t)c0 * ldi * esz, ldi * esz, row, nc,
                                                 cudaMemcpyHostToDevice, ctx->stream));
        if (apply_step(ctx, prof, ctx->d_buf[0], nullptr, ctx->d_buf[1], nc, c / s1, s1, 0.0, is_complex, false, false,
                       /*with_nl=*/false, /*with_veff=*/false))
            return drain_streams(ctx, 1);
        CHEFSI_CUDA_DRAIN(ctx, cudaMemcpy2DAsync((char *)y + (size_t)c0 * ldo * esz, ldo * esz, ctx->d_buf[1], pitch, row, nc,
                                                 cudaMemcpyDeviceToHost, ctx->stream));
    }
    CHEFSI_CUDA_DRAIN(ctx, cudaStreamSynchronize(ctx->stream));
    prof.finish();
    return 0;
}
extern "C" int chefsi_laplacian_mult(chefsi_ctx_t *ctx, int ncol, double a, double c, const double *x, size_t ldi, double *y,
                                     size_t ldo)
{
    return ctx ? lapmult_host(ctx, ncol, a, c, x, ldi, y, ldo, false) : 1;
}
extern "C" int chefsi_laplacian_mult_kpt(chefsi_ctx_t *ctx, int ncol, double a, double c, const void *x, size_t ldi, void *y,
                                         size_t ldo)
{
    return ctx ? lapmult_host(ctx, ncol, a, c, x, ldi, y, ldo, true) : 1;
}

/* ---- (D_dir + c) x: the first-derivative stencil along one lattice direction (kernels: gradient.cu) -------------
 * Gradient_vectors_dir[_kpt] (src/gradVecRoutines.c:32-51, src/gradVecRoutinesKpt.c:35-55): the operator of the GGA
 * density gradient, of the local / nonlocal force and stress terms; SURVEY.md 8f-4 "gradient ops sharing the stencil". */
static int gradmult_host(chefsi_ctx *ctx, int ncol, double c, const void *x, size_t ldi, void *Dx, size_t ldo, int dir, double kdir,
                         bool is_complex)
{
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (dir < 0 || dir > 2) return chefsi_fail(ctx, "gradient_mult: dir must be 0, 1 or 2");
    if (ncol <= 0) return 0;
    if (ldi < ctx->Nd || ldo < ctx->Nd) return chefsi_fail(ctx, "leading dimension smaller than the grid");
    if (ctx->multi) return multi_gradmult_host(ctx, ncol, c, x, ldi, Dx, ldo, dir, kdir, is_complex);
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t esz = is_complex ? 2 * sizeof(double) : sizeof(double);
    const int chunk = chunk_columns(ctx, ncol, esz);
    if (ensure_bufs(ctx, (size_t)chunk * ctx->ld * esz)) return 1;
    const size_t row = ctx->Nd * esz, pitch = ctx->ld * esz;
    if (chunk == ncol && want_staging(ctx, x, Dx, (size_t)ncol * row)) {
        if (staged_h2d(ctx, 0, ctx->d_buf[0], pitch, x, ldi * esz, row, ncol)) return drain_streams(ctx, 1);
        const int n = launch_gradient(ctx, ctx->d_buf[0], ctx->d_buf[1], ncol, dir, c, kdir, is_complex);
        if (n < 0) return drain_streams(ctx, 1);
        ctx->stats.kernel_launches += n;
        if (staged_d2h_begin(ctx, 1, ctx->d_buf[1], pitch, row, ncol)) return drain_streams(ctx, 1);
        CHEFSI_CUDA_DRAIN(ctx, cudaStreamSynchronize(ctx->stream));
        staged_d2h_finish(ctx, 1, Dx, ldo * esz, row, ncol);
        return 0;
    }
    /* large blocks: H2D of chunk k+1 and D2H of chunk k-1 overlap the kernel of chunk k through two buffer pairs
       (the kernel is one pass over the chunk, so the call is bound by the host link: both directions busy) */
    if (ensure_bufs2(ctx, (size_t)chunk * ctx->ld * esz) || ensure_pipe_events(ctx)) return 1;
    void *in[2] = {ctx->d_buf[0], ctx->d_buf2[0]}, *out[2] = {ctx->d_buf[1], ctx->d_buf2[1]};
    cudaEvent_t *ev_h2d = ctx->pipe_ev, *ev_out = ctx->pipe_ev + 6, *ev_d2h = ctx->pipe_ev + 9;
    int k = 0;
    for (int c0 = 0; c0 < ncol; c0 += chunk, k++) {
        const int nc = (ncol - c0 < chunk) ? ncol - c0 : chunk, s = k & 1;
        /* pair s is free again when the kernel of chunk k-2 has read in[s] and its result has left out[s] */
        if (k >= 2) CHEFSI_CUDA_DRAIN(ctx, cudaStreamWaitEvent(ctx->h2d_stream, ev_out[s], 0));
        CHEFSI_CUDA_DRAIN(ctx, cudaMemcpy2DAsync(in[s], pitch, (const char *)x + (size_t)c0 * ldi * esz, ldi * esz, row, nc,
                                                 cudaMemcpyHostToDevice, ctx->h2d_stream));
        CHEFSI_CUDA_DRAIN(ctx, cudaEventRecord(ev_h2d[s], ctx->h2d_stream));
        CHEFSI_CUDA_DRAIN(ctx, cudaStreamWaitEvent(ctx->stream, ev_h2d[s], 0));
        if (k >= 2) CHEFSI_CUDA_DRAIN(ctx, cudaStreamWaitEvent(ctx->stream, ev_d2h[s], 0));
        const int n = launch_gradient(ctx, in[s], out[s], nc, dir, c, kdir, is_complex);
        if (n < 0) return drain_streams(ctx, 1);
        ctx->stats.kernel_launches += n;
        CHEFSI_CUDA_DRAIN(ctx, cudaEventRecord(ev_out[s], ctx->stream));
        CHEFSI_CUDA_DRAIN(ctx, cudaStreamWaitEvent(ctx->d2h_stream, ev_out[s], 0));
        CHEFSI_CUDA_DRAIN(ctx, cudaMemcpy2DAsync((char *)Dx + (size_t)c0 * ldo * esz, ldo * esz, out[s], pitch, row, nc,
                                                 cudaMemcpyDeviceToHost, ctx->d2h_stream));
        CHEFSI_CUDA_DRAIN(ctx, cudaEventRecord(ev_d2h[s], ctx->d2h_stream));
    }
    for (int s = 0; s < 2 && s < k; s++) CHEFSI_CUDA_DRAIN(ctx, cudaStreamWaitEvent(ctx->stream, ev_d2h[s], 0));
    CHEFSI_CUDA_DRAIN(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int chefsi_gradient_mult_device(chefsi_ctx_t *ctx, int ncol, double c, const void *x, void *Dx, int dir, double kdir,
                                           int is_complex)
{
    if (!ctx) return 1;
    if (ctx->multi) return chefsi_fail(ctx, kMultiDeviceApi);
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (dir < 0 || dir > 2) return chefsi_fail(ctx, "gradient_mult: dir must be 0, 1 or 2");
    if (ncol <= 0) return 0;
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    const int n = launch_gradient(ctx, x, Dx, ncol, dir, c, kdir, is_complex != 0);
    if (n < 0) return 1;
    ctx->stats.kernel_launches += n;
    return 0;
}
extern "C" int chefsi_gradient_mult(chefsi_ctx_t *ctx, int ncol, double c, const double *x, size_t ldi, double *Dx, size_t ldo,
                                    int dir)
{
    return ctx ? gradmult_host(ctx, ncol, c, x, ldi, Dx, ldo, dir, 0.0, false) : 1;
}
extern "C" int chefsi_gradient_mult_kpt(chefsi_ctx_t *ctx, int ncol, double c, const void *x, size_t ldi, void *Dx, size_t ldo,
                                        int dir, double kdir)
{
    return ctx ? gradmult_host(ctx, ncol, c, x, ldi, Dx, ldo, dir, kdir, true) : 1;
}

/* ---- misc ---------------------------------------------------------------------------------------- */
extern "C" int chefsi_fill_random_device(chefsi_ctx_t *ctx, void *buf, int ncol, long long first_col,
                                         unsigned long long seed, int is_complex)
{
    if (!ctx) return 1;
    if (ctx->multi) return chefsi_fail(ctx, kMultiDeviceApi);
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    const int n = launch_fill_random(ctx, buf, ncol, first_col, seed, is_complex != 0);
    if (n < 0) return 1;
    ctx->stats.kernel_launches += n;
    return 0;
}

/* dense block with an arbitrary leading dimension <-> internal layout (which differs only in ld): strided D2D copies */
extern "C" int chefsi_pack_device(chefsi_ctx_t *ctx, const void *dense, size_t ld_dense, void *packed, int ncol,
                                  int is_complex)
{
    if (!ctx) return 1;
    if (ctx->multi) return chefsi_fail(ctx, kMultiDeviceApi);
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (ld_dense < ctx->Nd) return chefsi_fail(ctx, "leading dimension smaller than the grid");
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t esz = is_complex ? 16 : 8;
    if (ncol > 0)
        CHEFSI_CUDA(ctx, cudaMemcpy2DAsync(packed, ctx->ld * esz, dense, ld_dense * esz, ctx->Nd * esz, ncol,
                                           cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}
extern "C" int chefsi_unpack_device(chefsi_ctx_t *ctx, const void *packed, void *dense, size_t ld_dense, int ncol,
                                    int is_complex)
{
    if (!ctx) return 1;
    if (ctx->multi) return chefsi_fail(ctx, kMultiDeviceApi);
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (ld_dense < ctx->Nd) return chefsi_fail(ctx, "leading dimension smaller than the grid");
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t esz = is_complex ? 16 : 8;
    if (ncol > 0)
        CHEFSI_CUDA(ctx, cudaMemcpy2DAsync(dense, ld_dense * esz, packed, ctx->ld * esz, ctx->Nd * esz, ncol,
                                           cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

extern "C" int chefsi_host_register(chefsi_ctx_t *ctx, void *ptr, size_t bytes)
{
    if (!ctx || !ptr || !bytes) return 1;
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable); /* portable: every device of a multi-device context copies from it */
    if (e != cudaSuccess) { cudaGetLastError(); return chefsi_fail(ctx, "cudaHostRegister(%zu bytes): %s", bytes, cudaGetErrorString(e)); }
    return 0;
}
extern "C" int chefsi_host_unregister(chefsi_ctx_t *ctx, void *ptr)
{
    if (!ctx || !ptr) return 1;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return chefsi_fail(ctx, "cudaHostUnregister: %s", cudaGetErrorString(e)); }
    return 0;
}

extern "C" int chefsi_get_stats(const chefsi_ctx_t *ctx, chefsi_stats_t *out)
{
    if (!ctx || !out) return 1;
    *out = ctx->stats;
    if (ctx->multi) return 0; /* the leader's copy is refreshed by every split call (sum of launches, max of times) */
    if (ctx->d_sync) { /* word 1 of the round-barrier block counts producers that gave up waiting */
        unsigned int t = 0;
        if (cudaSetDevice(ctx->device) == cudaSuccess &&
            cudaMemcpyAsync(&t, ctx->d_sync + 1, sizeof(t), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
            cudaStreamSynchronize(ctx->stream) == cudaSuccess)
            out->round_barrier_timeouts = t;
        else cudaGetLastError();
    }
    return 0;
}
/* ---- domain-split building blocks (see include/chefsi_b200.h) ---------------------------------------- */
extern "C" int chefsi_stencil_step_device(chefsi_ctx_t *ctx, const double *x, const double *xprev, double *out, int ncol,
                                          double c, double s1, double s2)
{
    if (!ctx) return 1;
    if (ctx->multi) return chefsi_fail(ctx, kMultiDeviceApi);
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    Profiler prof(ctx);
    const int rc = apply_step(ctx, prof, x, xprev, out, ncol, c, s1, s2, false, false, false, /*with_nl=*/false);
    prof.finish();
    return rc;
}

extern "C" int chefsi_nloc_project_device(chefsi_ctx_t *ctx, const double *x, int ncol, double *alpha_out)
{
    if (!ctx || !alpha_out) return 1;
    if (ctx->multi) return chefsi_fail(ctx, kMultiDeviceApi);
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (ctx->nl.n_img == 0 || ctx->nl.ntot == 0) return chefsi_fail(ctx, "nloc_project: no projectors on this context");
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    const int saved_min = ctx->alpha_reduce_min;
    ctx->alpha_reduce_min = -1; /* per-atom sums wanted for THIS call only: later filter calls keep their own threshold */
    int n = launch_nloc(ctx, NLOC_PROJECT, const_cast<double *>(x), ctx->ld, ncol, 0.0, false);
    if (n >= 0) { ctx->stats.kernel_launches += n; n = launch_alpha_reduce(ctx, ncol, false); }
    ctx->alpha_reduce_min = saved_min;
    if (n < 0) return 1;
    ctx->stats.kernel_launches += n;
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(alpha_out, ctx->d_alpha_sum, sizeof(double) * (size_t)ctx->nl.ntot * ncol,
                                     cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

extern "C" int chefsi_nloc_expand_device(chefsi_ctx_t *ctx, double *out, int ncol, double scale, const double *alpha_in)
{
    if (!ctx || !alpha_in) return 1;
    if (ctx->multi) return chefsi_fail(ctx, kMultiDeviceApi);
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (ctx->nl.n_img == 0 || ctx->nl.ntot == 0) return chefsi_fail(ctx, "nloc_expand: no projectors on this context");
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    const int saved_min = ctx->alpha_reduce_min;
    ctx->alpha_reduce_min = -1;
    if (nloc_ensure_alpha(ctx, ncol, false)) { ctx->alpha_reduce_min = saved_min; return 1; }
    cudaError_t ce = cudaMemcpyAsync(ctx->d_alpha_sum, alpha_in, sizeof(double) * (size_t)ctx->nl.ntot * ncol,
                                     cudaMemcpyDeviceToDevice, ctx->stream);
    if (ce != cudaSuccess) { ctx->alpha_reduce_min = saved_min; return chefsi_fail(ctx, "nloc_expand: %s", cudaGetErrorString(ce)); }
    ctx->alpha_sum_external = 1; /* expand reads the buffer as it is */
    const int n = launch_nloc(ctx, NLOC_EXPAND, out, ctx->ld, ncol, scale, false);
    ctx->alpha_sum_external = 0;
    ctx->alpha_reduce_min = saved_min;
    if (n < 0) return 1;
    ctx->stats.kernel_launches += n;
    return 0;
}

extern "C" int chefsi_set_profiling(chefsi_ctx_t *ctx, int on)
{
    if (!ctx) return 1;
    ctx->profiling = on;
    if (ctx->multi) multi_set_profiling(ctx, on);
    return 0;
}
extern "C" void *chefsi_stream(chefsi_ctx_t *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int chefsi_multi_info(const chefsi_ctx_t *ctx, int *ndev, int *uses_nccl, unsigned long long *bcast_calls,
                                 unsigned long long *bcast_bytes)
{
    if (!ctx) return 1;
    if (ndev) *ndev = multi_size(ctx);
    if (uses_nccl) *uses_nccl = multi_uses_nccl(ctx);
    unsigned long long c = 0, b = 0;
    multi_bcast_stats(ctx, &c, &b);
    if (bcast_calls) *bcast_calls = c;
    if (bcast_bytes) *bcast_bytes = b;
    return 0;
}
