/*
 * rayleigh_ritz.cu -- the two steps that close the per-SCF device-resident loop:
 *
 *   chefsi_subspace_eig[_kpt]        Hp q = lambda Mp q for the Ns x Ns matrices the projection left on the device.
 *                                    Replaces DP_Solve_Generalized_EigenProblem (src/eigenSolver.c:1262-1375:
 *                                    LAPACKE_dsygvd(itype 1, 'V', 'U')) and its k-point twin (src/eigenSolverKpt.c:836-:
 *                                    LAPACKE_zhegvd).  The reference calls a LAPACK library here (and hooks an accelerator
 *                                    DSYGV / ZHEGV in its SPARCX_ACCEL build, :1267-1298); this is the same library call
 *                                    on the device, cusolverDnDsygvd / cusolverDnZhegvd, loaded with dlopen so that the
 *                                    filter path has no cuSOLVER dependency.  Hp, Mp never leave the device; the
 *                                    eigenvectors stay there for chefsi_subspace_rotate (Q == NULL).
 *
 *   chefsi_density_accumulate[_kpt]  rho[i] += sum_n g[n] |X[i + n ld]|^2 for one k-point / spin block: the loop body of
 *                                    CalculateDensity_psi (src/electronDensity.c:104-200).  When the band store is on
 *                                    (chefsi_band_store), chefsi_subspace_rotate keeps a device copy of every rotated
 *                                    block, keyed by its host address, and the density reads that copy: per SCF
 *                                    iteration only Nd doubles per block cross PCIe for the density.
 */
#include <cusolverDn.h>
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "chefsi_internal.h"

/* ---- cuSOLVER, loaded at run time ------------------------------------------------------------------------------ */
struct EigState {
    void *lib = nullptr;
    cusolverDnHandle_t handle = nullptr;
    decltype(&cusolverDnCreate) create = nullptr;
    decltype(&cusolverDnDestroy) destroy = nullptr;
    decltype(&cusolverDnSetStream) setStream = nullptr;
    decltype(&cusolverDnDsygvd_bufferSize) dsygvd_bufferSize = nullptr;
    decltype(&cusolverDnDsygvd) dsygvd = nullptr;
    decltype(&cusolverDnZhegvd_bufferSize) zhegvd_bufferSize = nullptr;
    decltype(&cusolverDnZhegvd) zhegvd = nullptr;
    void *d_work = nullptr;
    size_t work_bytes = 0;
    double *d_lambda = nullptr; /* eigenvalues, then one int of devInfo */
    size_t lambda_n = 0;
};

struct BandBlock {
    const void *host = nullptr;
    int ncol = 0, is_complex = 0;
    bool valid = false;
    void *dev = nullptr;
    size_t bytes = 0;
};
struct BandStore {
    std::vector<BandBlock> blocks;
    double *d_rho = nullptr, *d_g = nullptr, *d_part = nullptr;
    size_t rho_n = 0, g_n = 0, part_n = 0;
    double *h_rho = nullptr; /* pinned */
    double *h_g = nullptr;   /* pinned */
    size_t h_g_n = 0;
};

namespace {

int eig_load(chefsi_ctx *ctx)
{
    if (ctx->eig && ctx->eig->handle) return 0;
    if (!ctx->eig) ctx->eig = new EigState();
    EigState *E = ctx->eig;
    if (!E->lib) {
        const char *names[] = {getenv("CHEFSI_B200_CUSOLVER_LIB"), "libcusolver.so.11", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11",
                               "/usr/local/cuda/lib64/libcusolver.so"};
        for (const char *nm : names) {
            if (!nm) continue;
            E->lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (E->lib) break;
        }
        if (!E->lib) return chefsi_fail(ctx, "subspace_eig: cuSOLVER could not be loaded (%s); set CHEFSI_B200_CUSOLVER_LIB", dlerror());
#define CHEFSI_SYM(field, name)                                                       \
    E->field = (decltype(E->field))dlsym(E->lib, name);                               \
    if (!E->field) return chefsi_fail(ctx, "subspace_eig: cuSOLVER lacks %s", name)
        CHEFSI_SYM(create, "cusolverDnCreate");
        CHEFSI_SYM(destroy, "cusolverDnDestroy");
        CHEFSI_SYM(setStream, "cusolverDnSetStream");
        CHEFSI_SYM(dsygvd_bufferSize, "cusolverDnDsygvd_bufferSize");
        CHEFSI_SYM(dsygvd, "cusolverDnDsygvd");
        CHEFSI_SYM(zhegvd_bufferSize, "cusolverDnZhegvd_bufferSize");
        CHEFSI_SYM(zhegvd, "cusolverDnZhegvd");
#undef CHEFSI_SYM
    }
    cusolverStatus_t s = E->create(&E->handle);
    if (s != CUSOLVER_STATUS_SUCCESS) { E->handle = nullptr; return chefsi_fail(ctx, "cusolverDnCreate failed (%d)", (int)s); }
    s = E->setStream(E->handle, ctx->stream);
    if (s != CUSOLVER_STATUS_SUCCESS) return chefsi_fail(ctx, "cusolverDnSetStream failed (%d)", (int)s);
    return 0;
}

int ensure_small(chefsi_ctx *ctx, int ncol, bool is_complex)
{
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

int subspace_eig(chefsi_ctx *ctx, int ncol, const void *Hp, const void *Mp, size_t ldp, double *lambda, void *Q, size_t ldq, bool is_complex)
{
    if (!ctx || !lambda || ncol <= 0) return 1;
    if (ctx->multi) { /* Ns x Ns matrices do not split: the first device solves, from the host copies */
        if (!Hp || !Mp) return chefsi_fail(ctx, "subspace_eig: a multi-device context needs the host copies of Hp and Mp");
        chefsi_ctx *k = multi_first(ctx);
        const int rc = subspace_eig(k, ncol, Hp, Mp, ldp, lambda, Q, ldq, is_complex);
        if (rc) { strncpy(ctx->err, k->err, sizeof(ctx->err) - 1); ctx->err[sizeof(ctx->err) - 1] = 0; }
        return rc;
    }
    if ((Hp == nullptr) != (Mp == nullptr)) return chefsi_fail(ctx, "subspace_eig: pass both Hp and Mp, or neither");
    if (Hp && ldp < (size_t)ncol) return chefsi_fail(ctx, "subspace_eig: bad ldp");
    if (Q && ldq < (size_t)ncol) return chefsi_fail(ctx, "subspace_eig: bad ldq");
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    if (eig_load(ctx)) return 1;
    EigState *E = ctx->eig;
    const size_t esz = sizeof(double) * (is_complex ? 2 : 1), w = (size_t)ncol * esz;
    cudaStream_t st = ctx->stream;
    if (Hp) {
        if (ensure_small(ctx, ncol, is_complex)) return 1;
        CHEFSI_CUDA(ctx, cudaMemcpy2DAsync(ctx->d_small[0], w, Hp, ldp * esz, w, ncol, cudaMemcpyHostToDevice, st));
        CHEFSI_CUDA(ctx, cudaMemcpy2DAsync(ctx->d_small[1], w, Mp, ldp * esz, w, ncol, cudaMemcpyHostToDevice, st));
    } else if (ctx->small_ncol != ncol || ctx->small_complex != (int)is_complex) {
        return chefsi_fail(ctx, "subspace_eig: no projected matrices of %d %s columns on the device (call chefsi_subspace_project first, or pass Hp and Mp)",
                           ncol, is_complex ? "complex" : "real");
    }
    ctx->small_ncol = 0; /* the solver overwrites Hp (eigenvectors) and Mp (Cholesky factor) */
    ctx->q_ncol = 0;
    if ((size_t)ncol + 2 > E->lambda_n) {
        cudaFree(E->d_lambda);
        E->d_lambda = nullptr;
        E->lambda_n = 0;
        CHEFSI_CUDA(ctx, cudaMalloc(&E->d_lambda, ((size_t)ncol + 2) * sizeof(double)));
        E->lambda_n = (size_t)ncol + 2;
    }
    int *d_info = (int *)(E->d_lambda + ncol);
    int lwork = 0;
    cusolverStatus_t s;
    if (!is_complex)
        s = E->dsygvd_bufferSize(E->handle, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, ncol, (double *)ctx->d_small[0], ncol,
                                 (double *)ctx->d_small[1], ncol, E->d_lambda, &lwork);
    else
        s = E->zhegvd_bufferSize(E->handle, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, ncol, (cuDoubleComplex *)ctx->d_small[0],
                                 ncol, (cuDoubleComplex *)ctx->d_small[1], ncol, E->d_lambda, &lwork);
    if (s != CUSOLVER_STATUS_SUCCESS) return chefsi_fail(ctx, "subspace_eig: workspace query failed (%d)", (int)s);
    const size_t need = (size_t)lwork * esz;
    if (need > E->work_bytes) {
        cudaFree(E->d_work);
        E->d_work = nullptr;
        E->work_bytes = 0;
        CHEFSI_CUDA(ctx, cudaMalloc(&E->d_work, need));
        E->work_bytes = need;
    }
    if (!is_complex)
        s = E->dsygvd(E->handle, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, ncol, (double *)ctx->d_small[0], ncol,
                      (double *)ctx->d_small[1], ncol, E->d_lambda, (double *)E->d_work, lwork, d_info);
    else
        s = E->zhegvd(E->handle, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, ncol, (cuDoubleComplex *)ctx->d_small[0], ncol,
                      (cuDoubleComplex *)ctx->d_small[1], ncol, E->d_lambda, (cuDoubleComplex *)E->d_work, lwork, d_info);
    if (s != CUSOLVER_STATUS_SUCCESS) return chefsi_fail(ctx, "subspace_eig: the solver failed to start (%d)", (int)s);
    /* the rotation's complex branch splits Q into d_small[0], d_small[1]: the eigenvectors move to the Q slot */
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(ctx->d_small[2], ctx->d_small[0], (size_t)ncol * w, cudaMemcpyDeviceToDevice, st));
    int info = -1;
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(lambda, E->d_lambda, (size_t)ncol * sizeof(double), cudaMemcpyDeviceToHost, st));
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (Q) CHEFSI_CUDA(ctx, cudaMemcpy2DAsync(Q, ldq * esz, ctx->d_small[2], w, w, ncol, cudaMemcpyDeviceToHost, st));
    CHEFSI_CUDA(ctx, cudaStreamSynchronize(st));
    if (info != 0)
        return chefsi_fail(ctx, info > ncol ? "subspace_eig: Mp is not positive definite (leading minor %d)" : "subspace_eig: the eigensolver did not converge (info %d)",
                           info > ncol ? info - ncol : info);
    ctx->q_ncol = ncol;
    ctx->q_complex = is_complex;
    return 0;
}

/* ---- density ------------------------------------------------------------------------------------------------- */
constexpr int kThreads = 256;

/* part[s][i] = sum over the columns n of slab s (in order) of g[n] |x_n(i)|^2 */
template <int WORDS>
__global__ void __launch_bounds__(kThreads) density_kernel(const double *__restrict__ X, size_t ld, size_t Nd, int ncol, int cols_per_slab,
                                                           const double *__restrict__ g, double *__restrict__ part)
{
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= Nd) return;
    const int n0 = blockIdx.y * cols_per_slab, n1 = min(ncol, n0 + cols_per_slab);
    double s = 0.0;
    for (int n = n0; n < n1; n++) {
        const double gn = g[n];
        if (WORDS == 1) {
            const double v = X[(size_t)n * ld + i];
            s = fma(gn * v, v, s);
        } else {
            const double2 v = reinterpret_cast<const double2 *>(X)[(size_t)n * ld + i];
            s = fma(gn, fma(v.x, v.x, v.y * v.y), s);
        }
    }
    part[(size_t)blockIdx.y * Nd + i] = s;
}

__global__ void __launch_bounds__(kThreads) density_reduce_kernel(const double *__restrict__ part, size_t Nd, int slabs, double *__restrict__ rho)
{
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= Nd) return;
    double s = 0.0;
    for (int k = 0; k < slabs; k++) s += part[(size_t)k * Nd + i];
    rho[i] = s;
}

BandBlock *band_find(BandStore *B, const void *host, int ncol, bool is_complex)
{
    if (!B) return nullptr;
    for (BandBlock &b : B->blocks)
        if (b.valid && b.host == host && b.ncol == ncol && b.is_complex == (int)is_complex) return &b;
    return nullptr;
}

int density_accumulate(chefsi_ctx *ctx, const void *X, size_t ldx, int ncol, const double *g, double *rho, bool is_complex)
{
    if (!ctx || !X || !g || !rho || ncol <= 0) return 1;
    if (ctx->multi) { /* a streaming reduction over one block: the first device does it */
        chefsi_ctx *k = multi_first(ctx);
        const int rc = density_accumulate(k, X, ldx, ncol, g, rho, is_complex);
        if (rc) { strncpy(ctx->err, k->err, sizeof(ctx->err) - 1); ctx->err[sizeof(ctx->err) - 1] = 0; }
        return rc;
    }
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (ldx < ctx->Nd) return chefsi_fail(ctx, "density_accumulate: bad ldx");
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->bands) ctx->bands = new BandStore();
    BandStore *B = ctx->bands;
    const int words = is_complex ? 2 : 1;
    const size_t Nd = ctx->Nd, esz = sizeof(double) * words;
    cudaStream_t st = ctx->stream;
    const void *dX = nullptr;
    BandBlock *blk = band_find(B, X, ncol, is_complex);
    if (blk) {
        dX = blk->dev;
        ctx->stats.density_resident_blocks++;
    } else {
        /* not resident: upload the block into the subspace work block */
        if (is_complex ? chefsi_subspace_reserve_kpt(ctx, ncol) : chefsi_subspace_reserve(ctx, ncol)) return 1;
        CHEFSI_CUDA(ctx, cudaMemcpy2DAsync(ctx->d_res_W, ctx->ld * esz, X, ldx * esz, Nd * esz, ncol, cudaMemcpyHostToDevice, st));
        dX = ctx->d_res_W;
        ctx->stats.density_uploaded_blocks++;
    }
    /* column slabs so that small grids still fill the device; the slab partials are added in slab order */
    const int gx = (int)((Nd + kThreads - 1) / kThreads);
    int slabs = (2 * ctx->num_sms * 4 + gx - 1) / gx;
    if (slabs > (ncol + 7) / 8) slabs = (ncol + 7) / 8;
    if (slabs < 1) slabs = 1;
    if (slabs > 64) slabs = 64;
    const int cps = (ncol + slabs - 1) / slabs;
    slabs = (ncol + cps - 1) / cps;
    if (Nd > B->rho_n) {
        cudaFree(B->d_rho);
        if (B->h_rho) cudaFreeHost(B->h_rho);
        B->d_rho = B->h_rho = nullptr;
        B->rho_n = 0;
        CHEFSI_CUDA(ctx, cudaMalloc(&B->d_rho, Nd * sizeof(double)));
        CHEFSI_CUDA(ctx, cudaMallocHost(&B->h_rho, Nd * sizeof(double)));
        B->rho_n = Nd;
    }
    if ((size_t)slabs * Nd > B->part_n) {
        cudaFree(B->d_part);
        B->d_part = nullptr;
        B->part_n = 0;
        CHEFSI_CUDA(ctx, cudaMalloc(&B->d_part, (size_t)slabs * Nd * sizeof(double)));
        B->part_n = (size_t)slabs * Nd;
    }
    if ((size_t)ncol > B->g_n) {
        cudaFree(B->d_g);
        if (B->h_g) cudaFreeHost(B->h_g);
        B->d_g = nullptr;
        B->h_g = nullptr;
        B->g_n = 0;
        CHEFSI_CUDA(ctx, cudaMalloc(&B->d_g, (size_t)ncol * sizeof(double)));
        CHEFSI_CUDA(ctx, cudaMallocHost(&B->h_g, (size_t)ncol * sizeof(double)));
        B->g_n = ncol;
    }
    memcpy(B->h_g, g, (size_t)ncol * sizeof(double));
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(B->d_g, B->h_g, (size_t)ncol * sizeof(double), cudaMemcpyHostToDevice, st));
    const dim3 grid(gx, slabs);
    if (is_complex) density_kernel<2><<<grid, kThreads, 0, st>>>((const double *)dX, ctx->ld, Nd, ncol, cps, B->d_g, B->d_part);
    else density_kernel<1><<<grid, kThreads, 0, st>>>((const double *)dX, ctx->ld, Nd, ncol, cps, B->d_g, B->d_part);
    density_reduce_kernel<<<gx, kThreads, 0, st>>>(B->d_part, Nd, slabs, B->d_rho);
    ctx->stats.kernel_launches += 2;
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(B->h_rho, B->d_rho, Nd * sizeof(double), cudaMemcpyDeviceToHost, st));
    CHEFSI_CUDA(ctx, cudaStreamSynchronize(st));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return chefsi_fail(ctx, "density kernels: %s", cudaGetErrorString(e));
    for (size_t i = 0; i < Nd; i++) rho[i] += B->h_rho[i];
    if (blk) blk->valid = false; /* consumed: the next SCF iteration's rotation refreshes it */
    return 0;
}

}  // namespace

/* chefsi_api.cu: the rotation hands its result over before the block is reused */
int band_store_put(chefsi_ctx *ctx, const void *host, int ncol, bool is_complex, const void *d_block)
{
    BandStore *B = ctx->bands;
    if (!B || B->blocks.empty()) return 0;
    const size_t bytes = (size_t)ncol * ctx->ld * sizeof(double) * (is_complex ? 2 : 1);
    BandBlock *slot = nullptr;
    for (BandBlock &b : B->blocks)
        if (b.host == host) { slot = &b; break; }
    if (!slot)
        for (BandBlock &b : B->blocks)
            if (!b.valid && !b.host) { slot = &b; break; }
    if (!slot)
        for (BandBlock &b : B->blocks)
            if (!b.valid) { slot = &b; break; }
    if (!slot) { ctx->stats.band_store_misses++; return 0; } /* store full: the density call will upload this block */
    slot->valid = false;
    if (bytes > slot->bytes) {
        cudaFree(slot->dev);
        slot->dev = nullptr;
        slot->bytes = 0;
        if (cudaMalloc(&slot->dev, bytes) != cudaSuccess) { /* does not fit: not an error, the block is simply not kept */
            cudaGetLastError();
            slot->host = nullptr;
            ctx->stats.band_store_misses++;
            return 0;
        }
        slot->bytes = bytes;
    }
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(slot->dev, d_block, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    slot->host = host;
    slot->ncol = ncol;
    slot->is_complex = is_complex;
    slot->valid = true;
    return 0;
}

/* a host block that is about to be filtered (its contents change): drop its device copy */
void band_store_invalidate(chefsi_ctx *ctx, const void *host)
{
    if (!ctx->bands) return;
    for (BandBlock &b : ctx->bands->blocks)
        if (b.host == host) b.valid = false;
}

/* a new grid: every kept block has the old shape */
void band_store_clear(chefsi_ctx *ctx)
{
    if (!ctx->bands) return;
    for (BandBlock &b : ctx->bands->blocks) { b.valid = false; b.host = nullptr; }
}

void rayleigh_ritz_destroy(chefsi_ctx *ctx)
{
    if (ctx->eig) {
        EigState *E = ctx->eig;
        if (E->handle && E->destroy) E->destroy(E->handle);
        cudaFree(E->d_work);
        cudaFree(E->d_lambda);
        /* the library stays loaded: unloading cuSOLVER at exit races with its own teardown */
        delete E;
        ctx->eig = nullptr;
    }
    if (ctx->bands) {
        BandStore *B = ctx->bands;
        for (BandBlock &b : B->blocks) cudaFree(b.dev);
        cudaFree(B->d_rho); cudaFree(B->d_g); cudaFree(B->d_part);
        if (B->h_rho) cudaFreeHost(B->h_rho);
        if (B->h_g) cudaFreeHost(B->h_g);
        delete B;
        ctx->bands = nullptr;
    }
}

extern "C" int chefsi_band_store(chefsi_ctx_t *ctx, int max_blocks)
{
    if (!ctx) return 1;
    if (ctx->multi) return max_blocks > 0 ? chefsi_fail(ctx, "band_store: single-device contexts only") : 0;
    if (max_blocks < 0 || max_blocks > 4096) return chefsi_fail(ctx, "band_store: max_blocks out of range");
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->bands) ctx->bands = new BandStore();
    BandStore *B = ctx->bands;
    CHEFSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t i = max_blocks; i < B->blocks.size(); i++) cudaFree(B->blocks[i].dev);
    B->blocks.resize(max_blocks);
    return 0;
}

extern "C" int chefsi_subspace_eig(chefsi_ctx_t *ctx, int ncol, const double *Hp, const double *Mp, size_t ldp, double *lambda, double *Q,
                                   size_t ldq)
{
    return subspace_eig(ctx, ncol, Hp, Mp, ldp, lambda, Q, ldq, false);
}
extern "C" int chefsi_subspace_eig_kpt(chefsi_ctx_t *ctx, int ncol, const void *Hp, const void *Mp, size_t ldp, double *lambda, void *Q,
                                       size_t ldq)
{
    return subspace_eig(ctx, ncol, Hp, Mp, ldp, lambda, Q, ldq, true);
}
extern "C" int chefsi_density_accumulate(chefsi_ctx_t *ctx, const double *X, size_t ldx, int ncol, const double *g, double *rho)
{
    return density_accumulate(ctx, X, ldx, ncol, g, rho, false);
}
extern "C" int chefsi_density_accumulate_kpt(chefsi_ctx_t *ctx, const void *X, size_t ldx, int ncol, const double *g, double *rho)
{
    return density_accumulate(ctx, X, ldx, ncol, g, rho, true);
}
