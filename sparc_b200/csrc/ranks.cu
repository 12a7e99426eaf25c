/*
 * ranks.cu -- the Rayleigh-Ritz products of a band-parallel run with ONE PROCESS PER GPU.
 *
 * SPARC splits the Ns orbitals over its band communicator (NB = ceil(Ns / P), src/parallelization.c:403-428); the filter
 * needs no communication along that axis, but the projection Mp = Y^T Y, Hp = Y^T H Y and the rotation X = Y Q need every
 * pair of column blocks.  The reference re-distributes the block first (BP2DP: MPI_Alltoallv, src/parallelization.c:2535,
 * eigenSolver.c:977-990; pdgemr2d + pdgemm in Project_Hamiltonian, :1504-1582).  Here nothing is re-distributed: rank I
 * computes the column block I of Hp / Mp / Y Q, and the A operand of its GEMM launches is the OTHER rank's resident
 * block Y_J, mapped into this process with CUDA IPC -- the gemm kernels' cp.async loads pull the tiles over NVLink
 * (or through the local HBM when two ranks share a device) while the DMMAs of the previous tiles run, i.e. the
 * all-gather is fused into the product.  multi.cu is the same scheme inside one process (plain peer pointers).
 *
 * The caller (sparc_b200/band_parallel.py: torch.distributed for the handle exchange, the barriers and the all-gather
 * of the small Hp / Mp column blocks) owns the ordering between ranks:
 *     every rank: Y_I resident (KEEP_Y filter or chefsi_rank_load)        -> barrier
 *     chefsi_rank_project                                                 -> all-gather of the Ns x nc_I blocks, eigensolve
 *     chefsi_rank_rotate_prepare (complex: T_I = i Y_I)                   -> barrier
 *     chefsi_rank_rotate                                                  -> barrier (before anybody overwrites Y)
 */
#include <cstring>

#include "chefsi_internal.h"

namespace {

struct RankBufs {
    void *d[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}; /* Hp block, Mp block, Q block, Re Q, Im Q */
    size_t bytes = 0;
};

RankBufs *rank_bufs(chefsi_ctx *ctx, size_t need)
{
    RankBufs *rb = (RankBufs *)ctx->rank_state;
    if (!rb) ctx->rank_state = rb = new RankBufs();
    if (need > rb->bytes) {
        for (void *&p : rb->d) { cudaFree(p); p = nullptr; }
        rb->bytes = 0;
        for (void *&p : rb->d)
            if (cudaMalloc(&p, need) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        rb->bytes = need;
    }
    return rb;
}

int check_ranks(chefsi_ctx *ctx, const char *what, int is_complex, int nranks, int rank, const int *ncols, int *ncol_total, int *c0)
{
    if (ctx->multi) return chefsi_fail(ctx, "%s: takes a single-device context (one process per GPU)", what);
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (nranks < 1 || rank < 0 || rank >= nranks || !ncols) return chefsi_fail(ctx, "%s: bad rank layout", what);
    int tot = 0;
    for (int r = 0; r < nranks; r++) {
        if (ncols[r] < 0) return chefsi_fail(ctx, "%s: negative column count", what);
        if (r == rank) *c0 = tot;
        tot += ncols[r];
    }
    if (ncols[rank] <= 0) return chefsi_fail(ctx, "%s: this rank owns no columns", what);
    if (ctx->res_ncol != ncols[rank] || ctx->res_complex != (is_complex != 0))
        return chefsi_fail(ctx, "%s: the rank's block of %d %s columns is not resident (filter with CHEFSI_FLAG_KEEP_Y or call chefsi_rank_load)",
                           what, ncols[rank], is_complex ? "complex" : "real");
    *ncol_total = tot;
    return 0;
}

}  // namespace

void rank_state_destroy(chefsi_ctx *ctx)
{
    RankBufs *rb = (RankBufs *)ctx->rank_state;
    if (!rb) return;
    for (void *p : rb->d) cudaFree(p);
    delete rb;
    ctx->rank_state = nullptr;
}

/* make the rank's block resident from host memory (what a KEEP_Y filter leaves behind) */
extern "C" int chefsi_rank_load(chefsi_ctx_t *ctx, const void *Y, size_t ldy, int ncol, int is_complex)
{
    if (!ctx || !Y) return 1;
    if (ctx->multi) return chefsi_fail(ctx, "rank_load: takes a single-device context");
    if (ncol <= 0 || ldy < ctx->Nd) return chefsi_fail(ctx, "rank_load: bad dimensions");
    if (is_complex ? chefsi_subspace_reserve_kpt(ctx, ncol) : chefsi_subspace_reserve(ctx, ncol)) return 1;
    const size_t esz = sizeof(double) * (is_complex ? 2 : 1);
    CHEFSI_CUDA(ctx, cudaMemcpy2DAsync(ctx->d_res_Y, ctx->ld * esz, Y, ldy * esz, ctx->Nd * esz, ncol, cudaMemcpyHostToDevice, ctx->stream));
    CHEFSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->res_ncol = ncol;
    ctx->res_host = Y;
    ctx->res_complex = is_complex != 0;
    ctx->res_unwritten_host = nullptr;
    return 0;
}

/* device address of the resident blocks (0: Y, 1: W = H Y / Y Q, 2: T), for ranks that share one process */
extern "C" void *chefsi_resident_ptr(chefsi_ctx_t *ctx, int which)
{
    if (!ctx || ctx->multi) return nullptr;
    return which == 0 ? ctx->d_res_Y : (which == 1 ? ctx->d_res_W : (which == 2 ? ctx->d_res_T : nullptr));
}

extern "C" int chefsi_ipc_export(chefsi_ctx_t *ctx, int which, void *handle64)
{
    if (!ctx || !handle64) return 1;
    static_assert(sizeof(cudaIpcMemHandle_t) == CHEFSI_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    void *p = chefsi_resident_ptr(ctx, which);
    if (!p) return chefsi_fail(ctx, "ipc_export: block %d is not allocated (chefsi_subspace_reserve[_kpt] first)", which);
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    CHEFSI_CUDA(ctx, cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, p));
    return 0;
}

extern "C" int chefsi_ipc_open(chefsi_ctx_t *ctx, const void *handle64, void **dptr)
{
    if (!ctx || !handle64 || !dptr) return 1;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    CHEFSI_CUDA(ctx, cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int chefsi_ipc_close(chefsi_ctx_t *ctx, void *dptr)
{
    if (!ctx || !dptr) return 1;
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    CHEFSI_CUDA(ctx, cudaIpcCloseMemHandle(dptr));
    return 0;
}

/* Hp and Mp are Hermitian, so every off-diagonal block pair (J, I) / (I, J) has to be formed only once.  Rank I forms
 * block (J, I) -- rows of rank J, its own columns -- when J is one of the ranks that follow it cyclically at a distance
 * below P / 2, and for even P the two ranks of an antipodal pair share that block half and half: a balanced share of
 * (P + 1) / 2 instead of P blocks per rank (the diagonal one counts half: upper-triangle tiles, mirrored).  The host
 * mirrors the rest after the all-gather (sparc_b200/band_parallel.py: assemble_hermitian).  The rule itself:
 * rank_block_part, chefsi_internal.h. */
/* column block `rank` of Mp = Y^H Y and Hp = Y^H H Y: rows = all Ns columns of all ranks, columns = this rank's.
 * peerY[J]: device address of rank J's resident block in THIS process (chefsi_ipc_open, or chefsi_resident_ptr when
 * the ranks share a process); entry `rank` is ignored.  Hp_blk / Mp_blk: host, Ns x ncols[rank], column-major, ld = ldp.
 * share != 0: only the blocks rank_forms_block assigns to this rank are formed, the others are returned as zeros. */
static int rank_project_impl(chefsi_ctx_t *ctx, int is_complex, int nranks, int rank, const int *ncols, void *const *peerY,
                             void *Hp_blk, void *Mp_blk, size_t ldp, int share)
{
    if (!ctx || !Hp_blk || !Mp_blk) return 1;
    int ncol = 0, c0I = 0;
    if (check_ranks(ctx, "rank_project", is_complex, nranks, rank, ncols, &ncol, &c0I)) return 1;
    if (ldp < (size_t)ncol) return chefsi_fail(ctx, "rank_project: ldp smaller than the total number of columns");
    for (int J = 0; J < nranks; J++)
        if (J != rank && ncols[J] > 0 && (!share || rank_forms_block(J, rank, nranks)) && (!peerY || !peerY[J]))
            return chefsi_fail(ctx, "rank_project: no address for rank %d's block", J);
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    const int words = is_complex ? 2 : 1, ncI = ncols[rank];
    const size_t esz = sizeof(double) * words;
    RankBufs *rb = rank_bufs(ctx, (size_t)ncol * ncI * esz);
    if (!rb) return chefsi_fail(ctx, "rank_project: no device memory for the %d x %d blocks", ncol, ncI);
    if (share) {
        CHEFSI_CUDA(ctx, cudaMemsetAsync(rb->d[0], 0, (size_t)ncol * ncI * esz, ctx->stream));
        CHEFSI_CUDA(ctx, cudaMemsetAsync(rb->d[1], 0, (size_t)ncol * ncI * esz, ctx->stream));
    }
    /* W_I = H Y_I (c = 0): local */
    if (is_complex ? chefsi_hamiltonian_mult_kpt_device(ctx, ncI, 0.0, ctx->d_res_Y, ctx->d_res_W)
                   : chefsi_hamiltonian_mult_device(ctx, ncI, 0.0, (const double *)ctx->d_res_Y, (double *)ctx->d_res_W))
        return 1;
    const size_t K = ctx->Nd * words, ldv = ctx->ld * words;
    double *dHp = (double *)rb->d[0], *dMp = (double *)rb->d[1];
    const double *Yi = (const double *)ctx->d_res_Y, *Wi = (const double *)ctx->d_res_W;
    for (int pass = 0; pass < (is_complex ? 2 : 1); pass++) {
        const double *By = Yi, *Bw = Wi;
        if (pass == 1) { /* imaginary parts: Im(A^H B) = A_view^T (-i B)_view, -i B formed locally (subspace.cu) */
            if (launch_rot90(ctx, ctx->d_res_Y, ctx->d_res_T, ctx->Nd, ctx->ld, ncI, -1.0) < 0) return 1;
            By = (const double *)ctx->d_res_T;
        }
        for (int which = 0; which < 2; which++) { /* 0: Mp (B = Y_I), 1: Hp (B = H Y_I) */
            if (pass == 1 && which == 1) {
                CHEFSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); /* d_res_T is reused */
                if (launch_rot90(ctx, ctx->d_res_W, ctx->d_res_T, ctx->Nd, ctx->ld, ncI, -1.0) < 0) return 1;
                Bw = (const double *)ctx->d_res_T;
            }
            const double *B = which ? Bw : By;
            double *Cblk = (which ? dHp : dMp) + pass;
            int c0J = 0;
            for (int J = 0; J < nranks; c0J += ncols[J], J++) {
                if (ncols[J] <= 0) continue;
                int r0 = 0, r1 = ncols[J], q0 = 0, q1 = ncI; /* the part of block (J, rank) this rank forms */
                if (share) rank_block_part(J, rank, nranks, ncols[J], ncI, &r0, &r1, &q0, &q1);
                if (r1 <= r0 || q1 <= q0) continue;
                const double *A = J == rank ? Yi : (const double *)peerY[J]; /* another rank's block: IPC / peer memory */
                /* the diagonal block is a Hermitian product of its own: upper-triangle tiles, mirrored (real part
                   symmetric, imaginary part antisymmetric) */
                const int sym = (share && J == rank) ? (pass == 0 ? +1 : -1) : 0;
                const int nl = launch_gemm_tn(ctx, A + (size_t)r0 * ldv, ldv, B + (size_t)q0 * ldv, ldv, r1 - r0, q1 - q0, K, 1.0,
                                              Cblk + ((size_t)q0 * ncol + c0J + r0) * words, ncol, words, sym);
                if (nl < 0) return 1;
                ctx->stats.kernel_launches += nl;
            }
        }
    }
    const size_t w = (size_t)ncol * esz;
    CHEFSI_CUDA(ctx, cudaMemcpy2DAsync(Mp_blk, ldp * esz, dMp, w, w, ncI, cudaMemcpyDeviceToHost, ctx->stream));
    CHEFSI_CUDA(ctx, cudaMemcpy2DAsync(Hp_blk, ldp * esz, dHp, w, w, ncI, cudaMemcpyDeviceToHost, ctx->stream));
    CHEFSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int chefsi_rank_project(chefsi_ctx_t *ctx, int is_complex, int nranks, int rank, const int *ncols, void *const *peerY,
                                   void *Hp_blk, void *Mp_blk, size_t ldp)
{
    return rank_project_impl(ctx, is_complex, nranks, rank, ncols, peerY, Hp_blk, Mp_blk, ldp, 0);
}

/* the same with every Hermitian block pair formed once: blocks of other ranks' shares come back as zeros */
extern "C" int chefsi_rank_project_shared(chefsi_ctx_t *ctx, int is_complex, int nranks, int rank, const int *ncols,
                                          void *const *peerY, void *Hp_blk, void *Mp_blk, size_t ldp)
{
    return rank_project_impl(ctx, is_complex, nranks, rank, ncols, peerY, Hp_blk, Mp_blk, ldp, 1);
}

/* 1 when rank `rank` of `nranks` forms block (rows of rank J, its own columns) in chefsi_rank_project_shared */
extern "C" int chefsi_rank_forms_block(int J, int rank, int nranks)
{
    return (nranks > 0 && J >= 0 && J < nranks && rank >= 0 && rank < nranks && rank_forms_block(J, rank, nranks)) ? 1 : 0;
}

/* the part of block (rows of rank J, columns of `rank`) that `rank` forms: local rows [r0, r1) x columns [c0, c1) */
extern "C" void chefsi_rank_block_part(int J, int rank, int nranks, int ncJ, int ncI, int *r0, int *r1, int *c0, int *c1)
{
    rank_block_part(J, rank, nranks, ncJ, ncI, r0, r1, c0, c1);
}

/* complex data: T_I = i Y_I, which the other ranks read next to Y_I (Y Q = Y Q_r + (i Y) Q_i on the real views) */
extern "C" int chefsi_rank_rotate_prepare(chefsi_ctx_t *ctx, int is_complex)
{
    if (!ctx) return 1;
    if (ctx->multi) return chefsi_fail(ctx, "rank_rotate_prepare: takes a single-device context");
    if (ctx->res_ncol <= 0) return chefsi_fail(ctx, "rank_rotate_prepare: no resident block");
    if (!is_complex) return 0;
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    if (launch_rot90(ctx, ctx->d_res_Y, ctx->d_res_T, ctx->Nd, ctx->ld, ctx->res_ncol, 1.0) < 0) return 1;
    CHEFSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

/* X_I = sum_J Y_J Q[J, I]: Q_blk = the columns of Q this rank owns (host, Ns x ncols[rank], ld = ldq), peerT only for
 * complex data.  X_blk: host, ncols[rank] columns with leading dimension ldx. */
extern "C" int chefsi_rank_rotate(chefsi_ctx_t *ctx, int is_complex, int nranks, int rank, const int *ncols, void *const *peerY,
                                  void *const *peerT, const void *Q_blk, size_t ldq, void *X_blk, size_t ldx)
{
    if (!ctx || !Q_blk || !X_blk) return 1;
    int ncol = 0, c0I = 0;
    if (check_ranks(ctx, "rank_rotate", is_complex, nranks, rank, ncols, &ncol, &c0I)) return 1;
    if (ldq < (size_t)ncol || ldx < ctx->Nd) return chefsi_fail(ctx, "rank_rotate: bad dimensions");
    for (int J = 0; J < nranks; J++)
        if (J != rank && ncols[J] > 0 && (!peerY || !peerY[J] || (is_complex && (!peerT || !peerT[J]))))
            return chefsi_fail(ctx, "rank_rotate: no address for rank %d's block", J);
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    const int words = is_complex ? 2 : 1, ncI = ncols[rank];
    const size_t esz = sizeof(double) * words, w = (size_t)ncol * esz;
    RankBufs *rb = rank_bufs(ctx, (size_t)ncol * ncI * esz);
    if (!rb) return chefsi_fail(ctx, "rank_rotate: no device memory for the %d x %d block of Q", ncol, ncI);
    double *dQ = (double *)rb->d[2];
    CHEFSI_CUDA(ctx, cudaMemcpy2DAsync(dQ, w, Q_blk, ldq * esz, w, ncI, cudaMemcpyHostToDevice, ctx->stream));
    const double *Qr = dQ, *Qi = nullptr;
    if (is_complex) {
        if (launch_split_complex(ctx, dQ, ncol, ncol, ncI, (double *)rb->d[3], (double *)rb->d[4]) < 0) return 1;
        Qr = (const double *)rb->d[3];
        Qi = (const double *)rb->d[4];
    }
    const size_t K = ctx->Nd * words, ldv = ctx->ld * words;
    int first = 1, c0J = 0;
    for (int J = 0; J < nranks; c0J += ncols[J], J++) {
        if (ncols[J] <= 0) continue;
        const double *Yj = J == rank ? (const double *)ctx->d_res_Y : (const double *)peerY[J];
        int nl = launch_gemm_nn(ctx, Yj, ldv, Qr + c0J, ncol, K, ncols[J], ncI, (double *)ctx->d_res_W, ldv, first ? 0 : 1);
        if (nl < 0) return 1;
        ctx->stats.kernel_launches += nl;
        if (is_complex) {
            const double *Tj = J == rank ? (const double *)ctx->d_res_T : (const double *)peerT[J];
            nl = launch_gemm_nn(ctx, Tj, ldv, Qi + c0J, ncol, K, ncols[J], ncI, (double *)ctx->d_res_W, ldv, 1);
            if (nl < 0) return 1;
            ctx->stats.kernel_launches += nl;
        }
        first = 0;
    }
    CHEFSI_CUDA(ctx, cudaMemcpy2DAsync(X_blk, ldx * esz, ctx->d_res_W, ctx->ld * esz, ctx->Nd * esz, ncI, cudaMemcpyDeviceToHost, ctx->stream));
    CHEFSI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
