/*
 * chefsi_internal.h -- internal declarations shared by the CUDA translation units of
 * libchefsi_b200.so.  Not part of the C ABI (that is include/chefsi_b200.h).
 */
#ifndef CHEFSI_INTERNAL_H
#define CHEFSI_INTERNAL_H

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/chefsi_b200.h"

#define CHEFSI_MAXR CHEFSI_MAX_FDN

/* ---- stencil description handed to the kernels by value ------------------------------ */
struct MixedComp {
    int ext, ax1, ax2;        /* axes 0:x 1:y 2:z ; ax2 = -1 when the inner derivative is single */
    double c1[CHEFSI_MAXR + 1];
    double c2[CHEFSI_MAXR + 1];
    double wm[CHEFSI_MAXR + 1]; /* already scaled by a = -1/2 */
};

/* Device-resident layout of one orbital column ("internal layout") = the reference's dense x-fastest layout
 * (lapVecRoutines.c:313): grid point (i,j,k) lives at (k*Ny + j)*Nx + i; columns are ld elements apart
 * (ld >= Nd, a multiple of 16 doubles because TMA strides are multiples of 16 bytes).  Halos are never stored:
 * the streaming kernels get periodic images through wrapped TMA boxes and Dirichlet zeros through the TMA
 * out-of-bounds fill, the z-march kernels wrap when they commit a plane to their ring. */
struct Layout {
    int Nx, Ny, Nz;
    size_t plane;   /* Nx * Ny */
    size_t ld;      /* elements between columns (>= plane * Nz, multiple of 16) */
};
__host__ __device__ __forceinline__ size_t lay_pos(const Layout &L, int i, int j, int k)
{
    return ((size_t)k * L.Ny + j) * L.Nx + i;
}

struct StencilDesc {
    Layout lay;
    int Nx, Ny, Nz;
    int bc[3];
    int F;
    int nmix;
    double coef0;               /* a * (D2x[0] + D2y[0] + D2z[0]) */
    double wx[CHEFSI_MAXR + 1]; /* a * D2_x[r] ... */
    double wy[CHEFSI_MAXR + 1];
    double wz[CHEFSI_MAXR + 1];
    MixedComp mix[2];
    double ph_re[27], ph_im[27]; /* Bloch phases exp(i k.(ox Lx, oy Ly, oz Lz)), index (oz+1)*9+(oy+1)*3+(ox+1) */
};

/* out = s1 * ((-1/2 Lap + veff + c) x) - s2 * xprev      (xprev may be NULL when s2 == 0)
 * one launch handles ncol columns with leading dimension ld (elements of T).               */
struct StepArgs {
    const void *x;
    const void *xprev;
    void *out;
    const double *veff;
    size_t ld;
    int ncol;
    double c, s1, s2;
};

/* ---- nonlocal projector tables on the device ----------------------------------------- */
struct NlocDev {
    int n_atom = 0, n_img = 0, ntot = 0, max_nproj = 0;
    int overlap = 1;           /* 1: some grid point lies in more than one sphere -> atomics */
    int *IP_displ = nullptr;   /* [n_atom+1] */
    double *gamma = nullptr;   /* [ntot]     */
    int *img_atom = nullptr;   /* [n_img]    */
    int *img_ndc = nullptr;
    long long *pos_off = nullptr, *chiT_off = nullptr;
    int *grid_pos = nullptr;
    double *chiT = nullptr;       /* per image: ndc x np_pad, POINT-major (projector fastest), zero padded */
    int *img_aoff = nullptr;      /* [n_img+1] prefix sum of nproj over images: per-image alpha partials */
    int np_pad = 0;               /* projector padding the nloc kernels are instantiated for */
    long long img_proj_total = 0; /* img_aoff[n_img] */
    int max_parts = 0;            /* largest number of images/segments (= alpha partials) of any atom */
    double2 *img_phase = nullptr; /* [n_img] (cos theta, sin theta) for the current k-point */
    int *atom_img_off = nullptr;  /* CSR atom -> images */
    int *atom_img = nullptr;
    /* host copies needed to rebuild phases */
    double *h_img_coords = nullptr;
    long long total_pts = 0;      /* elements of grid_pos */
    long long n_chiT = 0;         /* elements of chiT */
};

struct MultiState; /* multi.cu */

struct chefsi_ctx {
    MultiState *multi = nullptr; /* != NULL: a leader context that owns one single-device child per GPU (multi.cu) */
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    bool have_grid = false;
    chefsi_grid_t grid{};
    StencilDesc desc{};
    size_t Nd = 0, ld = 0;
    Layout lay{};
    double *d_veff = nullptr;    /* Veff in the internal layout (one column) */
    cudaEvent_t pipe_ev[12] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool have_veff = false;
    NlocDev nl;
    double kvec[3] = {0, 0, 0};
    /* scratch owned by the host entry points */
    void *d_buf[3] = {nullptr, nullptr, nullptr};
    void *d_buf2[3] = {nullptr, nullptr, nullptr}; /* second and third trio of the host pipeline (chunks rotate through three) */
    void *d_buf3[3] = {nullptr, nullptr, nullptr};
    size_t buf2_bytes = 0;
    size_t buf_bytes = 0;
    /* small blocks in pageable host memory (the SCF test systems; single columns of Lanczos / the Poisson residual) go
       through these pinned staging buffers: one memcpy + one truly asynchronous copy instead of the driver's
       pageable path */
    /* subspace.cu: the filtered block kept on the device between ChebyshevFiltering, the projection and the rotation */
    void *d_res_Y = nullptr, *d_res_W = nullptr;   /* Y and a work block (H Y, then Y Q), ncol x ld each */
    void *d_res_T = nullptr;                       /* complex data only: -i H Y, then i Y */
    size_t res_bytes = 0, res_t_bytes = 0;
    int res_complex = 0;                           /* the resident Y is complex */
    int res_ncol = 0;                              /* columns of the resident Y (0: none) */
    const void *res_host = nullptr;                /* host address the resident Y stands for */
    const void *res_unwritten_host = nullptr;      /* host block whose copy-back was skipped (NO_Y_COPYBACK): its contents are stale */
    void *rank_state = nullptr;                    /* ranks.cu: Hp / Mp / Q column blocks of a one-process-per-GPU run */
    void *d_aar = nullptr;                         /* aar.cu: x, b, r, x_old, f, f_old, the two histories, scalars */
    size_t aar_bytes = 0;
    void *d_lanczos = nullptr;                     /* lanczos.cu: three vectors + scalars */
    size_t lanczos_bytes = 0;
    void *d_gemm_ws = nullptr;                     /* split-K partial tiles */
    size_t gemm_ws_bytes = 0;
    void *d_small[3] = {nullptr, nullptr, nullptr}; /* Hp, Mp, Q (Ns x Ns) */
    size_t small_bytes = 0;
    /* rayleigh_ritz.cu */
    int small_ncol = 0, small_complex = 0;         /* d_small[0], d_small[1] hold Hp, Mp of the last projection (0: not) */
    int q_ncol = 0, q_complex = 0;                 /* d_small[2] holds the eigenvectors of the last chefsi_subspace_eig */
    struct EigState *eig = nullptr;                /* cuSOLVER handle and workspace */
    struct BandStore *bands = nullptr;             /* device copies of the rotated blocks for the density */
    void *h_pin[3] = {nullptr, nullptr, nullptr};
    size_t h_pin_bytes = 0;
    int gemm_big_tiles = 1;                        /* subspace.cu: 128 x 128 CTA tiles for more than 64 columns (0: 64 x 64 everywhere) */
    int gemm_symmetric = 1;                        /* Y^T Y and Y^T H Y: upper-triangle tiles only, mirrored */
    int fast_small = 1;
    int small_brick = 1;                           /* launches with very few z-marching CTAs take the 3-D brick kernel */
    void *d_alpha[2] = {nullptr, nullptr}; /* per-image alpha partials: [cur] belongs to the current input */
    int alpha_sum_external = 0;            /* 1: d_alpha_sum was produced by chefsi_nloc_project_device: expand must not rebuild it */
    int alpha_reduce_min = 8;              /* atoms with more alpha partials than this get them summed by alpha_reduce_kernel */
    size_t alpha_sum_bytes = 0;
    void *d_alpha_sum = nullptr;           /* per-atom sums of the partials (only when nl.max_parts is large) */
    int alpha_cur = 0;
    size_t alpha_bytes = 0;
    int nloc_sort = 1;                     /* projector CTAs in order of decreasing sphere-segment size (CHEFSI_B200_NLOC_SORT) */
    /* stats */
    chefsi_stats_t stats{};
    int profiling = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int num_sms = 148;
    size_t max_smem_optin = 0;
    int force_general = 0;         /* 1: no TMA streaming kernel; 2: neither streaming nor z-march (3-D brick kernel only) */
    int stream_gridsync = 1;       /* round barrier between the streaming kernel's producers */
    int tma_l2promo = 3;           /* CUtensorMapL2promotion of the streaming kernel's tensor maps */
    unsigned int *d_sync = nullptr;
    unsigned int sync_arrivals = 0;
    char err[512] = {0};
};

int chefsi_fail(chefsi_ctx *ctx, const char *fmt, ...);
#define CHEFSI_CUDA(ctx, call)                                                                \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return chefsi_fail((ctx), "%s:%d: %s -> %s", __FILE__, __LINE__, #call,           \
                               cudaGetErrorString(e_));                                       \
    } while (0)

/* ---- kernel launchers (each returns the number of kernels it launched, <0 on error) --- */
int launch_stencil_general(chefsi_ctx *ctx, const StepArgs &a, bool is_complex);
bool stream_kpt_supported(const chefsi_ctx *ctx);
int launch_stencil_stream_kpt(chefsi_ctx *ctx, const StepArgs &a);
bool stencil_zmarch_supported(const chefsi_ctx *ctx);
int launch_stencil_zmarch(chefsi_ctx *ctx, const StepArgs &a, bool is_complex);
bool stream_dense_supported(const chefsi_ctx *ctx, bool is_complex);
int launch_stencil_stream_dense(chefsi_ctx *ctx, const StepArgs &a);
bool stream_mixed_supported(const chefsi_ctx *ctx, bool is_complex);
int launch_stencil_stream_mixed(chefsi_ctx *ctx, const StepArgs &a); /* -2: coefficient tables without the expected structure */

/* nloc.cu: see launch_nloc for the three modes */
enum { NLOC_PROJECT = 0, NLOC_FUSED = 1, NLOC_EXPAND = 2 };
int launch_nloc(chefsi_ctx *ctx, int mode, void *vec, size_t ld, int ncol, double scale, bool is_complex);
int nloc_padded_nproj(int max_nproj);
int launch_alpha_reduce(chefsi_ctx *ctx, int ncol, bool is_complex);
int nloc_ensure_alpha(chefsi_ctx *ctx, int ncol, bool is_complex); /* (re)allocate the alpha buffers for ncol columns */

/* multi.cu: the leader context's side of every entry point it supports */
chefsi_ctx *multi_first(chefsi_ctx *lead);
int multi_size(const chefsi_ctx *lead);
int multi_uses_nccl(const chefsi_ctx *lead);
void multi_destroy(chefsi_ctx *lead);
int multi_set_grid(chefsi_ctx *lead, const chefsi_grid_t *g);
int multi_set_kpoint(chefsi_ctx *lead, double k1, double k2, double k3);
int multi_set_veff(chefsi_ctx *lead, const double *veff_host);
int multi_set_projectors(chefsi_ctx *lead, const chefsi_nloc_t *nl);
int multi_filter_host(chefsi_ctx *lead, void *X, size_t ldi, void *Y, size_t ldo, int ncol, int m, double a, double b, double a0,
                      int flags, bool is_complex);
int multi_hmult_host(chefsi_ctx *lead, int ncol, double c, const void *x, size_t ldi, void *Hx, size_t ldo, bool is_complex);
int multi_lapmult_host(chefsi_ctx *lead, int ncol, double a, double c, const void *x, size_t ldi, void *y, size_t ldo, bool is_complex);
int multi_gradmult_host(chefsi_ctx *lead, int ncol, double c, const void *x, size_t ldi, void *Dx, size_t ldo, int dir, double kdir, bool is_complex);
int multi_synchronize(chefsi_ctx *lead);
void multi_set_profiling(chefsi_ctx *lead, int on);
int multi_subspace_reserve(chefsi_ctx *lead, int ncol, bool is_complex);
int multi_subspace_project(chefsi_ctx *lead, const void *Y, size_t ldy, int ncol, void *Hp, void *Mp, size_t ldp, bool is_complex);
int multi_subspace_rotate(chefsi_ctx *lead, const void *Q, size_t ldq, int ncol, void *X, size_t ldx, bool is_complex);
void multi_bcast_stats(const chefsi_ctx *lead, unsigned long long *calls, unsigned long long *bytes);
void chefsi_free_nloc(NlocDev &d);

/* subspace.cu */
int launch_gemm_tn(chefsi_ctx *ctx, const double *A, size_t lda, const double *B, size_t ldb, int M, int N, size_t K, double scale,
                   double *C, size_t ldc, int cstride, int sym = 0);
int launch_gemm_nn(chefsi_ctx *ctx, const double *A, size_t lda, const double *Q, size_t ldq, size_t K, int M, int N, double *C,
                   size_t ldc, int accumulate);
/* ranks.cu */
void rank_state_destroy(chefsi_ctx *ctx);
/* Hp and Mp are Hermitian: of every off-diagonal block pair (J, I) / (I, J) of a column-block split over P ranks or devices
   only one element of each mirrored pair is formed.  Owner I forms block (J, I) -- rows of J, its own columns -- for the
   owners J that follow it cyclically at a distance below P / 2; for even P the block at distance P / 2 is shared by its
   two owners: the lower one (I < J) forms the first half of its columns, the upper one the rows of the lower owner's
   block that mirror the other half.  rank_block_part gives the part of block (J, I) that owner I forms as local
   rows [r0, r1) x columns [c0, c1) (empty: r1 <= r0 or c1 <= c0); every owner forms (P + 1) / 2 blocks' worth. */
inline void rank_block_part(int J, int I, int P, int ncJ, int ncI, int *r0, int *r1, int *c0, int *c1)
{
    *r0 = 0; *r1 = ncJ; *c0 = 0; *c1 = ncI;
    if (J == I) return;
    const int dist = ((J - I) % P + P) % P;
    if (2 * dist < P) return;
    if (2 * dist == P) {
        if (I < J) *c1 = ncI / 2;   /* lower owner: columns [0, ncI / 2) of block (J, I) */
        else *r0 = ncJ / 2;         /* upper owner: rows [ncJ / 2, ncJ) of block (J, I), J being the lower one */
        return;
    }
    *r1 = 0;
}
/* block-level form of the rule: does owner I form anything of block (J, I)? */
inline bool rank_forms_block(int J, int I, int P)
{
    if (J == I) return true;
    const int dist = ((J - I) % P + P) % P;
    return 2 * dist <= P;
}
/* gradient.cu */
int launch_gradient(chefsi_ctx *ctx, const void *x, void *out, int ncol, int dir, double c, double kdir, bool is_complex);
int launch_rot90(chefsi_ctx *ctx, const void *in, void *out, size_t n, size_t ld, int ncol, double s);
int launch_split_complex(chefsi_ctx *ctx, const void *Q, size_t ldq, int M, int N, double *Qr, double *Qi);

/* rayleigh_ritz.cu */
int band_store_put(chefsi_ctx *ctx, const void *host, int ncol, bool is_complex, const void *d_block);
void band_store_invalidate(chefsi_ctx *ctx, const void *host);
void band_store_clear(chefsi_ctx *ctx);
void rayleigh_ritz_destroy(chefsi_ctx *ctx);

/* util.cu */
int launch_fill_random(chefsi_ctx *ctx, void *buf, int ncol, long long first_col, unsigned long long seed,
                       bool is_complex);

#endif
