/*
 * lanczos.cu -- Lanczos estimate of the extreme eigenvalues of H with every vector resident on the device
 * (SURVEY.md 8f-2), real (chefsi_lanczos) and complex / k-point (chefsi_lanczos_kpt) data.
 *
 * Lanczos_kpt (src/eigenSolverKpt.c:1361-1566) is the same loop on complex vectors, and its inner products are REAL:
 * VectorDotProduct_complex accumulates conj(a_i) b_i into a double (src/tools.c:815-826, the imaginary part is dropped)
 * and Vector2Norm_complex is the 2-norm.  Both are the real dot product / norm of the interleaved (re, im) view of
 * length 2 Nd, and a, b are real, so the only complex step is the H apply: the kernels below run unchanged on the
 * real view.
 *
 * Replaces the body of Lanczos (src/eigenSolver.c:1920-2129) at one rank: start from x0 / ||x0||, three-term
 * recurrence V_{j+1} = H V_j - a_{j+1} V_j - b_j V_{j-1}, a = <V_j, H V_j>, b = ||V_{j+1}||, and after every step the
 * extreme eigenvalues of the tridiagonal T = tridiag(b, a, b) (the reference calls LAPACKE_dsterf, :2075); stop when
 * both moved by less than their tolerances.  The reference ships the vector through Hamiltonian_vectors_mult once
 * per iteration (with the drop-in: one H2D + one D2H per iteration); here only x0 goes in and two numbers come out.
 * Per iteration: one H apply (the filter's kernels), one fused dot kernel, one fused update + norm kernel, one scale
 * kernel, and one 16-byte copy of (a, b) for the host-side convergence test.
 */
#include <cmath>
#include <cstring>
#include <vector>

#include "chefsi_internal.h"

namespace {

constexpr int kBlocks = 296, kThreads = 256;

__device__ __forceinline__ double block_sum(double v, double *sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
    if (w == 0) {
        s = (l < kThreads / 32) ? sh[l] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    }
    return s; /* valid in thread 0 */
}

/* last block to finish adds the per-block partials in index order (deterministic) and writes op(result) */
template <int OP /* 0: sum, 1: sqrt(sum) */>
__device__ __forceinline__ void finish(double part, double *partials, unsigned int *ticket, double *out)
{
    __shared__ bool last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = part;
        __threadfence();
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        double s = 0.0;
        for (unsigned i = 0; i < gridDim.x; i++) s += ((volatile double *)partials)[i];
        *out = OP ? sqrt(s) : s;
        *ticket = 0;
    }
}

/* out = <x, y> */
__global__ void __launch_bounds__(kThreads) dot_kernel(const double *__restrict__ x, const double *__restrict__ y, size_t n,
                                                       double *partials, unsigned int *ticket, double *out)
{
    __shared__ double sh[kThreads / 32];
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) s = fma(x[i], y[i], s);
    finish<0>(block_sum(s, sh), partials, ticket, out);
}

/* w -= a v + b u ; u = v ; out = ||w||      (a read from device memory, b a value known on the host)
   first step: u == nullptr -> w -= a v only */
__global__ void __launch_bounds__(kThreads) update_kernel(double *__restrict__ w, const double *__restrict__ v, double *__restrict__ u,
                                                          const double *__restrict__ a_ptr, double b, size_t n, double *partials,
                                                          unsigned int *ticket, double *out)
{
    __shared__ double sh[kThreads / 32];
    const double a = *a_ptr;
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        const double vi = v[i];
        double wi = w[i];
        if (u) { wi -= (a * vi + b * u[i]); u[i] = vi; }
        else wi -= a * vi;
        w[i] = wi;
        s = fma(wi, wi, s);
    }
    finish<1>(block_sum(s, sh), partials, ticket, out);
}

/* y = x * (*s_ptr == 0 ? 1 : 1 / *s_ptr)   (or x * s when direct) */
__global__ void scale_kernel(double *__restrict__ y, const double *__restrict__ x, const double *__restrict__ s_ptr, size_t n)
{
    const double s = *s_ptr, f = (s == 0.0) ? 1.0 : 1.0 / s;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = x[i] * f;
}

/* smallest and largest eigenvalue of the symmetric tridiagonal matrix (diagonal d[0..n), off-diagonal e[0..n-1)) by
   bisection on the Sturm count */
void tridiag_extremes(const std::vector<double> &d, const std::vector<double> &e, int n, double *lo_out, double *hi_out)
{
    double glo = d[0], ghi = d[0];
    for (int i = 0; i < n; i++) {
        const double r = (i > 0 ? fabs(e[i - 1]) : 0.0) + (i < n - 1 ? fabs(e[i]) : 0.0);
        glo = fmin(glo, d[i] - r);
        ghi = fmax(ghi, d[i] + r);
    }
    auto count_below = [&](double x) { /* number of eigenvalues < x */
        int c = 0;
        double q = d[0] - x;
        if (q < 0) c++;
        for (int i = 1; i < n; i++) {
            if (q == 0.0) q = 1e-300;
            q = d[i] - x - e[i - 1] * e[i - 1] / q;
            if (q < 0) c++;
        }
        return c;
    };
    auto kth = [&](int k) { /* k-th smallest, 0-based */
        double a = glo, b = ghi;
        for (int it = 0; it < 200; it++) {
            const double m = 0.5 * (a + b);
            if (m <= a || m >= b) break;
            if (count_below(m) > k) b = m; else a = m;
        }
        return 0.5 * (a + b);
    };
    *lo_out = kth(0);
    *hi_out = kth(n - 1);
}

}  // namespace

int apply_h_device(chefsi_ctx *ctx, const void *x, void *Hx, bool is_complex); /* chefsi_api.cu: one H apply (c = 0) of a resident column */

namespace {

/* words = 1: real column of Nd doubles; words = 2: complex column seen as 2 Nd doubles */
int lanczos_impl(chefsi_ctx *ctx, const void *x0, int words, double tol_min, double tol_max, int maxit, double *eigmin, double *eigmax,
                 int *iterations)
{
    if (!ctx || !x0 || !eigmin || !eigmax) return 1;
    if (ctx->multi) { /* a single vector does not split over devices: the first one iterates */
        chefsi_ctx *k = multi_first(ctx);
        const int rc = lanczos_impl(k, x0, words, tol_min, tol_max, maxit, eigmin, eigmax, iterations);
        if (rc) { strncpy(ctx->err, k->err, sizeof(ctx->err) - 1); ctx->err[sizeof(ctx->err) - 1] = 0; }
        return rc;
    }
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (maxit < 1) return chefsi_fail(ctx, "lanczos: maxit must be positive");
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool cplx = words == 2;
    const size_t n = (size_t)ctx->Nd * words, ldb = (size_t)ctx->ld * words * sizeof(double);
    /* workspace: 3 vectors, scalars a[maxit+1], b[maxit+1], partials, ticket */
    const size_t scal = (size_t)2 * (maxit + 2) + kBlocks + 8;
    const size_t need = 3 * ldb + scal * sizeof(double);
    if (need > ctx->lanczos_bytes) {
        cudaFree(ctx->d_lanczos);
        ctx->d_lanczos = nullptr;
        ctx->lanczos_bytes = 0;
        CHEFSI_CUDA(ctx, cudaMalloc(&ctx->d_lanczos, need));
        ctx->lanczos_bytes = need;
    }
    char *base = (char *)ctx->d_lanczos;
    double *Vjm1 = (double *)base, *Vj = (double *)(base + ldb), *Vjp1 = (double *)(base + 2 * ldb);
    double *da = (double *)(base + 3 * ldb), *db = da + (maxit + 2), *partials = db + (maxit + 2);
    unsigned int *ticket = (unsigned int *)(partials + kBlocks);
    double *tmp = partials + kBlocks + 2;
    cudaStream_t st = ctx->stream;
    CHEFSI_CUDA(ctx, cudaMemsetAsync(ticket, 0, 2 * sizeof(double), st));
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(Vjp1, x0, n * sizeof(double), cudaMemcpyHostToDevice, st));
    /* V_{j-1} = x0 / ||x0||                                                         (eigenSolver.c:1986-1992) */
    dot_kernel<<<kBlocks, kThreads, 0, st>>>(Vjp1, Vjp1, n, partials, ticket, tmp);
    double h_ab[2];
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(h_ab, tmp, sizeof(double), cudaMemcpyDeviceToHost, st));
    CHEFSI_CUDA(ctx, cudaStreamSynchronize(st));
    const double nrm0 = sqrt(h_ab[0]);
    if (!(nrm0 > 0.0)) return chefsi_fail(ctx, "lanczos: zero start vector");
    h_ab[0] = nrm0;
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(tmp, h_ab, sizeof(double), cudaMemcpyHostToDevice, st));
    scale_kernel<<<kBlocks, kThreads, 0, st>>>(Vjm1, Vjp1, tmp, n);
    /* V_j = H V_{j-1}; a0 = <V_{j-1}, V_j>; V_j -= a0 V_{j-1}; b0 = ||V_j||; V_j /= b0   (:2003-2040) */
    if (apply_h_device(ctx, Vjm1, Vjp1, cplx)) return 1;
    dot_kernel<<<kBlocks, kThreads, 0, st>>>(Vjm1, Vjp1, n, partials, ticket, da);
    update_kernel<<<kBlocks, kThreads, 0, st>>>(Vjp1, Vjm1, nullptr, da, 0.0, n, partials, ticket, db);
    scale_kernel<<<kBlocks, kThreads, 0, st>>>(Vj, Vjp1, db, n);
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(&h_ab[0], da, sizeof(double), cudaMemcpyDeviceToHost, st));
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(&h_ab[1], db, sizeof(double), cudaMemcpyDeviceToHost, st));
    CHEFSI_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->stats.kernel_launches += 5;
    std::vector<double> a(maxit + 2, 0.0), b(maxit + 2, 0.0);
    a[0] = h_ab[0];
    b[0] = h_ab[1];
    if (b[0] == 0.0) return chefsi_fail(ctx, "lanczos: the start vector is an eigenvector (H x0 parallel to x0)"); /* the reference re-randomises (:2020-2033); the caller falls back to it */
    double emin = 0.0, emax = 0.0, emin_pre = 0.0, emax_pre = 0.0;
    double err_min = tol_min + 1.0, err_max = tol_max + 1.0;
    int j = 0;
    while ((err_min > tol_min || err_max > tol_max) && j < maxit) {
        if (apply_h_device(ctx, Vj, Vjp1, cplx)) return 1;                                   /* V_{j+1} = H V_j          (:2048) */
        dot_kernel<<<kBlocks, kThreads, 0, st>>>(Vj, Vjp1, n, partials, ticket, da + j + 1);  /* a[j+1] = <V_j, V_{j+1}>  (:2054) */
        update_kernel<<<kBlocks, kThreads, 0, st>>>(Vjp1, Vj, Vjm1, da + j + 1, b[j], n, partials, ticket, db + j + 1); /* (:2056-2063) */
        scale_kernel<<<kBlocks, kThreads, 0, st>>>(Vj, Vjp1, db + j + 1, n);                  /* V_j = V_{j+1} / b[j+1]   (:2068-2071) */
        CHEFSI_CUDA(ctx, cudaMemcpyAsync(&h_ab[0], da + j + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
        CHEFSI_CUDA(ctx, cudaMemcpyAsync(&h_ab[1], db + j + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
        CHEFSI_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->stats.kernel_launches += 3;
        a[j + 1] = h_ab[0];
        b[j + 1] = h_ab[1];
        if (b[j + 1] == 0.0) break;                                                          /* (:2064-2066) */
        tridiag_extremes(a, b, j + 2, &emin, &emax);                                         /* (:2073-2083) */
        err_min = fabs(emin - emin_pre);
        err_max = fabs(emax - emax_pre);
        emin_pre = emin;
        emax_pre = emax;
        j++;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return chefsi_fail(ctx, "lanczos kernels: %s", cudaGetErrorString(e));
    *eigmin = emin;
    *eigmax = emax;
    if (iterations) *iterations = j;
    return 0;
}

}  // namespace

extern "C" int chefsi_lanczos(chefsi_ctx_t *ctx, const double *x0, double tol_min, double tol_max, int maxit, double *eigmin,
                              double *eigmax, int *iterations)
{
    return lanczos_impl(ctx, x0, 1, tol_min, tol_max, maxit, eigmin, eigmax, iterations);
}

extern "C" int chefsi_lanczos_kpt(chefsi_ctx_t *ctx, const void *x0, double tol_min, double tol_max, int maxit, double *eigmin,
                                  double *eigmax, int *iterations)
{
    return lanczos_impl(ctx, x0, 2, tol_min, tol_max, maxit, eigmin, eigmax, iterations);
}
