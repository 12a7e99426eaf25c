/*
 * aar.cu -- Alternating Anderson-Richardson solve of -(Lap + c) x = b with every vector resident on the device
 * (SURVEY.md 8f-4), real data, single-device contexts.
 *
 * Replaces AAR (src/linearSolver.c:38-146) for the one operator pair SPARC uses it with: residual
 * r = b + (Lap + c) x (poisson_residual, src/lapVecRoutines.c:61-79) and the Jacobi preconditioner f = -r / (D2x[0] +
 * D2y[0] + D2z[0] + c) (src/electrostatics.c:1682-1700) -- the Poisson solve of every SCF iteration
 * (electrostatics.c:1658) and the Kerker preconditioner of the mixing (mixing.c:501).  Same iteration, same
 * parameters, same stopping rule:
 *   f = M^-1 r;  history X(:,i) = x - x_old, F(:,i) = f - f_old;  every p-th step the Anderson extrapolation
 *   x = x_old - X G + beta (f - F G), G = pinv(F^T F) F^T f (mixing.c:48-140: dgemm/dgemv + LAPACKE_dgelsd), else the
 *   Richardson update x = x_old + omega f;  r = b + (Lap + c) x;  ||r|| is checked after Anderson steps only.
 * With the drop-in of round 2 every residual evaluation was one H2D + one D2H of a grid vector (Au_fcc211: 5 461 per
 * run); here x and b go in once, x comes out once, and the host sees (m^2 + m + 1) numbers per Anderson step.
 * The Laplacian is the filter's stencil kernel (its `- s2 * xprev` operand carries b: s2 = -1).
 */
#include <cmath>
#include <cstring>
#include <vector>

#include "chefsi_internal.h"

namespace {

constexpr int kBlocks = 296, kThreads = 256, kMaxM = 16;

__device__ __forceinline__ double block_sum(double v, double *sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
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

/* out[y] = <u_y, v_y> for gridDim.y vector pairs; per pair a fixed-order sum of the per-block partials (deterministic) */
struct DotPairs { const double *u[kMaxM * (kMaxM + 1) / 2 + kMaxM + 1]; const double *v[kMaxM * (kMaxM + 1) / 2 + kMaxM + 1]; };

__global__ void __launch_bounds__(kThreads) dots_kernel(const __grid_constant__ DotPairs P, size_t n, double *partials, unsigned int *tickets,
                                                        double *out)
{
    __shared__ double sh[kThreads / 32];
    __shared__ bool last;
    const double *__restrict__ u = P.u[blockIdx.y], *__restrict__ v = P.v[blockIdx.y];
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) s = fma(u[i], v[i], s);
    const double part = block_sum(s, sh);
    double *mine = partials + (size_t)blockIdx.y * gridDim.x;
    if (threadIdx.x == 0) {
        mine[blockIdx.x] = part;
        __threadfence();
        last = (atomicAdd(&tickets[blockIdx.y], 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        double t = 0.0;
        for (unsigned i = 0; i < gridDim.x; i++) t += ((volatile double *)mine)[i];
        out[blockIdx.y] = t;
        tickets[blockIdx.y] = 0;
    }
}

/* f = m_inv r; if (hist >= 0) X_h = x - x_old, F_h = f - f_old; x_old = x; f_old = f; if (richardson) x += omega f */
__global__ void aar_step_kernel(const double *__restrict__ r, double *__restrict__ x, double *__restrict__ x_old, double *__restrict__ f,
                                double *__restrict__ f_old, double *__restrict__ Xh, double *__restrict__ Fh, double m_inv, double omega,
                                int richardson, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double fi = m_inv * r[i], xi = x[i];
        if (Xh) {
            Xh[i] = xi - x_old[i];
            Fh[i] = fi - f_old[i];
        }
        x_old[i] = xi;
        f_old[i] = fi;
        f[i] = fi;
        if (richardson) x[i] = xi + omega * fi;
    }
}

struct Gammas { double g[kMaxM]; };
/* x = x_old - X G + beta (f - F G)      (mixing.c:48-103) */
__global__ void aar_anderson_kernel(double *__restrict__ x, const double *__restrict__ x_old, const double *__restrict__ f,
                                    const double *__restrict__ X, const double *__restrict__ F, const Gammas G, int m, double beta,
                                    size_t n, size_t ldh)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double xw = x_old[i], fw = f[i];
        for (int j = 0; j < m; j++) {
            xw = fma(-G.g[j], X[(size_t)j * ldh + i], xw);
            fw = fma(-G.g[j], F[(size_t)j * ldh + i], fw);
        }
        x[i] = xw + beta * fw;
    }
}

/* G = pinv(A) rhs for the symmetric positive semi-definite m x m matrix A (column-major), singular values below
   eps * largest dropped -- what LAPACKE_dgelsd(..., rcond = -1) returns for such a matrix (mixing.c:121).  Cyclic Jacobi. */
void pinv_solve(int m, std::vector<double> A, const std::vector<double> &rhs, double *G)
{
    std::vector<double> V((size_t)m * m, 0.0);
    for (int i = 0; i < m; i++) V[(size_t)i * m + i] = 1.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int p = 0; p < m; p++)
            for (int q = p + 1; q < m; q++) off += A[(size_t)q * m + p] * A[(size_t)q * m + p];
        if (off < 1e-300) break;
        for (int p = 0; p < m; p++)
            for (int q = p + 1; q < m; q++) {
                const double apq = A[(size_t)q * m + p];
                if (fabs(apq) < 1e-300) continue;
                const double app = A[(size_t)p * m + p], aqq = A[(size_t)q * m + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < m; k++) { /* columns p, q of A */
                    const double akp = A[(size_t)p * m + k], akq = A[(size_t)q * m + k];
                    A[(size_t)p * m + k] = c * akp - s * akq;
                    A[(size_t)q * m + k] = s * akp + c * akq;
                }
                for (int k = 0; k < m; k++) { /* rows p, q */
                    const double apk = A[(size_t)k * m + p], aqk = A[(size_t)k * m + q];
                    A[(size_t)k * m + p] = c * apk - s * aqk;
                    A[(size_t)k * m + q] = s * apk + c * aqk;
                }
                for (int k = 0; k < m; k++) {
                    const double vkp = V[(size_t)p * m + k], vkq = V[(size_t)q * m + k];
                    V[(size_t)p * m + k] = c * vkp - s * vkq;
                    V[(size_t)q * m + k] = s * vkp + c * vkq;
                }
            }
    }
    double lmax = 0.0;
    for (int k = 0; k < m; k++) lmax = fmax(lmax, fabs(A[(size_t)k * m + k]));
    const double cut = lmax * 2.220446049250313e-16;
    for (int i = 0; i < m; i++) G[i] = 0.0;
    for (int k = 0; k < m; k++) {
        const double lam = A[(size_t)k * m + k];
        if (!(fabs(lam) > cut)) continue;
        double proj = 0.0;
        for (int i = 0; i < m; i++) proj += V[(size_t)k * m + i] * rhs[i];
        proj /= lam;
        for (int i = 0; i < m; i++) G[i] += V[(size_t)k * m + i] * proj;
    }
}

}  // namespace

int lap_residual_device(chefsi_ctx *ctx, double c, const void *x, const void *b, void *r); /* chefsi_api.cu: r = b + (Lap + c) x */

extern "C" int chefsi_poisson_aar(chefsi_ctx_t *ctx, double c, double *x, const double *b, double omega, double beta, int m, int p,
                                  double tol, int max_iter, int *iterations, double *res_norm)
{
    if (!ctx || !x || !b) return 1;
    if (ctx->multi) { /* a single right-hand side does not split over devices: the first one solves it */
        chefsi_ctx *k = multi_first(ctx);
        const int rc = chefsi_poisson_aar(k, c, x, b, omega, beta, m, p, tol, max_iter, iterations, res_norm);
        if (rc) { strncpy(ctx->err, k->err, sizeof(ctx->err) - 1); ctx->err[sizeof(ctx->err) - 1] = 0; }
        return rc;
    }
    if (!ctx->have_grid) return chefsi_fail(ctx, "set_grid must be called first");
    if (m < 1 || m > kMaxM || p < 1) return chefsi_fail(ctx, "aar: history length must be 1..%d, p >= 1", kMaxM);
    CHEFSI_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = ctx->Nd, ldh = ctx->ld, vb = ldh * sizeof(double);
    const int npair = m * (m + 1) / 2 + m; /* F^T F (upper triangle) and F^T f */
    const size_t nscal = (size_t)(npair + 2) * kBlocks + 2 * (npair + 2) + 64;
    const size_t need = (6 + 2 * (size_t)m) * vb + nscal * sizeof(double);
    if (need > ctx->aar_bytes) {
        cudaFree(ctx->d_aar);
        ctx->d_aar = nullptr;
        ctx->aar_bytes = 0;
        CHEFSI_CUDA(ctx, cudaMalloc(&ctx->d_aar, need));
        ctx->aar_bytes = need;
    }
    char *base = (char *)ctx->d_aar;
    double *dx = (double *)base, *db = (double *)(base + vb), *dr = (double *)(base + 2 * vb), *dxo = (double *)(base + 3 * vb);
    double *df = (double *)(base + 4 * vb), *dfo = (double *)(base + 5 * vb);
    double *dX = (double *)(base + 6 * vb), *dF = (double *)(base + (6 + (size_t)m) * vb);
    double *partials = (double *)(base + (6 + 2 * (size_t)m) * vb);
    double *outs = partials + (size_t)(npair + 2) * kBlocks;
    unsigned int *tickets = (unsigned int *)(outs + (npair + 2));
    cudaStream_t st = ctx->stream;
    CHEFSI_CUDA(ctx, cudaMemsetAsync(tickets, 0, (npair + 2) * sizeof(double), st));
    CHEFSI_CUDA(ctx, cudaMemsetAsync(dX, 0, 2 * (size_t)m * vb, st)); /* the reference callocs the histories (linearSolver.c:66-67) */
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(dx, x, n * sizeof(double), cudaMemcpyHostToDevice, st));
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(db, b, n * sizeof(double), cudaMemcpyHostToDevice, st));
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(dxo, dx, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    /* Jacobi preconditioner (electrostatics.c:1689-1697) */
    const chefsi_grid_t &g = ctx->grid;
    double m_inv = g.D2_x[0] + g.D2_y[0] + g.D2_z[0] + c;
    if (fabs(m_inv) < 1e-14) m_inv = 1.0;
    m_inv = -1.0 / m_inv;
    std::vector<double> h((size_t)npair + 2, 0.0);
    DotPairs P;
    /* ||b|| and the first residual */
    P.u[0] = db; P.v[0] = db;
    dots_kernel<<<dim3(kBlocks, 1), kThreads, 0, st>>>(P, n, partials, tickets, outs);
    if (lap_residual_device(ctx, c, dx, db, dr)) return 1;
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(h.data(), outs, sizeof(double), cudaMemcpyDeviceToHost, st));
    CHEFSI_CUDA(ctx, cudaStreamSynchronize(st));
    const double b_2norm = sqrt(h[0]);
    tol *= b_2norm;
    double r_2norm = tol + 1.0;
    int iter = 0;
    ctx->stats.kernel_launches += 1;
    while (r_2norm > tol && iter < max_iter) {
        const int i_hist = iter > 0 ? (iter - 1) % m : -1;
        const bool anderson = ((iter + 1) % p == 0) && iter > 0;
        aar_step_kernel<<<kBlocks, kThreads, 0, st>>>(dr, dx, dxo, df, dfo, i_hist >= 0 ? dX + (size_t)i_hist * ldh : nullptr,
                                                      i_hist >= 0 ? dF + (size_t)i_hist * ldh : nullptr, m_inv, omega, anderson ? 0 : 1, n);
        ctx->stats.kernel_launches += 1;
        if (anderson) {
            int q = 0;
            for (int i = 0; i < m; i++)
                for (int j = i; j < m; j++, q++) { P.u[q] = dF + (size_t)i * ldh; P.v[q] = dF + (size_t)j * ldh; }
            for (int i = 0; i < m; i++, q++) { P.u[q] = dF + (size_t)i * ldh; P.v[q] = df; }
            dots_kernel<<<dim3(kBlocks, (unsigned)npair), kThreads, 0, st>>>(P, n, partials, tickets, outs);
            CHEFSI_CUDA(ctx, cudaMemcpyAsync(h.data(), outs, npair * sizeof(double), cudaMemcpyDeviceToHost, st));
            CHEFSI_CUDA(ctx, cudaStreamSynchronize(st));
            std::vector<double> FtF((size_t)m * m), Ftf(m);
            q = 0;
            for (int i = 0; i < m; i++)
                for (int j = i; j < m; j++, q++) FtF[(size_t)j * m + i] = FtF[(size_t)i * m + j] = h[q];
            for (int i = 0; i < m; i++, q++) Ftf[i] = h[q];
            Gammas G;
            for (int i = 0; i < kMaxM; i++) G.g[i] = 0.0;
            pinv_solve(m, FtF, Ftf, G.g);
            aar_anderson_kernel<<<kBlocks, kThreads, 0, st>>>(dx, dxo, df, dX, dF, G, m, beta, n, ldh);
            if (lap_residual_device(ctx, c, dx, db, dr)) return 1;
            P.u[0] = dr; P.v[0] = dr;
            dots_kernel<<<dim3(kBlocks, 1), kThreads, 0, st>>>(P, n, partials, tickets, outs);
            CHEFSI_CUDA(ctx, cudaMemcpyAsync(h.data(), outs, sizeof(double), cudaMemcpyDeviceToHost, st));
            CHEFSI_CUDA(ctx, cudaStreamSynchronize(st));
            r_2norm = sqrt(h[0]);
            ctx->stats.kernel_launches += 3;
        } else {
            if (lap_residual_device(ctx, c, dx, db, dr)) return 1;
        }
        iter++;
    }
    CHEFSI_CUDA(ctx, cudaMemcpyAsync(x, dx, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CHEFSI_CUDA(ctx, cudaStreamSynchronize(st));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return chefsi_fail(ctx, "aar kernels: %s", cudaGetErrorString(e));
    if (iterations) *iterations = iter;
    if (res_norm) *res_norm = r_2norm;
    return 0;
}
