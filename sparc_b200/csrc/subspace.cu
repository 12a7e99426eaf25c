/*
 * subspace.cu -- Rayleigh-Ritz projection and subspace rotation with the filtered block resident on the device
 * (SURVEY.md 8f-1), real (Gamma-point) data, single-device contexts.
 *
 *   project   HY = H Y (one more Hamiltonian apply on the kernels of the filter),  Mp = Y^T Y,  Hp = Y^T (H Y)
 *             -- DP_Project_Hamiltonian (src/eigenSolver.c:939-1086; Project_Hamiltonian :1477-1669)
 *   rotate    X = Y Q            -- DP_Subspace_Rotation (src/eigenSolver.c:1386-1443; Subspace_Rotation :1854-1918)
 *
 * The reference does these with cblas_dgemm on the host (its accelerator hook names the same three GEMMs:
 * ACCEL_DGEMM at eigenSolver.c:1006-1017,1405).  Here they are FP64 tensor-core products (mma.sync.m8n8k4.f64, the
 * only FP64 MMA on sm_100a -- tcgen05 has no FP64 kind): the one place on this path where the work is a dense
 * contraction.  Only the Ns x Ns matrices Hp, Mp, Q cross PCIe; Y never leaves the device between the filter and the
 * rotation, and the rotated block is what goes back to the caller.
 *
 *   gemm_tn  C(M x N) = A^T B, A: K x M, B: K x N column-major, K = Nd (10^4 .. 10^6), M = N = Ns.  CTA tile 64 x 64,
 *            K cut into slabs over CTAs (split-K: tall-skinny operands would otherwise fill only a few SMs); every
 *            slab writes its partial tile, a second kernel adds the slabs in a fixed order (deterministic).
 *   gemm_nn  C(K x N) = A Q,  A: K x M, Q: M x N.  CTA tile 128 rows x 64 columns, loop over M.
 * Complex (k-point) data: a complex column is a real column of twice the length (re, im interleaved), so
 *   Re(A^H B) = A_view^T B_view,  Im(A^H B) = A_view^T (-i B)_view,   (A Q)_view = A_view Q_r + (i A)_view Q_i
 * -- two real DMMA products each, plus one rotation pass (rot90_kernel) that forms -i B or i A.
 * Both stage 32-deep operand chunks with 16-byte cp.async into a 3-stage shared-memory ring; fragments are read
 * with bank-conflict-free pitches (pitch = 4 mod 16 doubles).
 */
#include <algorithm>
#include <cstdio>

#include "chefsi_internal.h"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

/* D(8x8) += A(8x4) B(4x8): a = A[lane/4][lane%4], b = B[lane%4][lane/4], c0,c1 = C[lane/4][2*(lane%4) + {0,1}] */
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int KC = 32;         /* depth of a staged operand chunk */
constexpr int PITCH = KC + 4;  /* 36 doubles: 4 mod 16 -> the 8 x 4 fragment reads of a half warp hit 16 distinct banks pairs */
constexpr int NSTG = 3;

/* ---- C = A^T B --------------------------------------------------------------------------------------------- */
/* grid: (tiles_m * tiles_n, nslab); block 256 threads = 8 warps as 2 (m) x 4 (n): warp tile 32 x 16 */
__global__ void __launch_bounds__(256)
gemm_tn_kernel(const double *__restrict__ A, size_t lda, const double *__restrict__ B, size_t ldb, int M, int N, size_t K,
               size_t kslab, int tiles_m, double *__restrict__ part /* [nslab][tiles][64*64] */)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sA = reinterpret_cast<double *>(smem_raw);            /* [NSTG][64][PITCH]: column m, k contiguous */
    double *sB = sA + NSTG * 64 * PITCH;
    const int tile = blockIdx.x, tm = tile % tiles_m, tn = tile / tiles_m;
    const int m0 = tm * 64, n0 = tn * 64;
    const size_t k_begin = (size_t)blockIdx.y * kslab;
    const size_t k_end = (k_begin + kslab < K) ? k_begin + kslab : K;
    const int nchunks = (k_end > k_begin) ? (int)((k_end - k_begin + KC - 1) / KC) : 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = (warp & 1) * 32, wn = (warp >> 1) * 16;

    auto issue = [&](int c, int buf) {
        const size_t k0 = k_begin + (size_t)c * KC;
        /* 128 columns (64 of A, 64 of B) x 16 chunks of 16 bytes = 2048 copies / 256 threads */
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const int idx = tid + t * 256;
            const int col = idx >> 4, q = idx & 15;        /* column 0..127, 16-byte chunk 0..15 */
            const bool isB = col >= 64;
            const int cc = isB ? col - 64 : col;
            const int gcol = (isB ? n0 : m0) + cc;
            const size_t k = k0 + 2 * q;
            double *dst = (isB ? sB : sA) + ((size_t)buf * 64 + cc) * PITCH + 2 * q;
            const bool colok = gcol < (isB ? N : M);
            if (colok && k + 1 < k_end) {
                cp_async16(dst, (isB ? B + (size_t)gcol * ldb : A + (size_t)gcol * lda) + k);
            } else {
                const double *src = isB ? B + (size_t)gcol * ldb : A + (size_t)gcol * lda;
                dst[0] = (colok && k < k_end) ? src[k] : 0.0;
                dst[1] = 0.0;
            }
        }
    };

    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < NSTG - 1; s++) {
        if (s < nchunks) issue(s, s);
        cp_async_commit();
    }
    for (int c = 0; c < nchunks; c++) {
        const int buf = c % NSTG;
        cp_async_wait<NSTG - 2>();
        __syncthreads();
        if (c + NSTG - 1 < nchunks) issue(c + NSTG - 1, (c + NSTG - 1) % NSTG);
        cp_async_commit();
        const double *a_base = sA + ((size_t)buf * 64 + wm + (lane >> 2)) * PITCH + (lane & 3);
        const double *b_base = sB + ((size_t)buf * 64 + wn + (lane >> 2)) * PITCH + (lane & 3);
#pragma unroll
        for (int kk = 0; kk < KC; kk += 4) {
            double af[4], bf[2];
#pragma unroll
            for (int i = 0; i < 4; i++) af[i] = a_base[(size_t)i * 8 * PITCH + kk];
#pragma unroll
            for (int j = 0; j < 2; j++) bf[j] = b_base[(size_t)j * 8 * PITCH + kk];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 2; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();
    /* partial tile of this slab, column-major 64 x 64 */
    double *out = part + ((size_t)blockIdx.y * gridDim.x + tile) * 4096;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int r = wm + i * 8 + (lane >> 2), cc = wn + j * 8 + 2 * (lane & 3);
            out[(size_t)cc * 64 + r] = acc[i][j][0];
            out[(size_t)(cc + 1) * 64 + r] = acc[i][j][1];
        }
}

/* C[m + n ldc] = scale * sum over slabs (fixed order) of the partial tiles */
__global__ void gemm_tn_reduce_kernel(const double *__restrict__ part, int nslab, int ntiles, int tiles_m, int M, int N, double scale,
                                      double *__restrict__ C, size_t ldc, int cstride)
{
    const size_t total = (size_t)ntiles * 4096;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int tile = (int)(e / 4096), r = (int)(e % 64), c = (int)((e / 64) % 64);
        const int m = (tile % tiles_m) * 64 + r, n = (tile / tiles_m) * 64 + c;
        if (m >= M || n >= N) continue;
        double s = 0.0;
        for (int sl = 0; sl < nslab; sl++) s += part[(size_t)sl * total + e];
        C[((size_t)n * ldc + m) * cstride] = scale * s;
    }
}

/* ---- C = A Q ----------------------------------------------------------------------------------------------- */
constexpr int RT = 128;          /* rows per CTA tile */
constexpr int RPITCH = RT + 4;   /* 132 = 4 mod 16 */
/* grid: (row tiles, column tiles of 64); block 256 = 8 warps as 4 (rows) x 2 (cols): warp tile 32 rows x 32 columns */
__global__ void __launch_bounds__(256)
gemm_nn_kernel(const double *__restrict__ A, size_t lda, const double *__restrict__ Qm, size_t ldq, size_t K, int M, int N,
               double *__restrict__ C, size_t ldc, int accumulate)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sA = reinterpret_cast<double *>(smem_raw);   /* [NSTG][KC (m)][RPITCH]: column m of A, rows contiguous */
    double *sQ = sA + NSTG * KC * RPITCH;                /* [NSTG][64 (n)][PITCH]: column n of Q, m contiguous */
    const size_t r0 = (size_t)blockIdx.x * RT;
    const int n0 = blockIdx.y * 64;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wr = (warp & 3) * 32, wn = (warp >> 2) * 32;
    const int nchunks = (M + KC - 1) / KC;

    auto issue = [&](int c, int buf) {
        const int mc0 = c * KC;
        /* A chunk: 32 columns x 128 rows = 32 x 64 16-byte chunks = 2048 copies */
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const int idx = tid + t * 256;
            const int mcol = idx >> 6, q = idx & 63;
            const size_t r = r0 + 2 * q;
            double *dst = sA + ((size_t)buf * KC + mcol) * RPITCH + 2 * q;
            const int gm = mc0 + mcol;
            if (gm < M && r + 1 < K) cp_async16(dst, A + (size_t)gm * lda + r);
            else {
                dst[0] = (gm < M && r < K) ? A[(size_t)gm * lda + r] : 0.0;
                dst[1] = 0.0;
            }
        }
        /* Q chunk: 64 columns (n) x 32 (m): 64 x 16 chunks = 1024 copies; element alignment of Q is only 8 bytes in general */
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const int idx = tid + t * 256;
            const int ncol = idx >> 5, mm = idx & 31;
            const int gn = n0 + ncol, gm = mc0 + mm;
            sQ[((size_t)buf * 64 + ncol) * PITCH + mm] = (gn < N && gm < M) ? Qm[(size_t)gn * ldq + gm] : 0.0;
        }
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < NSTG - 1; s++) {
        if (s < nchunks) issue(s, s);
        cp_async_commit();
    }
    for (int c = 0; c < nchunks; c++) {
        const int buf = c % NSTG;
        cp_async_wait<NSTG - 2>();
        __syncthreads();
        if (c + NSTG - 1 < nchunks) issue(c + NSTG - 1, (c + NSTG - 1) % NSTG);
        cp_async_commit();
        /* a = A[row = lane/4][m = lane%4] -> sA[m][row]; b = Q[m = lane%4][n = lane/4] -> sQ[n][m] */
        const double *a_base = sA + ((size_t)buf * KC + (lane & 3)) * RPITCH + wr + (lane >> 2);
        const double *b_base = sQ + ((size_t)buf * 64 + wn + (lane >> 2)) * PITCH + (lane & 3);
#pragma unroll
        for (int kk = 0; kk < KC; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; i++) af[i] = a_base[(size_t)kk * RPITCH + i * 8];
#pragma unroll
            for (int j = 0; j < 4; j++) bf[j] = b_base[(size_t)j * 8 * PITCH + kk];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const size_t r = r0 + wr + i * 8 + (lane >> 2);
            const int cc = n0 + wn + j * 8 + 2 * (lane & 3);
            if (r < K) {
                if (cc < N) C[(size_t)cc * ldc + r] = acc[i][j][0] + (accumulate ? C[(size_t)cc * ldc + r] : 0.0);
                if (cc + 1 < N) C[(size_t)(cc + 1) * ldc + r] = acc[i][j][1] + (accumulate ? C[(size_t)(cc + 1) * ldc + r] : 0.0);
            }
        }
}

/* out = s * i * in on interleaved complex columns: (re, im) -> s * (-im, re).  s = +1: i z, s = -1: -i z */
__global__ void rot90_kernel(const double2 *__restrict__ in, double2 *__restrict__ out, size_t n, size_t ld, double s)
{
    const double2 *ci = in + (size_t)blockIdx.y * ld;
    double2 *co = out + (size_t)blockIdx.y * ld;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double2 z = ci[i];
        co[i] = make_double2(-s * z.y, s * z.x);
    }
}

/* Q (complex, M x N, interleaved, ld ldq) -> Qr, Qi (real M x N, ld M) */
__global__ void split_complex_kernel(const double2 *__restrict__ Q, size_t ldq, int M, int N, double *__restrict__ Qr, double *__restrict__ Qi)
{
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)M * N; e += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(e % M), n = (int)(e / M);
        const double2 z = Q[(size_t)n * ldq + m];
        Qr[e] = z.x;
        Qi[e] = z.y;
    }
}

int ensure_bytes(chefsi_ctx *ctx, void **p, size_t *have, size_t need)
{
    if (need <= *have) return 0;
    cudaFree(*p);
    *p = nullptr;
    *have = 0;
    cudaError_t e = cudaMalloc(p, need);
    if (e != cudaSuccess) { cudaGetLastError(); chefsi_fail(ctx, "subspace: cudaMalloc(%zu): %s", need, cudaGetErrorString(e)); return 1; }
    *have = need;
    return 0;
}

}  // namespace

/* C(M x N, ld ldc, device) = scale * A^T B with A: K x M (lda), B: K x N (ldb), all device, column-major */
int launch_gemm_tn(chefsi_ctx *ctx, const double *A, size_t lda, const double *B, size_t ldb, int M, int N, size_t K, double scale,
                   double *C, size_t ldc, int cstride)
{
    const int tiles_m = (M + 63) / 64, tiles_n = (N + 63) / 64, ntiles = tiles_m * tiles_n;
    /* slabs: enough CTAs to fill the GPU twice, at least 8 chunks per slab */
    size_t nslab = (size_t)std::max(1, (2 * ctx->num_sms + ntiles - 1) / ntiles);
    const size_t max_slab = (K + 8 * KC - 1) / (8 * KC);
    if (nslab > max_slab) nslab = max_slab ? max_slab : 1;
    if (nslab > 65535) nslab = 65535;
    size_t kslab = (K + nslab - 1) / nslab;
    kslab = (kslab + KC - 1) / KC * KC;                 /* slabs start on even element offsets */
    nslab = (K + kslab - 1) / kslab;
    if (ensure_bytes(ctx, &ctx->d_gemm_ws, &ctx->gemm_ws_bytes, nslab * ntiles * 4096 * sizeof(double))) return -1;
    const size_t smem = (size_t)2 * NSTG * 64 * PITCH * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { chefsi_fail(ctx, "cudaFuncSetAttribute(gemm_tn): %s", cudaGetErrorString(e)); return -1; }
    gemm_tn_kernel<<<dim3((unsigned)ntiles, (unsigned)nslab), 256, smem, ctx->stream>>>(A, lda, B, ldb, M, N, K, kslab, tiles_m,
                                                                                          (double *)ctx->d_gemm_ws);
    gemm_tn_reduce_kernel<<<std::min(4 * ctx->num_sms, (ntiles * 4096 + 255) / 256), 256, 0, ctx->stream>>>(
        (const double *)ctx->d_gemm_ws, (int)nslab, ntiles, tiles_m, M, N, scale, C, ldc, cstride);
    e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "gemm_tn launch: %s", cudaGetErrorString(e)); return -1; }
    return 2;
}

/* C(K x N, ldc) = A Q with A: K x M (lda), Q: M x N (ldq); C must not alias A */
int launch_gemm_nn(chefsi_ctx *ctx, const double *A, size_t lda, const double *Q, size_t ldq, size_t K, int M, int N, double *C,
                   size_t ldc, int accumulate)
{
    const size_t smem = ((size_t)NSTG * KC * RPITCH + (size_t)NSTG * 64 * PITCH) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(gemm_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { chefsi_fail(ctx, "cudaFuncSetAttribute(gemm_nn): %s", cudaGetErrorString(e)); return -1; }
    const size_t rt = (K + RT - 1) / RT;
    if (rt > 0x7fffffffULL) { chefsi_fail(ctx, "gemm_nn: too many row tiles"); return -1; }
    gemm_nn_kernel<<<dim3((unsigned)rt, (unsigned)((N + 63) / 64)), 256, smem, ctx->stream>>>(A, lda, Q, ldq, K, M, N, C, ldc, accumulate);
    e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "gemm_nn launch: %s", cudaGetErrorString(e)); return -1; }
    return 1;
}

/* out = s * i * in for ncol interleaved complex columns of n elements (column stride ld complex elements) */
int launch_rot90(chefsi_ctx *ctx, const void *in, void *out, size_t n, size_t ld, int ncol, double s)
{
    if (ncol <= 0) return 0;
    rot90_kernel<<<dim3(64, (unsigned)ncol), 256, 0, ctx->stream>>>((const double2 *)in, (double2 *)out, n, ld, s);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "rot90 launch: %s", cudaGetErrorString(e)); return -1; }
    return 1;
}

int launch_split_complex(chefsi_ctx *ctx, const void *Q, size_t ldq, int M, int N, double *Qr, double *Qi)
{
    split_complex_kernel<<<64, 256, 0, ctx->stream>>>((const double2 *)Q, ldq, M, N, Qr, Qi);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "split launch: %s", cudaGetErrorString(e)); return -1; }
    return 1;
}
