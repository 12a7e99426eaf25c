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
__device__ __forceinline__ void cp_async8(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

/* D(8x8) += A(8x4) B(4x8): a = A[lane/4][lane%4], b = B[lane%4][lane/4], c0,c1 = C[lane/4][2*(lane%4) + {0,1}] */
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int NWARP_M = 2, NWARP_N = 4; /* gemm_tn: 8 warps as 2 (m) x 4 (n) */

/* ---- C = A^T B --------------------------------------------------------------------------------------------- */
/* grid: (tiles_m * tiles_n, nslab); block 256 threads = 8 warps as 2 (m) x 4 (n).  FM x FN 8 x 8 fragments per warp:
 *   <4, 2, 32, 3>:  64 x  64 CTA tile, 32-deep chunks, 3 stages (110 KB, 2 CTAs per SM)   -- up to 64 columns
 *   <8, 4, 16, 4>: 128 x 128 CTA tile, 16-deep chunks, 4 stages (164 KB, 1 CTA per SM)    -- half the L2 -> SM bytes per
 *                  FMA and 32 independent DMMA chains per warp between two shared-memory fragment loads
 * PITCH = KCH + 4 doubles (4 mod 16): the 8 x 4 fragment reads of a warp hit 32 distinct banks. */
template <int FM, int FN, int KCH, int STG>
__global__ void __launch_bounds__(256)
gemm_tn_kernel(const double *__restrict__ A, size_t lda, const double *__restrict__ B, size_t ldb, int M, int N, size_t K,
               size_t kslab, int tiles_m, int upper, double *__restrict__ part /* [nslab][tiles][TM*TN] */)
{
    constexpr int TM = NWARP_M * 8 * FM, TN = NWARP_N * 8 * FN, PITCH = KCH + 4, QCH = KCH / 2 /* 16-byte chunks per column */;
    constexpr int COPIES = (TM + TN) * QCH / 256;
    static_assert((TM + TN) * QCH % 256 == 0, "copy loop");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sA = reinterpret_cast<double *>(smem_raw);            /* [STG][TM][PITCH]: column m, k contiguous */
    double *sB = sA + STG * TM * PITCH;                           /* [STG][TN][PITCH] */
    const int tile = blockIdx.x;
    int tm, tn;
    if (upper) { /* tiles of the upper triangle only, column by column: column tn holds tm = 0..tn */
        tn = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
        while ((tn + 1) * (tn + 2) / 2 <= tile) tn++;
        while (tn * (tn + 1) / 2 > tile) tn--;
        tm = tile - tn * (tn + 1) / 2;
    } else {
        tm = tile % tiles_m;
        tn = tile / tiles_m;
    }
    const int m0 = tm * TM, n0 = tn * TN;
    const size_t k_begin = (size_t)blockIdx.y * kslab;
    const size_t k_end = (k_begin + kslab < K) ? k_begin + kslab : K;
    const int nchunks = (k_end > k_begin) ? (int)((k_end - k_begin + KCH - 1) / KCH) : 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = (warp % NWARP_M) * 8 * FM, wn = (warp / NWARP_M) * 8 * FN;

    auto issue = [&](int c, int buf) {
        const size_t k0 = k_begin + (size_t)c * KCH;
#pragma unroll
        for (int t = 0; t < COPIES; t++) {
            const int idx = tid + t * 256;
            const int col = idx / QCH, q = idx % QCH;       /* column 0..TM+TN-1, 16-byte chunk */
            const bool isB = col >= TM;
            const int cc = isB ? col - TM : col;
            const int gcol = (isB ? n0 : m0) + cc;
            const size_t k = k0 + 2 * q;
            double *dst = (isB ? sB + ((size_t)buf * TN + cc) * PITCH : sA + ((size_t)buf * TM + cc) * PITCH) + 2 * q;
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

    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; i++)
#pragma unroll
        for (int j = 0; j < FN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STG - 1; s++) {
        if (s < nchunks) issue(s, s);
        cp_async_commit();
    }
    for (int c = 0; c < nchunks; c++) {
        const int buf = c % STG;
        cp_async_wait<STG - 2>();
        __syncthreads();
        if (c + STG - 1 < nchunks) issue(c + STG - 1, (c + STG - 1) % STG);
        cp_async_commit();
        const double *a_base = sA + ((size_t)buf * TM + wm + (lane >> 2)) * PITCH + (lane & 3);
        const double *b_base = sB + ((size_t)buf * TN + wn + (lane >> 2)) * PITCH + (lane & 3);
#pragma unroll
        for (int kk = 0; kk < KCH; kk += 4) {
            double af[FM], bf[FN];
#pragma unroll
            for (int i = 0; i < FM; i++) af[i] = a_base[(size_t)i * 8 * PITCH + kk];
#pragma unroll
            for (int j = 0; j < FN; j++) bf[j] = b_base[(size_t)j * 8 * PITCH + kk];
#pragma unroll
            for (int i = 0; i < FM; i++)
#pragma unroll
                for (int j = 0; j < FN; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();
    /* partial tile of this slab, column-major TM x TN */
    double *out = part + ((size_t)blockIdx.y * gridDim.x + tile) * (TM * TN);
#pragma unroll
    for (int i = 0; i < FM; i++)
#pragma unroll
        for (int j = 0; j < FN; j++) {
            const int r = wm + i * 8 + (lane >> 2), cc = wn + j * 8 + 2 * (lane & 3);
            out[(size_t)cc * TM + r] = acc[i][j][0];
            out[(size_t)(cc + 1) * TM + r] = acc[i][j][1];
        }
}

/* C[m + n ldc] = scale * sum over slabs (fixed order) of the partial T x T tiles.  sym != 0: the tiles are those of the
 * upper triangle; every element with m <= n is written to (m, n) and, times sym, to (n, m) (sym = +1: symmetric product,
 * -1: the antisymmetric imaginary part of a Hermitian one) */
__global__ void gemm_tn_reduce_kernel(const double *__restrict__ part, int nslab, int ntiles, int tiles_m, int T, int M, int N, double scale,
                                      double *__restrict__ C, size_t ldc, int cstride, int sym)
{
    const size_t tsz = (size_t)T * T, total = (size_t)ntiles * tsz;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int tile = (int)(e / tsz), r = (int)(e % T), c = (int)((e / T) % T);
        int tm, tn;
        if (sym) {
            tn = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
            while ((tn + 1) * (tn + 2) / 2 <= tile) tn++;
            while (tn * (tn + 1) / 2 > tile) tn--;
            tm = tile - tn * (tn + 1) / 2;
        } else {
            tm = tile % tiles_m;
            tn = tile / tiles_m;
        }
        const int m = tm * T + r, n = tn * T + c;
        if (m >= M || n >= N || (sym && m > n)) continue;
        double s = 0.0;
        for (int sl = 0; sl < nslab; sl++) s += part[(size_t)sl * total + e];
        C[((size_t)n * ldc + m) * cstride] = scale * s;
        if (sym && m < n) C[((size_t)m * ldc + n) * cstride] = sym * scale * s;
    }
}

/* ---- C = A Q ----------------------------------------------------------------------------------------------- */
constexpr int RT = 128;          /* rows per CTA tile */
constexpr int RPITCH = RT + 4;   /* 132 = 4 mod 16 */
/* 1-D grid, column tile fastest (the CTAs that share a row tile of A run together: A comes from DRAM once, then from
 * L2); block 256 = 8 warps as 4 (rows) x 2 (cols); warp tile 32 rows x 8 FNN columns:
 *   <4, 32, 3>: 128 x  64 CTA tile (up to 64 columns)      <8, 16, 4>: 128 x 128 CTA tile */
template <int FNN, int KCH, int STG>
__global__ void __launch_bounds__(256)
gemm_nn_kernel(const double *__restrict__ A, size_t lda, const double *__restrict__ Qm, size_t ldq, size_t K, int M, int N, int tiles_n,
               double *__restrict__ C, size_t ldc, int accumulate)
{
    constexpr int CT = 2 * 8 * FNN, PITCH = KCH + 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sA = reinterpret_cast<double *>(smem_raw);   /* [STG][KCH (m)][RPITCH]: column m of A, rows contiguous */
    double *sQ = sA + STG * KCH * RPITCH;                /* [STG][CT (n)][PITCH]: column n of Q, m contiguous */
    const size_t r0 = (size_t)(blockIdx.x / tiles_n) * RT;
    const int n0 = (int)(blockIdx.x % tiles_n) * CT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wr = (warp & 3) * 32, wn = (warp >> 2) * 8 * FNN;
    const int nchunks = (M + KCH - 1) / KCH;

    auto issue = [&](int c, int buf) {
        const int mc0 = c * KCH;
        /* A chunk: KCH columns x 128 rows = KCH x 64 16-byte chunks */
#pragma unroll
        for (int t = 0; t < KCH / 4; t++) {
            const int idx = tid + t * 256;
            const int mcol = idx >> 6, q = idx & 63;
            const size_t r = r0 + 2 * q;
            double *dst = sA + ((size_t)buf * KCH + mcol) * RPITCH + 2 * q;
            const int gm = mc0 + mcol;
            if (gm < M && r + 1 < K) cp_async16(dst, A + (size_t)gm * lda + r);
            else {
                dst[0] = (gm < M && r < K) ? A[(size_t)gm * lda + r] : 0.0;
                dst[1] = 0.0;
            }
        }
        /* Q chunk: CT columns (n) x KCH (m); a column of Q is only 8-byte aligned in general: 8-byte copies */
#pragma unroll
        for (int t = 0; t < CT * KCH / 256; t++) {
            const int idx = tid + t * 256;
            const int ncol = idx / KCH, mm = idx % KCH;
            const int gn = n0 + ncol, gm = mc0 + mm;
            double *dst = sQ + ((size_t)buf * CT + ncol) * PITCH + mm;
            if (gn < N && gm < M) cp_async8(dst, Qm + (size_t)gn * ldq + gm);
            else *dst = 0.0;
        }
    };

    double acc[4][FNN][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < FNN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STG - 1; s++) {
        if (s < nchunks) issue(s, s);
        cp_async_commit();
    }
    for (int c = 0; c < nchunks; c++) {
        const int buf = c % STG;
        cp_async_wait<STG - 2>();
        __syncthreads();
        if (c + STG - 1 < nchunks) issue(c + STG - 1, (c + STG - 1) % STG);
        cp_async_commit();
        /* a = A[row = lane/4][m = lane%4] -> sA[m][row]; b = Q[m = lane%4][n = lane/4] -> sQ[n][m] */
        const double *a_base = sA + ((size_t)buf * KCH + (lane & 3)) * RPITCH + wr + (lane >> 2);
        const double *b_base = sQ + ((size_t)buf * CT + wn + (lane >> 2)) * PITCH + (lane & 3);
#pragma unroll
        for (int kk = 0; kk < KCH; kk += 4) {
            double af[4], bf[FNN];
#pragma unroll
            for (int i = 0; i < 4; i++) af[i] = a_base[(size_t)kk * RPITCH + i * 8];
#pragma unroll
            for (int j = 0; j < FNN; j++) bf[j] = b_base[(size_t)j * 8 * PITCH + kk];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < FNN; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < FNN; j++) {
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

template <int FM, int FN, int KCH, int STG>
static int run_gemm_tn(chefsi_ctx *ctx, const double *A, size_t lda, const double *B, size_t ldb, int M, int N, size_t K, double scale,
                       double *C, size_t ldc, int cstride, int sym, int ctas_per_sm)
{
    constexpr int T = NWARP_M * 8 * FM;
    static_assert(T == NWARP_N * 8 * FN, "square CTA tiles");
    if (sym && M != N) { chefsi_fail(ctx, "gemm_tn: a symmetric product needs M == N"); return -1; }
    const int tiles_m = (M + T - 1) / T, tiles_n = (N + T - 1) / T, ntiles = sym ? tiles_m * (tiles_m + 1) / 2 : tiles_m * tiles_n;
    /* split K into slabs over CTAs (tall-skinny operands give a few tiles only): at least two waves of CTAs, and among
       the slab counts of the next wave the one that fills whole waves best -- a tail wave of a few CTAs costs a full
       slab time (Ns = 512, 128 x 128 tiles: 16 tiles x 19 slabs = 2.05 waves would run at 68 %) */
    const size_t wave = (size_t)ctx->num_sms * ctas_per_sm;
    const size_t max_slab = std::max<size_t>(1, (K + 8 * KCH - 1) / (8 * KCH)); /* at least 8 chunks per slab */
    const size_t lo = std::max<size_t>(1, (2 * wave + ntiles - 1) / ntiles), hi = lo + (wave + ntiles - 1) / ntiles + 1;
    size_t nslab = std::min(lo, max_slab);
    double best = -1.0;
    for (size_t cand = lo; cand <= hi && cand <= max_slab; cand++) {
        const size_t ctas = cand * ntiles, waves = (ctas + wave - 1) / wave;
        const double eff = (double)ctas / (double)(waves * wave);
        if (eff > best + 1e-9) { best = eff; nslab = cand; }
    }
    if (nslab > 65535) nslab = 65535;
    size_t kslab = (K + nslab - 1) / nslab;
    kslab = (kslab + KCH - 1) / KCH * KCH;                 /* slabs start on even element offsets */
    nslab = (K + kslab - 1) / kslab;
    if (ensure_bytes(ctx, &ctx->d_gemm_ws, &ctx->gemm_ws_bytes, nslab * ntiles * T * T * sizeof(double))) return -1;
    const size_t smem = (size_t)2 * STG * T * (KCH + 4) * sizeof(double);
    auto kern = gemm_tn_kernel<FM, FN, KCH, STG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { chefsi_fail(ctx, "cudaFuncSetAttribute(gemm_tn): %s", cudaGetErrorString(e)); return -1; }
    kern<<<dim3((unsigned)ntiles, (unsigned)nslab), 256, smem, ctx->stream>>>(A, lda, B, ldb, M, N, K, kslab, tiles_m, sym != 0,
                                                                              (double *)ctx->d_gemm_ws);
    gemm_tn_reduce_kernel<<<std::min(4 * ctx->num_sms, (ntiles * T * T + 255) / 256), 256, 0, ctx->stream>>>(
        (const double *)ctx->d_gemm_ws, (int)nslab, ntiles, tiles_m, T, M, N, scale, C, ldc, cstride, sym);
    e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "gemm_tn launch: %s", cudaGetErrorString(e)); return -1; }
    return 2;
}

/* C(M x N, ld ldc, device) = scale * A^T B with A: K x M (lda), B: K x N (ldb), all device, column-major.
   sym = +1 / -1: the caller knows the product is symmetric / antisymmetric (Y^T Y, Y^T H Y and the imaginary parts of their
   complex forms): only the tiles of the upper triangle are computed and mirrored, about half the work.  What consumes
   these matrices (dsygvd / zhegvd with uplo = 'U', eigenSolver.c:1318) reads the upper triangle only. */
int launch_gemm_tn(chefsi_ctx *ctx, const double *A, size_t lda, const double *B, size_t ldb, int M, int N, size_t K, double scale,
                   double *C, size_t ldc, int cstride, int sym)
{
    if (!ctx->gemm_symmetric) sym = 0;
    if (std::max(M, N) > 64 && ctx->gemm_big_tiles)
        return run_gemm_tn<8, 4, 16, 4>(ctx, A, lda, B, ldb, M, N, K, scale, C, ldc, cstride, sym, 1);
    return run_gemm_tn<4, 2, 32, 3>(ctx, A, lda, B, ldb, M, N, K, scale, C, ldc, cstride, sym, 2);
}

template <int FNN, int KCH, int STG>
static int run_gemm_nn(chefsi_ctx *ctx, const double *A, size_t lda, const double *Q, size_t ldq, size_t K, int M, int N, double *C,
                       size_t ldc, int accumulate)
{
    constexpr int CT = 2 * 8 * FNN;
    const size_t smem = ((size_t)STG * KCH * RPITCH + (size_t)STG * CT * (KCH + 4)) * sizeof(double);
    auto kern = gemm_nn_kernel<FNN, KCH, STG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { chefsi_fail(ctx, "cudaFuncSetAttribute(gemm_nn): %s", cudaGetErrorString(e)); return -1; }
    const size_t rt = (K + RT - 1) / RT, tiles_n = (size_t)(N + CT - 1) / CT;
    if (rt * tiles_n > 0x7fffffffULL) { chefsi_fail(ctx, "gemm_nn: too many tiles"); return -1; }
    kern<<<(unsigned)(rt * tiles_n), 256, smem, ctx->stream>>>(A, lda, Q, ldq, K, M, N, (int)tiles_n, C, ldc, accumulate);
    e = cudaGetLastError();
    if (e != cudaSuccess) { chefsi_fail(ctx, "gemm_nn launch: %s", cudaGetErrorString(e)); return -1; }
    return 1;
}

/* C(K x N, ldc) = A Q with A: K x M (lda), Q: M x N (ldq); C must not alias A */
int launch_gemm_nn(chefsi_ctx *ctx, const double *A, size_t lda, const double *Q, size_t ldq, size_t K, int M, int N, double *C,
                   size_t ldc, int accumulate)
{
    if (N > 64 && ctx->gemm_big_tiles) return run_gemm_nn<8, 16, 4>(ctx, A, lda, Q, ldq, K, M, N, C, ldc, accumulate);
    return run_gemm_nn<4, 32, 3>(ctx, A, lda, Q, ldq, K, M, N, C, ldc, accumulate);
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
