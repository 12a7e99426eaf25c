/*
 * tma_ring.cuh -- the PTX building blocks shared by the TMA-streamed stencil kernels (stencil_stream_dense.cu,
 * stencil_stream_kpt.cu): shared-memory mbarriers, the tiled TMA load that completes on
 * one, vector stores, and the driver entry point that encodes tensor maps.  sm_100a.
 */
#ifndef CHEFSI_TMA_RING_CUH
#define CHEFSI_TMA_RING_CUH

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tma_ring {

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
/* try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint expires)
   instead of re-issuing try_wait + branch every few hundred cycles.  Without it the producer lane's spinning on
   the `empty` barriers was 18 % of all instructions the streaming kernel executed (ncu source view, profiles/)
   and competed with the two consumer warps of its scheduler for issue slots. */
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
}
/* TMA tiled load of one 4-D box (x, y, z, column), completion counted in bytes on an mbarrier */
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void stg256(double *p, const double (&v)[4])
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
__device__ __forceinline__ void stg128(double *p, double v0, double v1)
{
    asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v0), "d"(v1) : "memory");
}
/* origin of tile t along an axis of N points tiled by T: the last tile is shifted inwards */
__device__ __forceinline__ int tile_origin(int t, int T, int N) { return min(t * T, N - T); }

/* cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda) */
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

}  // namespace tma_ring
#endif
