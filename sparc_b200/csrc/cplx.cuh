/*
 * cplx.cuh -- the handful of arithmetic helpers that let one kernel body serve both the
 * Gamma-point (double) and k-point (double2 = re,im; layout of C99 double _Complex) variants.
 * All stencil weights are real, so "complex" work is two independent real lanes except for
 * the Bloch-phase multiply.
 */
#ifndef CHEFSI_CPLX_CUH
#define CHEFSI_CPLX_CUH

#include <cuda_runtime.h>

namespace cplx {

template <typename T> struct is_complex { static constexpr bool value = false; };
template <> struct is_complex<double2> { static constexpr bool value = true; };

template <typename T> __host__ __device__ __forceinline__ T zero();
template <> __host__ __device__ __forceinline__ double zero<double>() { return 0.0; }
template <> __host__ __device__ __forceinline__ double2 zero<double2>() { return make_double2(0.0, 0.0); }

__device__ __forceinline__ double add(double a, double b) { return a + b; }
__device__ __forceinline__ double sub(double a, double b) { return a - b; }
__device__ __forceinline__ double mul(double a, double w) { return a * w; }
__device__ __forceinline__ double fma(double a, double w, double acc) { return ::fma(a, w, acc); }
__device__ __forceinline__ double mul_phase(double a, double, double) { return a; }

__device__ __forceinline__ double2 add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 mul(double2 a, double w) { return make_double2(a.x * w, a.y * w); }
__device__ __forceinline__ double2 fma(double2 a, double w, double2 acc)
{
    return make_double2(::fma(a.x, w, acc.x), ::fma(a.y, w, acc.y));
}
/* a * (re + i im) */
__device__ __forceinline__ double2 mul_phase(double2 a, double re, double im)
{
    return make_double2(a.x * re - a.y * im, a.x * im + a.y * re);
}

}  // namespace cplx
#endif
