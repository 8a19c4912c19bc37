// Shared device/host helpers for the MMBert sm_100a kernels.
// Everything in csrc/ is compiled with -gencode arch=compute_100a,code=sm_100a.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mmbert_sm100.h"

namespace mmb {

// ------------------------------------------------------------------ error plumbing (host)
void set_last_error(const char* fmt, ...);
int check_launch(const char* what, int kernels = 1);  // MMB_OK / MMB_ECUDA after kernel launch(es); counts them

#define MMB_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) {                                           \
            ::mmb::set_last_error(__VA_ARGS__);                  \
            return MMB_EINVAL;                                   \
        }                                                        \
    } while (0)

#define MMB_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            ::mmb::set_last_error("%s failed: %s", #call, cudaGetErrorString(e__)); \
            return MMB_ECUDA;                                                       \
        }                                                                           \
    } while (0)

int num_sms();

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// Exact (erf) GELU, the activation HF BERT uses ("gelu" -> F.gelu default).
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
// d/dx gelu(x) = Phi(x) + x * phi(x)
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// Fast erf for the GEMM epilogues: Abramowitz & Stegun 7.1.25, erf(z) = 1 - (a1 t + a2 t^2 + a3 t^3) exp(-z^2),
// t = 1 / (1 + 0.47047 z), |error| <= 2.5e-5 — two orders of magnitude below the bf16 rounding (2^-9 relative)
// of the outputs it feeds — in ~14 straight-line instructions (MUFU.RCP + MUFU.EX2 via the .approx.ftz PTX forms;
// erff and the IEEE __frcp_rn / exp2f forms compile to branchy code with subroutine calls that made the GELU
// epilogue, not the MMA, the bound of the FFN1 GEMM).  gelu_fast_grad shares the exponential: exp(-z^2) with
// z = x / sqrt(2) is also the Gaussian density.
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void erf_parts(float x, float& erf_abs, float& expmz2) {
    const float t = rcp_approx(fmaf(0.47047f * 0.70710678118654752f, fabsf(x), 1.0f));
    float poly = fmaf(0.7478556f, t, -0.0958798f);
    poly = fmaf(poly, t, 0.3480242f);
    expmz2 = ex2_approx(x * x * -0.72134752044448170f);   // exp(-x^2 / 2)
    erf_abs = fmaf(-poly * t, expmz2, 1.0f);
}
__device__ __forceinline__ float gelu_fast(float x) {
    float e, g;
    erf_parts(x, e, g);
    return 0.5f * x * (1.0f + copysignf(e, x));
}
__device__ __forceinline__ float gelu_fast_grad(float x) {
    float e, g;
    erf_parts(x, e, g);
    return fmaf(x * 0.3989422804014327f, g, 0.5f * (1.0f + copysignf(e, x)));
}

// Counter-based RNG for dropout: a 64-bit counter (stream, element index) mixed with the seed by a
// splitmix64-style finaliser. Forward and backward regenerate identical masks from (seed, stream, idx);
// nothing is stored.  Returns 32 uniform bits.
__device__ __forceinline__ uint32_t rng_bits(uint64_t seed, uint32_t stream, uint64_t idx) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1) + ((uint64_t)stream << 40) * 0xD1B54A32D192ED03ull;
    z ^= (uint64_t)stream * 0xA24BAED4963EE407ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 16);
}
// keep-decision for dropout probability p: keep iff bits >= p * 2^32
__device__ __forceinline__ bool rng_keep(uint64_t seed, uint32_t stream, uint64_t idx, uint32_t thresh) {
    return rng_bits(seed, stream, idx) >= thresh;
}
#endif  // __CUDACC__

inline uint32_t dropout_threshold(float p) {
    if (p <= 0.f) return 0u;
    double t = (double)p * 4294967296.0;
    if (t >= 4294967295.0) return 0xFFFFFFFFu;
    return (uint32_t)t;
}

}  // namespace mmb
