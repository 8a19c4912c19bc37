// Shared device/host helpers for the MMBert sm_100a kernels.
// Everything in csrc/ is compiled with -gencode arch=compute_100a,code=sm_100a.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

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
int persistent_sms();   // num_sms() minus the SMs reserved for concurrent collectives (mmb_set_reserved_sms)

// Opt-in to more than 48 KB of dynamic shared memory.  The attribute is per device and must be set before the first
// launch there: one flag word per call site, bit = device ordinal (mod 64).  Two threads racing only repeat the call.
#define MMB_ENSURE_SMEM(bytes, ...)                                                                                  \
    do {                                                                                                             \
        static std::atomic<unsigned long long> done__{0};                                                            \
        int dev__ = 0;                                                                                               \
        MMB_CUDA(cudaGetDevice(&dev__));                                                                             \
        const unsigned long long bit__ = 1ull << (dev__ & 63);                                                       \
        if (!(done__.load(std::memory_order_acquire) & bit__)) {                                                     \
            MMB_CUDA(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));  \
            done__.fetch_or(bit__, std::memory_order_release);                                                       \
        }                                                                                                            \
    } while (0)

// ------------------------------------------------------------------ programmatic dependent launch
// A training step is ~345 kernels in one stream, most of them persistent grids that fill the chip; between two of them the
// chip drains, the next grid is scheduled and runs its prologue (barrier init, TMEM allocation, descriptor prefetch) before
// any useful work starts.  Launched with cudaLaunchAttributeProgrammaticStreamSerialization, kernel N+1's CTAs are placed as
// soon as every CTA of kernel N has executed griddepcontrol.launch_dependents (pdl_trigger(), first statement of every
// converted kernel) and an SM has room; they run their prologue and then block in griddepcontrol.wait (pdl_wait()) until
// kernel N has COMPLETED and its writes are visible.  Rules every converted kernel keeps:
//   * every thread executes pdl_wait() before its first global-memory access (read OR write) and before any return — a
//     grid that finished without waiting would let ITS dependent overtake the grid before it;
//   * only shared-memory / TMEM / barrier set-up and reads of kernel parameters happen before it.
// Launched without the attribute both instructions are no-ops.  MMB_PDL=1 turns the attribute on (=2: also under stream
// capture).  DEFAULT OFF: measured on B200 (same-box A/B) it buys nothing here — every heavy kernel is a
// persistent grid whose CTAs each take a whole SM (~200 KB of shared memory, the full register file), so a dependent CTA
// can only be placed once its predecessor's CTA on that SM has exited, and the grid still starts when the slowest one ends.
int pdl_mode();      // 0 = off (default), 1 = on outside stream capture, 2 = always
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    int on = pdl_mode();
    if (on == 1) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) on = 0;
    }
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = on ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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

// Counter-based RNG for dropout.  Forward and backward regenerate identical masks from (seed, stream, row, column):
// nothing is stored.  32-bit murmur3-style mixing only (no 64-bit multiplies): a per-row key is hashed once from
// (seed, stream, row id), then ONE hash per PAIR of adjacent columns yields two 16-bit uniform values, i.e. about
// 4 integer instructions per dropout decision.  keep iff the 16-bit value >= thresh16 = round(p * 65536).
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}
__device__ __forceinline__ uint32_t rng_row_key(uint64_t seed, uint32_t stream, uint32_t row_id) {
    uint32_t h = (uint32_t)seed ^ (stream * 0x9E3779B9u);
    uint32_t k = row_id * 0xcc9e2d51u;
    k = rotl32(k, 15) * 0x1b873593u;
    h = rotl32(h ^ k, 13) * 5u + 0xe6546b64u;
    k = (uint32_t)(seed >> 32) * 0xcc9e2d51u;
    k = rotl32(k, 15) * 0x1b873593u;
    h = rotl32(h ^ k, 13) * 5u + 0xe6546b64u;
    return fmix32(h ^ 12u);
}
// 32 random bits for column pair `pair` (columns 2*pair and 2*pair+1) of the row: low 16 bits -> even column
__device__ __forceinline__ uint32_t rng_pair(uint32_t row_key, uint32_t pair) {
    uint32_t k = pair * 0xcc9e2d51u;
    k = rotl32(k, 15) * 0x1b873593u;
    return fmix32(row_key ^ k);
}
__device__ __forceinline__ bool rng_keep_lo(uint32_t bits, uint32_t thresh16) { return (bits & 0xFFFFu) >= thresh16; }
__device__ __forceinline__ bool rng_keep_hi(uint32_t bits, uint32_t thresh16) { return (bits >> 16) >= thresh16; }

// Attention-probability dropout (tcgen05 kernels): the decision for (query q, key k) is
//     keep  <=>  (qkey[q] * kkey[k]) mod 2^32  >=  thresh16 << 16
// with two ODD 32-bit keys hashed once per query row / key column of a (sequence, head).  One IMAD + one compare
// per probability, and — unlike a per-row stream — equally cheap whether a thread walks along the keys of one
// query (forward, dQ) or along the queries of one key (dK/dV).  For a fixed odd qkey the map kkey -> product is a
// bijection of the odd residues, so every row (and every column) of the mask is an independent uniform draw.
__device__ __forceinline__ uint32_t attn_drop_qkey(uint64_t seed, uint32_t stream, uint32_t prob_row) {
    return rng_row_key(seed, stream, prob_row) | 1u;
}
__device__ __forceinline__ uint32_t attn_drop_kkey(uint64_t seed, uint32_t stream, uint32_t prob_row) {
    return rng_row_key(seed ^ 0x9E3779B97F4A7C15ull, stream ^ 0x5bd1e995u, prob_row) | 1u;
}
__device__ __forceinline__ bool attn_keep(uint32_t qkey, uint32_t kkey, uint32_t thresh32) { return qkey * kkey >= thresh32; }

// packed fp32 x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2 — half the issue slots of the scalar forms)
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
// Packed (two values per instruction) erf-GELU and its derivative for the GEMM epilogues, which are bound by issue slots and
// by the MUFU pipe, not by the MMA.  ONE exponential per value: with ax = |x|,
//     erfc(ax / sqrt 2) = 2 Phi(-ax) ~= 2^q(ax),   q = ax (c1 + ax (c2 + ax (c3 + ax (c4 + ax c5))))
// (weighted least-squares fit on [0, 8]; c5 < 0, so q -> -inf and 2^q -> 0 beyond it), hence
//     Phi(x) = 1/2 + copysign(1/2 - 2^q / 2, x)            gelu(x)  = x Phi(x)
//     phi(x) = Phi'(x) = -(ln 2 / 2) q'(ax) 2^q            gelu'(x) = Phi(x) + x phi(x)
// — the density comes out of the SAME exponential through the polynomial's derivative, so the reciprocal of the
// Abramowitz-Stegun form (a second MUFU op per value) is gone.  Max abs error against erf-GELU in fp32: 4.4e-6 (gelu),
// 1.0e-5 (gelu'), both far below the bf16 rounding of the outputs.  -DMMB_GELU_AS restores the two-MUFU form (A/B runs).
#ifndef MMB_GELU_AS
__device__ __forceinline__ void gelu_fast2(float2 x, float2& g, float2& dg) {
    constexpr float c1 = -1.151042103767395f, c2 = -0.45956283807754517f, c3 = -0.05201205611228943f,
                    c4 = 0.0070396410301327705f, c5 = -0.00044516444904729724f;
    constexpr float h = -0.34657359027997264f;      // -ln 2 / 2
    constexpr float d0 = h * c1, d1 = 2.f * h * c2, d2 = 3.f * h * c3, d3 = 4.f * h * c4, d4 = 5.f * h * c5;
    const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
    float2 q = fma2(ax, make_float2(c5, c5), make_float2(c4, c4));
    q = fma2(q, ax, make_float2(c3, c3));
    q = fma2(q, ax, make_float2(c2, c2));
    q = fma2(q, ax, make_float2(c1, c1));
    q = mul2(q, ax);
    const float2 e = make_float2(ex2_approx(q.x), ex2_approx(q.y));                          // erfc(|x| / sqrt 2)
    const float2 t = fma2(e, make_float2(-0.5f, -0.5f), make_float2(0.5f, 0.5f));
    const float2 cdf = add2(make_float2(0.5f, 0.5f), make_float2(copysignf(t.x, x.x), copysignf(t.y, x.y)));   // Phi(x)
    g = mul2(x, cdf);
    float2 r = fma2(ax, make_float2(d4, d4), make_float2(d3, d3));
    r = fma2(r, ax, make_float2(d2, d2));
    r = fma2(r, ax, make_float2(d1, d1));
    r = fma2(r, ax, make_float2(d0, d0));                                                    // phi(x) = r e
    dg = fma2(mul2(x, r), e, cdf);                                                           // Phi(x) + x phi(x)
}
#else
__device__ __forceinline__ void gelu_fast2(float2 x, float2& g, float2& dg) {
    const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
    const float2 den = fma2(make_float2(0.47047f * 0.70710678118654752f, 0.47047f * 0.70710678118654752f), ax,
                            make_float2(1.0f, 1.0f));
    const float2 t = make_float2(rcp_approx(den.x), rcp_approx(den.y));
    float2 poly = fma2(make_float2(0.7478556f, 0.7478556f), t, make_float2(-0.0958798f, -0.0958798f));
    poly = fma2(poly, t, make_float2(0.3480242f, 0.3480242f));
    const float2 xx = mul2(x, mul2(x, make_float2(-0.72134752044448170f, -0.72134752044448170f)));
    const float2 e = make_float2(ex2_approx(xx.x), ex2_approx(xx.y));                       // exp(-x^2 / 2)
    const float2 erf_abs = fma2(mul2(poly, t), make_float2(-e.x, -e.y), make_float2(1.0f, 1.0f));
    const float2 serf = make_float2(copysignf(erf_abs.x, x.x), copysignf(erf_abs.y, x.y));
    const float2 cdf = fma2(serf, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));      // Phi(x)
    g = mul2(x, cdf);
    dg = fma2(mul2(x, make_float2(0.3989422804014327f, 0.3989422804014327f)), e, cdf);     // Phi(x) + x phi(x)
}
#endif
#endif  // __CUDACC__

// 16-bit dropout threshold and the matching unbiased rescale 1 / (1 - thresh16 / 65536)
inline uint32_t dropout_threshold(float p) {
    if (p <= 0.f) return 0u;
    double t = (double)p * 65536.0 + 0.5;
    if (t >= 65535.0) return 65535u;
    return (uint32_t)t;
}
inline float dropout_inv_keep(float p) {
    const uint32_t t = dropout_threshold(p);
    return t == 0u ? 1.0f : (float)(1.0 / (1.0 - (double)t / 65536.0));
}

}  // namespace mmb
