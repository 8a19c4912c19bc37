// Masked self-attention BACKWARD on 5th-generation tensor cores (tcgen05 + TMEM + TMA), head dim 64.
// Autograd of modeling_bert.py:115-140 (eager_attention_forward) for the packed batch; same contract as
// mmb_attn_bwd in include/mmbert_sm100.h (dQ|dK|dV written into dqkv [rows, 3H], D = rowsum(dO ∘ O) precomputed).
//
// Two kernels, neither of which transposes anything or needs atomics:
//   dQ  kernel  (CTA = 128 queries = TMEM lanes; steps of 64 keys)
//        S  = Q K^T,  dP  = dO V^T   -> TMEM;  dS  = P ∘ (dP - D)  -> bf16 -> swizzled smem;  dQ += dS K
//   dKV kernel  (CTA = 128 keys    = TMEM lanes; steps of 64 queries)
//        S^T = K Q^T, dP^T = V dO^T  -> TMEM;  P^T, dS^T -> smem;  dV += P^T dO,  dK += dS^T Q
// Every thread owns one TMEM lane (row) and 32 of the step's 64 columns.  The probabilities are rebuilt from the
// forward's log-sum-exp (no running maximum), so the element-wise stage is ~7 instructions per probability:
// packed FFMA2/FADD2/FMUL2, one MUFU.EX2, and a one-IMAD dropout decision (common.cuh: attn_keep).
//
// Pipeline per CTA (two CTAs are resident per SM so one's element-wise stage overlaps the other's MMAs):
//   one elected thread issues   [accumulating MMAs of step i]  [score MMAs of step i+1]  commit
//   everyone waits for that commit, runs the element-wise stage of step i+1, one __syncthreads, repeat.
// tcgen05.mma retire in issue order, so the single commit per step also tells that step i's operand buffers
// (the TMA double buffers and the P / dS tiles) are free again.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace mmb {

int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1);

namespace {

constexpr int kRows = 128;           // TMEM lanes per CTA
constexpr int kStep = 64;            // columns per step
constexpr int kHd = 64;              // head dim
constexpr int kBwdThreads = 256;
constexpr float kLog2eB = 1.4426950408889634f;
constexpr int kBig = kRows * kHd * 2;     // 16 KB: 128-row operand tile (128 B rows, SWIZZLE_128B)
constexpr int kSmall = kStep * kHd * 2;   // 8 KB: 64-row operand tile
// dQ kernel: Q, dO (big) + 2 x (K, V) (small) + dS (big) = 80 KB;  dKV kernel: K, V + 2 x (Q, dO) + P^T + dS^T = 96 KB
constexpr int kDqSmem = 3 * kBig + 4 * kSmall + 2048 + 1024;
constexpr int kDkvSmem = 4 * kBig + 4 * kSmall + 2048 + 1024;

struct BwdTcParams {
    __nv_bfloat16* dqkv;
    const float* lse;
    const float* dsum;
    const float* keybias;
    const int* cu_seqlens;
    const int* kv_end;
    int H, nheads, total_rows;
    float scale_log2, scale;
    uint32_t thresh32;
    float inv_keep;
    uint64_t seed;
    uint32_t rng_stream;
};

__device__ __forceinline__ void sts128_b(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// instruction descriptor: fp32 accumulate, bf16 x bf16, M = 128, N = 64; B operand K-major or MN-major
__device__ __forceinline__ constexpr uint32_t idesc_m128_n64(bool b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | ((uint32_t)(kStep >> 3) << 17) |
           ((uint32_t)(kRows >> 4) << 24);
}

// D[128 x 64] (+)= A[128 x 64] * B^T,  A and B both K-major tiles with 128-byte rows (contraction over the 64 columns)
__device__ __forceinline__ void mma_kk(uint32_t tmem_d, uint32_t sA, uint32_t sB, bool accumulate) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
        ptx::umma_bf16(tmem_d, ptx::umma_desc_sw128(sA + k * 32, 0, 1024), ptx::umma_desc_sw128(sB + k * 32, 0, 1024),
                       idesc_m128_n64(false), (accumulate || k > 0) ? 1u : 0u);
}
// D[128 x 64] (+)= A[128 x 64] * B,  A K-major, B = [64 (contraction rows)][64] read in place as an MN-major operand
__device__ __forceinline__ void mma_kmn(uint32_t tmem_d, uint32_t sA, uint32_t sB, bool accumulate) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
        ptx::umma_bf16(tmem_d, ptx::umma_desc_sw128(sA + k * 32, 0, 1024), ptx::umma_desc_sw128(sB + k * 2048, 8192, 1024),
                       idesc_m128_n64(true), (accumulate || k > 0) ? 1u : 0u);
}

__device__ __forceinline__ int effective_keys_tc(const int* kv_end, int seq, int S) {
    if (kv_end == nullptr) return S;
    const int e = kv_end[seq];
    return (e <= 0 || e > S) ? S : e;
}

struct Smem {
    uint32_t base;
    uint8_t* ptr;
};
__device__ __forceinline__ Smem align_smem(uint8_t* raw) {
    uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    return Smem{ptx::smem_u32(p), p};
}

// ============================================================================================== dQ
template <bool kDrop>
__global__ void __launch_bounds__(kBwdThreads, 2)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv128, const __grid_constant__ CUtensorMap tm_qkv64,
                      const __grid_constant__ CUtensorMap tm_do128, const BwdTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const Smem sm = align_smem(smem_raw);
    const uint32_t sQ = sm.base, sdO = sQ + kBig, sK = sdO + kBig, sV = sK + 2 * kSmall, sdS = sV + 2 * kSmall;
    uint8_t* aux = sm.ptr + 3 * kBig + 4 * kSmall;
    float* sBias = reinterpret_cast<float*>(aux);                 // [2][64]
    uint32_t* sKk = reinterpret_cast<uint32_t*>(aux + 512);       // [2][64] dropout key-column keys
    const uint32_t bars = sdS + kBig + 1024;
    const uint32_t bar_qdo = bars, bar_kv0 = bars + 8, bar_kv1 = bars + 16, bar_mma = bars + 24;
    const uint32_t aux_u32 = sdS + kBig;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + 1024 + 64);

    const int seq = blockIdx.z, head = blockIdx.y, qt = blockIdx.x;
    const int row0 = p.cu_seqlens[seq];
    const int S = p.cu_seqlens[seq + 1] - row0;
    if (qt * kRows >= S) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int half = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const int nkt = (effective_keys_tc(p.kv_end, seq, S) + kStep - 1) / kStep;
    const uint32_t prob_base = ((uint32_t)seq * (uint32_t)p.nheads + (uint32_t)head) * (uint32_t)S;

    if (tid == 0) {
        ptx::prefetch_tensormap(&tm_qkv128);
        ptx::prefetch_tensormap(&tm_qkv64);
        ptx::prefetch_tensormap(&tm_do128);
        ptx::mbar_init(bar_qdo, 1);
        ptx::mbar_init(bar_kv0, 1);
        ptx::mbar_init(bar_kv1, 1);
        ptx::mbar_init(bar_mma, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc<256>(ptx::smem_u32(tmem_slot));
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_bits = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_bits + half * 32;            // S   : columns [0, 64)
    const uint32_t tdP = tmem + 64 + lane_bits + half * 32;      // dP  : columns [64, 128)
    const uint32_t tdQ = tmem + 128 + lane_bits + half * 32;     // dQ  : columns [128, 192)

    const int col_q = head * kHd, col_k = p.H + head * kHd, col_v = 2 * p.H + head * kHd;
    auto load_kv = [&](int kt) {   // thread 0 only
        const uint32_t bar = (kt & 1) ? bar_kv1 : bar_kv0;
        ptx::mbar_expect_tx(bar, 2 * kSmall);
        ptx::tma_load_2d(sK + (kt & 1) * kSmall, &tm_qkv64, bar, col_k, row0 + kt * kStep);
        ptx::tma_load_2d(sV + (kt & 1) * kSmall, &tm_qkv64, bar, col_v, row0 + kt * kStep);
    };
    auto load_cols = [&](int kt, int i) {   // bias / dropout key of key column i of step kt
        const int k = kt * kStep + i;
        sBias[(kt & 1) * kStep + i] = k < S ? p.keybias[row0 + k] * kLog2eB : -INFINITY;
        if (kDrop) sKk[(kt & 1) * kStep + i] = attn_drop_kkey(p.seed, p.rng_stream, prob_base + (uint32_t)k);
    };
    if (tid == 0) {
        ptx::mbar_expect_tx(bar_qdo, 2 * kBig);
        ptx::tma_load_2d(sQ, &tm_qkv128, bar_qdo, col_q, row0 + qt * kRows);
        ptx::tma_load_2d(sdO, &tm_do128, bar_qdo, head * kHd, row0 + qt * kRows);
        load_kv(0);
        if (nkt > 1) load_kv(1);
    }
    if (tid < 2 * kStep && (tid >> 6) < nkt) load_cols(tid >> 6, tid & 63);
    const int q = qt * kRows + r;
    const float negL = q < S ? -p.lse[(int64_t)head * p.total_rows + row0 + q] : -INFINITY;   // padded rows: P = 0
    const float negD = q < S ? -p.dsum[(int64_t)head * p.total_rows + row0 + q] : 0.f;
    const uint32_t qkey = kDrop ? attn_drop_qkey(p.seed, p.rng_stream, prob_base + (uint32_t)q) : 0u;
    __syncthreads();
    if (tid == 0) {
        ptx::mbar_wait(bar_qdo, 0);
        ptx::mbar_wait(bar_kv0, 0);
        ptx::tc_fence_after();
        mma_kk(tmem, sQ, sK, false);            // S_0  = Q K_0^T
        mma_kk(tmem + 64, sdO, sV, false);      // dP_0 = dO V_0^T
        ptx::umma_commit(bar_mma);
    }
    __syncwarp();
    const int r7 = r & 7;
    const uint32_t ds_row = sdS + r * 128;
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), negL2 = make_float2(negL, negL),
                 negD2 = make_float2(negD, negD), ik2 = make_float2(p.inv_keep, p.inv_keep);

    for (int kt = 0; kt < nkt; ++kt) {
        const int st = kt & 1;
        ptx::mbar_wait(bar_mma, kt & 1);
        ptx::tc_fence_after();
        // the K/V stage and the column vectors of step kt-1 are free: refill them for step kt+1
        if (kt >= 1 && kt + 1 < nkt) {
            if (tid == 0) load_kv(kt + 1);
            if (tid >= 64 && tid < 128) load_cols(kt + 1, tid - 64);
        }
        uint32_t s_raw[32], dp_raw[32];
        ptx::tmem_ld_32x32(tS, s_raw);
        ptx::tmem_ld_32x32(tdP, dp_raw);
        ptx::tmem_ld_wait();
        const uint32_t bias_a = aux_u32 + (st * kStep + half * 32) * 4, kk_a = bias_a + 512;
#pragma unroll
        for (int j = 0; j < 4; ++j) {          // 8 columns -> one 16-byte chunk of the dS row
            uint32_t packed[4];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int i = 2 * j + h2;      // float4 index: columns 4i .. 4i+3
                const float4 b = lds_f4(bias_a + i * 16);
                float2 x0 = fma2(make_float2(__uint_as_float(s_raw[4 * i]), __uint_as_float(s_raw[4 * i + 1])), sc2,
                                 make_float2(b.x, b.y));
                float2 x1 = fma2(make_float2(__uint_as_float(s_raw[4 * i + 2]), __uint_as_float(s_raw[4 * i + 3])), sc2,
                                 make_float2(b.z, b.w));
                x0 = add2(x0, negL2);
                x1 = add2(x1, negL2);
                const float2 p0 = make_float2(ex2_approx(x0.x), ex2_approx(x0.y));
                const float2 p1 = make_float2(ex2_approx(x1.x), ex2_approx(x1.y));
                float2 d0 = make_float2(__uint_as_float(dp_raw[4 * i]), __uint_as_float(dp_raw[4 * i + 1]));
                float2 d1 = make_float2(__uint_as_float(dp_raw[4 * i + 2]), __uint_as_float(dp_raw[4 * i + 3]));
                if (kDrop) {
                    const uint4 kk = lds_u4(kk_a + i * 16);
                    d0.x = attn_keep(qkey, kk.x, p.thresh32) ? d0.x : 0.f;
                    d0.y = attn_keep(qkey, kk.y, p.thresh32) ? d0.y : 0.f;
                    d1.x = attn_keep(qkey, kk.z, p.thresh32) ? d1.x : 0.f;
                    d1.y = attn_keep(qkey, kk.w, p.thresh32) ? d1.y : 0.f;
                }
                // dS = P ∘ (dP_dropped / keep - D)
                const float2 e0 = mul2(p0, fma2(d0, ik2, negD2));
                const float2 e1 = mul2(p1, fma2(d1, ik2, negD2));
                packed[2 * h2] = pack_bf16x2(e0.x, e0.y);
                packed[2 * h2 + 1] = pack_bf16x2(e1.x, e1.y);
            }
            sts128_b(ds_row + (((half * 4 + j) ^ r7) << 4), packed[0], packed[1], packed[2], packed[3]);
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            ptx::tc_fence_after();
            mma_kmn(tmem + 128, sdS, sK + st * kSmall, kt > 0);        // dQ += dS K_kt
            if (kt + 1 < nkt) {
                ptx::mbar_wait((st ^ 1) ? bar_kv1 : bar_kv0, ((kt + 1) >> 1) & 1);
                ptx::tc_fence_after();
                mma_kk(tmem, sQ, sK + (st ^ 1) * kSmall, false);
                mma_kk(tmem + 64, sdO, sV + (st ^ 1) * kSmall, false);
            }
            ptx::umma_commit(bar_mma);
        }
        __syncwarp();
    }
    ptx::mbar_wait(bar_mma, nkt & 1);
    ptx::tc_fence_after();
    {
        uint32_t raw[32];
        ptx::tmem_ld_32x32(tdQ, raw);
        ptx::tmem_ld_wait();
        if (q < S) {
            __nv_bfloat16* out = p.dqkv + (int64_t)(row0 + q) * (3 * p.H) + head * kHd + half * 32;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 u;
                u.x = pack_bf16x2(__uint_as_float(raw[8 * j + 0]) * p.scale, __uint_as_float(raw[8 * j + 1]) * p.scale);
                u.y = pack_bf16x2(__uint_as_float(raw[8 * j + 2]) * p.scale, __uint_as_float(raw[8 * j + 3]) * p.scale);
                u.z = pack_bf16x2(__uint_as_float(raw[8 * j + 4]) * p.scale, __uint_as_float(raw[8 * j + 5]) * p.scale);
                u.w = pack_bf16x2(__uint_as_float(raw[8 * j + 6]) * p.scale, __uint_as_float(raw[8 * j + 7]) * p.scale);
                *reinterpret_cast<uint4*>(out + 8 * j) = u;
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<256>(tmem);
    }
}

// ============================================================================================== dK, dV
template <bool kDrop>
__global__ void __launch_bounds__(kBwdThreads, 2)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv128, const __grid_constant__ CUtensorMap tm_qkv64,
                       const __grid_constant__ CUtensorMap tm_do64, const BwdTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const Smem sm = align_smem(smem_raw);
    const uint32_t sK = sm.base, sV = sK + kBig, sQ = sV + kBig, sdO = sQ + 2 * kSmall, sP = sdO + 2 * kSmall, sdS = sP + kBig;
    uint8_t* aux = sm.ptr + 4 * kBig + 4 * kSmall;
    float* sNegL = reinterpret_cast<float*>(aux);                 // [2][64]
    float* sNegD = reinterpret_cast<float*>(aux + 512);           // [2][64]
    uint32_t* sQk = reinterpret_cast<uint32_t*>(aux + 1024);      // [2][64] dropout query-row keys
    const uint32_t bars = sdS + kBig + 1536;
    const uint32_t bar_kv = bars, bar_q0 = bars + 8, bar_q1 = bars + 16, bar_mma = bars + 24;
    const uint32_t aux_u32 = sdS + kBig;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + 1536 + 64);

    const int seq = blockIdx.z, head = blockIdx.y, kt = blockIdx.x;
    const int row0 = p.cu_seqlens[seq];
    const int S = p.cu_seqlens[seq + 1] - row0;
    if (kt * kRows >= S) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int half = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const int k = kt * kRows + r;                     // this thread's key
    const int64_t ld = 3 * (int64_t)p.H;
    if (kt * kRows >= effective_keys_tc(p.kv_end, seq, S)) {
        // every key of this tile is masked: P == 0 exactly, so dK = dV = 0
        if (k < S) {
            __nv_bfloat16* o = p.dqkv + (int64_t)(row0 + k) * ld + p.H + head * kHd + half * 32;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                *reinterpret_cast<uint4*>(o + 8 * j) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(o + p.H + 8 * j) = make_uint4(0, 0, 0, 0);
            }
        }
        return;
    }
    const int nqt = (S + kStep - 1) / kStep;
    const uint32_t prob_base = ((uint32_t)seq * (uint32_t)p.nheads + (uint32_t)head) * (uint32_t)S;

    if (tid == 0) {
        ptx::prefetch_tensormap(&tm_qkv128);
        ptx::prefetch_tensormap(&tm_qkv64);
        ptx::prefetch_tensormap(&tm_do64);
        ptx::mbar_init(bar_kv, 1);
        ptx::mbar_init(bar_q0, 1);
        ptx::mbar_init(bar_q1, 1);
        ptx::mbar_init(bar_mma, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc<256>(ptx::smem_u32(tmem_slot));
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_bits = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_bits + half * 32;            // S^T  : columns [0, 64)
    const uint32_t tdP = tmem + 64 + lane_bits + half * 32;      // dP^T : columns [64, 128)
    const uint32_t tdV = tmem + 128 + lane_bits + half * 32;     // dV   : columns [128, 192)
    const uint32_t tdK = tmem + 192 + lane_bits + half * 32;     // dK   : columns [192, 256)

    const int col_q = head * kHd, col_k = p.H + head * kHd, col_v = 2 * p.H + head * kHd;
    auto load_q = [&](int it) {   // thread 0 only
        const uint32_t bar = (it & 1) ? bar_q1 : bar_q0;
        ptx::mbar_expect_tx(bar, 2 * kSmall);
        ptx::tma_load_2d(sQ + (it & 1) * kSmall, &tm_qkv64, bar, col_q, row0 + it * kStep);
        ptx::tma_load_2d(sdO + (it & 1) * kSmall, &tm_do64, bar, head * kHd, row0 + it * kStep);
    };
    auto load_cols = [&](int it, int i) {   // -LSE, -D and the dropout key of query column i of step it
        const int q = it * kStep + i;
        const int64_t g = (int64_t)head * p.total_rows + row0 + q;
        sNegL[(it & 1) * kStep + i] = q < S ? -p.lse[g] : -INFINITY;      // padded queries: P = 0
        sNegD[(it & 1) * kStep + i] = q < S ? -p.dsum[g] : 0.f;
        if (kDrop) sQk[(it & 1) * kStep + i] = attn_drop_qkey(p.seed, p.rng_stream, prob_base + (uint32_t)q);
    };
    if (tid == 0) {
        ptx::mbar_expect_tx(bar_kv, 2 * kBig);
        ptx::tma_load_2d(sK, &tm_qkv128, bar_kv, col_k, row0 + kt * kRows);
        ptx::tma_load_2d(sV, &tm_qkv128, bar_kv, col_v, row0 + kt * kRows);
        load_q(0);
        if (nqt > 1) load_q(1);
    }
    if (tid < 2 * kStep && (tid >> 6) < nqt) load_cols(tid >> 6, tid & 63);
    const float bias = k < S ? p.keybias[row0 + k] * kLog2eB : -INFINITY;        // keys beyond the sequence: P = 0
    const uint32_t kkey = kDrop ? attn_drop_kkey(p.seed, p.rng_stream, prob_base + (uint32_t)k) : 0u;
    __syncthreads();
    if (tid == 0) {
        ptx::mbar_wait(bar_kv, 0);
        ptx::mbar_wait(bar_q0, 0);
        ptx::tc_fence_after();
        mma_kk(tmem, sK, sQ, false);            // S^T_0  = K Q_0^T
        mma_kk(tmem + 64, sV, sdO, false);      // dP^T_0 = V dO_0^T
        ptx::umma_commit(bar_mma);
    }
    __syncwarp();
    const int r7 = r & 7;
    const uint32_t p_row = sP + r * 128, ds_row = sdS + r * 128;
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), bias2 = make_float2(bias, bias),
                 ik2 = make_float2(p.inv_keep, p.inv_keep);

    for (int it = 0; it < nqt; ++it) {
        const int st = it & 1;
        ptx::mbar_wait(bar_mma, it & 1);
        ptx::tc_fence_after();
        if (it >= 1 && it + 1 < nqt) {
            if (tid == 0) load_q(it + 1);
            if (tid >= 64 && tid < 128) load_cols(it + 1, tid - 64);
        }
        uint32_t s_raw[32], dp_raw[32];
        ptx::tmem_ld_32x32(tS, s_raw);
        ptx::tmem_ld_32x32(tdP, dp_raw);
        ptx::tmem_ld_wait();
        const uint32_t negL_a = aux_u32 + (st * kStep + half * 32) * 4, negD_a = negL_a + 512, qk_a = negL_a + 1024;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t pk[4], dk[4];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int i = 2 * j + h2;
                const float4 nl = lds_f4(negL_a + i * 16), nd = lds_f4(negD_a + i * 16);
                float2 x0 = fma2(make_float2(__uint_as_float(s_raw[4 * i]), __uint_as_float(s_raw[4 * i + 1])), sc2, bias2);
                float2 x1 = fma2(make_float2(__uint_as_float(s_raw[4 * i + 2]), __uint_as_float(s_raw[4 * i + 3])), sc2, bias2);
                x0 = add2(x0, make_float2(nl.x, nl.y));
                x1 = add2(x1, make_float2(nl.z, nl.w));
                const float2 p0 = make_float2(ex2_approx(x0.x), ex2_approx(x0.y));
                const float2 p1 = make_float2(ex2_approx(x1.x), ex2_approx(x1.y));
                float2 d0 = make_float2(__uint_as_float(dp_raw[4 * i]), __uint_as_float(dp_raw[4 * i + 1]));
                float2 d1 = make_float2(__uint_as_float(dp_raw[4 * i + 2]), __uint_as_float(dp_raw[4 * i + 3]));
                float2 pd0 = p0, pd1 = p1;     // dropped probabilities feed dV (the 1/keep rescale is applied to dV once)
                if (kDrop) {
                    const uint4 qk = lds_u4(qk_a + i * 16);
                    const bool k0 = attn_keep(qk.x, kkey, p.thresh32), k1 = attn_keep(qk.y, kkey, p.thresh32),
                               k2 = attn_keep(qk.z, kkey, p.thresh32), k3 = attn_keep(qk.w, kkey, p.thresh32);
                    d0.x = k0 ? d0.x : 0.f;
                    d0.y = k1 ? d0.y : 0.f;
                    d1.x = k2 ? d1.x : 0.f;
                    d1.y = k3 ? d1.y : 0.f;
                    pd0.x = k0 ? p0.x : 0.f;
                    pd0.y = k1 ? p0.y : 0.f;
                    pd1.x = k2 ? p1.x : 0.f;
                    pd1.y = k3 ? p1.y : 0.f;
                }
                const float2 e0 = mul2(p0, fma2(d0, ik2, make_float2(nd.x, nd.y)));
                const float2 e1 = mul2(p1, fma2(d1, ik2, make_float2(nd.z, nd.w)));
                pk[2 * h2] = pack_bf16x2(pd0.x, pd0.y);
                pk[2 * h2 + 1] = pack_bf16x2(pd1.x, pd1.y);
                dk[2 * h2] = pack_bf16x2(e0.x, e0.y);
                dk[2 * h2 + 1] = pack_bf16x2(e1.x, e1.y);
            }
            const uint32_t off = (uint32_t)(((half * 4 + j) ^ r7) << 4);
            sts128_b(p_row + off, pk[0], pk[1], pk[2], pk[3]);
            sts128_b(ds_row + off, dk[0], dk[1], dk[2], dk[3]);
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            ptx::tc_fence_after();
            mma_kmn(tmem + 128, sP, sdO + st * kSmall, it > 0);        // dV += P^T  dO_it
            mma_kmn(tmem + 192, sdS, sQ + st * kSmall, it > 0);        // dK += dS^T Q_it
            if (it + 1 < nqt) {
                ptx::mbar_wait((st ^ 1) ? bar_q1 : bar_q0, ((it + 1) >> 1) & 1);
                ptx::tc_fence_after();
                mma_kk(tmem, sK, sQ + (st ^ 1) * kSmall, false);
                mma_kk(tmem + 64, sV, sdO + (st ^ 1) * kSmall, false);
            }
            ptx::umma_commit(bar_mma);
        }
        __syncwarp();
    }
    ptx::mbar_wait(bar_mma, nqt & 1);
    ptx::tc_fence_after();
    {
        uint32_t rv[32], rk[32];
        ptx::tmem_ld_32x32(tdV, rv);
        ptx::tmem_ld_32x32(tdK, rk);
        ptx::tmem_ld_wait();
        if (k < S) {
            __nv_bfloat16* oK = p.dqkv + (int64_t)(row0 + k) * ld + p.H + head * kHd + half * 32;
            __nv_bfloat16* oV = oK + p.H;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 u, v;
                u.x = pack_bf16x2(__uint_as_float(rk[8 * j + 0]) * p.scale, __uint_as_float(rk[8 * j + 1]) * p.scale);
                u.y = pack_bf16x2(__uint_as_float(rk[8 * j + 2]) * p.scale, __uint_as_float(rk[8 * j + 3]) * p.scale);
                u.z = pack_bf16x2(__uint_as_float(rk[8 * j + 4]) * p.scale, __uint_as_float(rk[8 * j + 5]) * p.scale);
                u.w = pack_bf16x2(__uint_as_float(rk[8 * j + 6]) * p.scale, __uint_as_float(rk[8 * j + 7]) * p.scale);
                v.x = pack_bf16x2(__uint_as_float(rv[8 * j + 0]) * p.inv_keep, __uint_as_float(rv[8 * j + 1]) * p.inv_keep);
                v.y = pack_bf16x2(__uint_as_float(rv[8 * j + 2]) * p.inv_keep, __uint_as_float(rv[8 * j + 3]) * p.inv_keep);
                v.z = pack_bf16x2(__uint_as_float(rv[8 * j + 4]) * p.inv_keep, __uint_as_float(rv[8 * j + 5]) * p.inv_keep);
                v.w = pack_bf16x2(__uint_as_float(rv[8 * j + 6]) * p.inv_keep, __uint_as_float(rv[8 * j + 7]) * p.inv_keep);
                *reinterpret_cast<uint4*>(oK + 8 * j) = u;
                *reinterpret_cast<uint4*>(oV + 8 * j) = v;
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<256>(tmem);
    }
}

template <bool kDrop>
int launch_both(const CUtensorMap& q128, const CUtensorMap& q64, const CUtensorMap& do128, const CUtensorMap& do64,
                const BwdTcParams& p, dim3 grid, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        MMB_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<kDrop>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDqSmem));
        MMB_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<kDrop>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDkvSmem));
        attr_set = true;
    }
    attn_bwd_dkv_tc_kernel<kDrop><<<grid, kBwdThreads, kDkvSmem, stream>>>(q128, q64, do64, p);
    int rc = check_launch("attn_bwd_dkv_tc_kernel");
    if (rc != MMB_OK) return rc;
    attn_bwd_dq_tc_kernel<kDrop><<<grid, kBwdThreads, kDqSmem, stream>>>(q128, q64, do128, p);
    return check_launch("attn_bwd_dq_tc_kernel");
}

}  // namespace

// dsum (D = rowsum(dO ∘ O)) must already be in a->dsum (attn.cu: attn_bwd_dsum_kernel).
int launch_attn_bwd_tc(const mmb_attn_args* a, cudaStream_t stream) {
    CUtensorMap q128, q64, do128, do64;
    const uint64_t rows = (uint64_t)a->total_rows, H = (uint64_t)a->H;
    int rc = make_tmap_bf16(&q128, a->qkv, 3 * H, rows, 3 * H, 64, 128);
    if (rc == MMB_OK) rc = make_tmap_bf16(&q64, a->qkv, 3 * H, rows, 3 * H, 64, 64);
    if (rc == MMB_OK) rc = make_tmap_bf16(&do128, a->dctx, H, rows, H, 64, 128);
    if (rc == MMB_OK) rc = make_tmap_bf16(&do64, a->dctx, H, rows, H, 64, 64);
    if (rc != MMB_OK) return rc;
    BwdTcParams p;
    p.dqkv = (__nv_bfloat16*)a->dqkv;
    p.lse = a->lse;
    p.dsum = a->dsum;
    p.keybias = a->keybias;
    p.cu_seqlens = a->cu_seqlens;
    p.kv_end = a->kv_end;
    p.H = a->H;
    p.nheads = a->nheads;
    p.total_rows = a->total_rows;
    p.scale = 1.0f / sqrtf((float)kHd);
    p.scale_log2 = p.scale * kLog2eB;
    p.thresh32 = dropout_threshold(a->p_drop) << 16;
    p.inv_keep = dropout_inv_keep(a->p_drop);
    p.seed = a->seed;
    p.rng_stream = a->rng_stream;
    dim3 grid((a->max_seqlen + kRows - 1) / kRows, a->nheads, a->nseq);
    return p.thresh32 ? launch_both<true>(q128, q64, do128, do64, p, grid, stream)
                      : launch_both<false>(q128, q64, do128, do64, p, grid, stream);
}

}  // namespace mmb
