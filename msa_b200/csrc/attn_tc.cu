// Masked self-attention forward on 5th-generation tensor cores (tcgen05 + TMEM + TMA), head dim 64.
//
// One CTA = one 128-query tile of one (sequence, head); it walks the sequence's keys in tiles of 128:
//   S = Q K^T      tcgen05.mma 128 x 128 x 64 (A = Q, B = K, both K-major SWIZZLE_128B tiles filled by TMA)  -> TMEM
//   softmax        every thread owns ONE query row (= its TMEM lane): row max / sum need no shuffles;
//                  scale, additive key mask ((1 - m) * -10000 as the reference), online rescale, dropout
//   O += P V       P (bf16) goes through swizzled shared memory as the A operand, V is read in place as an
//                  MN-major B operand; the fp32 accumulator O lives in TMEM and is rescaled there
// Two CTAs are resident per SM (256 TMEM columns and ~113 KB shared memory each), so one CTA's softmax
// (CUDA cores) overlaps the other's MMAs (tensor pipe) without intra-CTA warp specialisation.
//
// Same contract as the mma.sync kernel in attn.cu (mmb_attn_fwd): Q|K|V read in place from [rows, 3H],
// context written to [rows, H], log2-domain LSE to [heads, rows].
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace mmb {

constexpr int kTQ = 128;   // queries per CTA
constexpr int kTK = 128;   // keys per tile
constexpr int kHD = 64;    // head dim
constexpr int kTcThreads = 256;
constexpr float kLog2eTc = 1.4426950408889634f;

constexpr int kQBytes = kTQ * kHD * 2;        // 16 KB
constexpr int kKBytes = kTK * kHD * 2;        // 16 KB
constexpr int kPBytes = kTQ * kTK * 2;        // 32 KB: two 64-key SWIZZLE_128B slabs
// Q 16 KB + K 16 KB (single buffer: free again as soon as S = Q K^T has retired, i.e. reloaded under the softmax)
// + V 2 x 16 KB + P 32 KB + key bias 1 KB + barriers: ~98 KB, two CTAs per SM
constexpr int kTcSmem = kQBytes + kKBytes + 2 * kKBytes + kPBytes + 6 * kTK * 4 + 128 + 1024;

struct AttnTcParams {
    __nv_bfloat16* ctx;
    float* lse;
    const float* keybias;
    const int* cu_seqlens;
    const int* kv_end;
    int H, nheads, total_rows;
    float scale_log2;
    uint32_t thresh;
    float inv_keep;
    uint64_t seed;
    uint32_t rng_stream;
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ptx_ex2(float x) {   // MUFU.EX2 (flush-to-zero): ex2(-inf) = 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void sts128_tc(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// 256 threads: warps 0-3 own key columns 0-63 of the score tile, warps 4-7 columns 64-127 (a warp may only touch
// the TMEM lane quarter 32*(warp%4)..+31, so the two groups share the rows and split the columns); 2 CTAs per SM
// = 16 resident warps, which is what hides the tcgen05.ld / MUFU / LDS latencies of the softmax.
__global__ void __launch_bounds__(kTcThreads, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sQ = ptx::smem_u32(smem);
    const uint32_t sK = sQ + kQBytes;                 // 16 KB
    const uint32_t sV = sK + kKBytes;                 // [2][16 KB]
    const uint32_t sP = sV + 2 * kKBytes;             // 32 KB
    float* sBias = reinterpret_cast<float*>(smem + kQBytes + 3 * kKBytes + kPBytes);   // [2][128]
    float* sMax = sBias + 2 * kTK;                                                      // [2][128] partial row maxima
    uint32_t* sKk = reinterpret_cast<uint32_t*>(sMax + 2 * kTK);                       // [2][128] dropout key-column keys
    const uint32_t bars = sP + kPBytes + 6 * kTK * 4;
    const uint32_t bar_q = bars, bar_k = bars + 8, bar_v0 = bars + 16, bar_v1 = bars + 24, bar_s = bars + 32, bar_o = bars + 40;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kQBytes + 3 * kKBytes + kPBytes + 6 * kTK * 4 + 64);

    const int seq = blockIdx.z, head = blockIdx.y, qt = blockIdx.x;
    const int row0 = p.cu_seqlens[seq];
    const int S = p.cu_seqlens[seq + 1] - row0;
    if (qt * kTQ >= S) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int half = warp >> 2;                       // which 64 key columns of the tile
    const int r = (warp & 3) * 32 + lane;             // query row inside the tile = TMEM lane
    int s_eff = S;                                    // keys behind kv_end are masked in whole tiles: skipped (exact)
    if (p.kv_end != nullptr) {
        const int e = p.kv_end[seq];
        if (e > 0 && e < S) s_eff = e;
    }
    const int nkv = (s_eff + kTK - 1) / kTK;

    if (tid == 0) {
        ptx::prefetch_tensormap(&tm_qkv);
        ptx::mbar_init(bar_q, 1);
        ptx::mbar_init(bar_k, 1);
        ptx::mbar_init(bar_v0, 1);
        ptx::mbar_init(bar_v1, 1);
        ptx::mbar_init(bar_s, 1);
        ptx::mbar_init(bar_o, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc<256>(ptx::smem_u32(tmem_slot));
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_bits = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_bits + half * 64;          // this thread's 64 score columns
    const uint32_t tO = tmem + 128 + lane_bits + half * 32;    // this thread's 32 output columns

    const int col_q = head * kHD, col_k = p.H + head * kHD, col_v = 2 * p.H + head * kHD;
    auto load_k = [&](int kt) {    // thread 0 only
        ptx::mbar_expect_tx(bar_k, kKBytes);
        ptx::tma_load_2d(sK, &tm_qkv, bar_k, col_k, row0 + kt * kTK);
    };
    auto load_v = [&](int kt) {    // thread 0 only
        const uint32_t bar = (kt & 1) ? bar_v1 : bar_v0;
        ptx::mbar_expect_tx(bar, kKBytes);
        ptx::tma_load_2d(sV + (kt & 1) * kKBytes, &tm_qkv, bar, col_v, row0 + kt * kTK);
    };
    if (tid == 0) {
        ptx::mbar_expect_tx(bar_q, kQBytes);
        ptx::tma_load_2d(sQ, &tm_qkv, bar_q, col_q, row0 + qt * kTQ);
        load_k(0);
        load_v(0);
        if (nkv > 1) load_v(1);
    }
    // key bias (log2 domain) of the first two tiles; keys beyond the sequence are -inf
    const uint32_t prob_base = (uint32_t)head * (uint32_t)p.total_rows + (uint32_t)row0;   // + index in the sequence
    sBias[tid] = tid < S ? p.keybias[row0 + tid] * kLog2eTc : -INFINITY;
    if (p.thresh != 0u) sKk[tid] = attn_drop_kkey(p.seed, p.rng_stream, prob_base + (uint32_t)tid);
    __syncthreads();

    // instruction descriptors: S: M=128,N=128, both K-major; PV: M=128,N=64, A K-major (P), B MN-major (V)
    const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTK >> 3) << 17) | ((uint32_t)(kTQ >> 4) << 24);
    const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(kHD >> 3) << 17) |
                             ((uint32_t)(kTQ >> 4) << 24);

    const int q = qt * kTQ + r;     // this thread's query row in the sequence
    float m_run = -INFINITY, l_run = 0.f;   // l_run: partial sum over this thread's column half
    const uint32_t qkey = p.thresh ? attn_drop_qkey(p.seed, p.rng_stream, prob_base + (uint32_t)q) : 0u;
    const uint32_t thresh32 = p.thresh << 16;
    const int r7 = r & 7;
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);

    for (int kt = 0; kt < nkv; ++kt) {
        const int st = kt & 1;
        // ---- S = Q K^T
        if (tid == 0) {
            if (kt == 0) ptx::mbar_wait(bar_q, 0);
            ptx::mbar_wait(bar_k, kt & 1);
            ptx::tc_fence_after();
#pragma unroll
            for (int k = 0; k < kHD / 16; ++k) {
                const uint64_t da = ptx::umma_desc_sw128(sQ + k * 32, 0, 1024);
                const uint64_t db = ptx::umma_desc_sw128(sK + k * 32, 0, 1024);
                ptx::umma_bf16(tmem, da, db, idesc_s, k > 0 ? 1u : 0u);
            }
            ptx::umma_commit(bar_s);
        }
        __syncwarp();
        ptx::mbar_wait(bar_s, kt & 1);
        ptx::tc_fence_after();
        if (tid == 0 && kt + 1 < nkv) load_k(kt + 1);   // K is free again: fetch the next tile under the softmax
        __syncwarp();
        const float4* bias4 = reinterpret_cast<const float4*>(sBias + st * kTK + half * 64);
        // ---- pass 1: partial row maximum over this thread's 64 columns, combined with the other half via smem
        float m_part = -INFINITY;
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
            uint32_t raw[32];
            ptx::tmem_ld_32x32(tS + c, raw);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 b = bias4[(c >> 2) + i];
                const float2 x0 = fma2(make_float2(__uint_as_float(raw[4 * i + 0]), __uint_as_float(raw[4 * i + 1])), sc2,
                                       make_float2(b.x, b.y));
                const float2 x1 = fma2(make_float2(__uint_as_float(raw[4 * i + 2]), __uint_as_float(raw[4 * i + 3])), sc2,
                                       make_float2(b.z, b.w));
                m_part = fmaxf(m_part, fmaxf(fmaxf(x0.x, x0.y), fmaxf(x1.x, x1.y)));
            }
        }
        sMax[half * kTK + r] = m_part;
        // previous P V must have retired before O is rescaled and P is overwritten (waited before the barrier so
        // that thread 0 can refill the freed V stage right after it)
        if (kt > 0) {
            ptx::mbar_wait(bar_o, (kt - 1) & 1);
            ptx::tc_fence_after();
        }
        // Lazy rescaling: the running reference maximum only moves (and O / l are only rescaled) when some row of the
        // tile exceeds it by more than 2^8 — otherwise the stale reference is kept: probabilities stay <= 256, which
        // fp32 sums and bf16 P hold without loss, and the TMEM round trip of O is skipped.  Block-uniform decision.
        const int moved = __syncthreads_or(m_part > m_run + 8.f);
        float m_new = m_run, corr = 1.f;
        if (moved) {
            m_new = fmaxf(m_run, fmaxf(sMax[r], sMax[kTK + r]));
        }
        // a fully masked-so-far row (all -inf) keeps m = -inf: use 0 as the reference to avoid inf - inf
        const float m_ref = m_new == -INFINITY ? 0.f : m_new;
        if (moved) corr = ptx_ex2(m_run - m_ref);
        if (kt > 0) {
            if (moved) {
                uint32_t raw[32];
                ptx::tmem_ld_32x32(tO, raw);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * corr);
                tmem_st_32x32(tO, raw);
                tmem_st_wait();
            }
            if (tid == 0 && kt + 1 < nkv) load_v(kt + 1);      // the V stage consumed by tile kt-1 is free
            if (kt + 1 < nkv && tid < kTK) {
                const int k0 = (kt + 1) * kTK;
                sBias[(st ^ 1) * kTK + tid] = (k0 + tid) < S ? p.keybias[row0 + k0 + tid] * kLog2eTc : -INFINITY;
                if (p.thresh != 0u) sKk[(st ^ 1) * kTK + tid] = attn_drop_kkey(p.seed, p.rng_stream, prob_base + (uint32_t)(k0 + tid));
            }
        } else if (nkv > 1 && tid < kTK) {
            // (already filled by the prologue: tiles 0 and 1 are loaded up front)
        }
        // ---- pass 2: probabilities -> bf16 P row of slab `half` (K-major SWIZZLE_128B), partial running sum
        float l_tile = 0.f;
        const float2 nm2 = make_float2(-m_ref, -m_ref);
        const uint32_t prow = sP + half * (kTQ * 128) + r * 128;
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
            uint32_t raw[32];
            ptx::tmem_ld_32x32(tS + c, raw);
            ptx::tmem_ld_wait();
            float pv[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 b = bias4[(c >> 2) + i];
                const float2 x0 = fma2(make_float2(__uint_as_float(raw[4 * i + 0]), __uint_as_float(raw[4 * i + 1])), sc2,
                                       add2(make_float2(b.x, b.y), nm2));
                const float2 x1 = fma2(make_float2(__uint_as_float(raw[4 * i + 2]), __uint_as_float(raw[4 * i + 3])), sc2,
                                       add2(make_float2(b.z, b.w), nm2));
                pv[4 * i + 0] = ptx_ex2(x0.x);
                pv[4 * i + 1] = ptx_ex2(x0.y);
                pv[4 * i + 2] = ptx_ex2(x1.x);
                pv[4 * i + 3] = ptx_ex2(x1.y);
            }
            {
                float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 32; i += 2) acc2 = add2(acc2, make_float2(pv[i], pv[i + 1]));
                l_tile += acc2.x + acc2.y;
            }
            if (p.thresh != 0u) {     // dropped entries become 0; the 1/(1-p) rescale is applied once in the epilogue
                const uint4* kk4 = reinterpret_cast<const uint4*>(sKk + st * kTK + half * 64 + c);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 kk = kk4[i];
                    pv[4 * i + 0] = attn_keep(qkey, kk.x, thresh32) ? pv[4 * i + 0] : 0.f;
                    pv[4 * i + 1] = attn_keep(qkey, kk.y, thresh32) ? pv[4 * i + 1] : 0.f;
                    pv[4 * i + 2] = attn_keep(qkey, kk.z, thresh32) ? pv[4 * i + 2] : 0.f;
                    pv[4 * i + 3] = attn_keep(qkey, kk.w, thresh32) ? pv[4 * i + 3] : 0.f;
                }
            }
            const int ch0 = c >> 3;   // first 16-byte chunk of this 32-key group inside the 64-key slab row
#pragma unroll
            for (int j = 0; j < 4; ++j)
                sts128_tc(prow + (((ch0 + j) ^ r7) << 4), pack_bf16x2(pv[8 * j + 0], pv[8 * j + 1]),
                          pack_bf16x2(pv[8 * j + 2], pv[8 * j + 3]), pack_bf16x2(pv[8 * j + 4], pv[8 * j + 5]),
                          pack_bf16x2(pv[8 * j + 6], pv[8 * j + 7]));
        }
        l_run = l_run * corr + l_tile;
        m_run = m_new;
        // make P visible to the tensor core (async proxy), order the TMEM accesses, then issue O += P V
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            ptx::mbar_wait(st ? bar_v1 : bar_v0, (kt >> 1) & 1);
            ptx::tc_fence_after();
#pragma unroll
            for (int k = 0; k < kTK / 16; ++k) {
                const uint64_t da = ptx::umma_desc_sw128(sP + (k >> 2) * (kTQ * 128) + (k & 3) * 32, 0, 1024);
                const uint64_t db = ptx::umma_desc_sw128(sV + st * kKBytes + k * 2048, 8192, 1024);
                ptx::umma_bf16(tmem + 128, da, db, idesc_o, (kt > 0 || k > 0) ? 1u : 0u);
            }
            ptx::umma_commit(bar_o);
        }
        __syncwarp();
    }
    // ---- epilogue: O * inv_keep / l -> bf16 context row (this thread: 32 of the 64 columns), LSE
    sMax[half * kTK + r] = l_run;          // combine the two column halves' partial sums
    ptx::mbar_wait(bar_o, (nkv - 1) & 1);
    ptx::tc_fence_after();
    __syncthreads();
    const float l_tot = sMax[r] + sMax[kTK + r];
    const float inv_l = p.inv_keep / l_tot;
    {
        uint32_t raw[32];
        ptx::tmem_ld_32x32(tO, raw);
        ptx::tmem_ld_wait();
        if (q < S) {
            __nv_bfloat16* out = p.ctx + (int64_t)(row0 + q) * p.H + head * kHD + half * 32;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 u;
                u.x = pack_bf16x2(__uint_as_float(raw[8 * j + 0]) * inv_l, __uint_as_float(raw[8 * j + 1]) * inv_l);
                u.y = pack_bf16x2(__uint_as_float(raw[8 * j + 2]) * inv_l, __uint_as_float(raw[8 * j + 3]) * inv_l);
                u.z = pack_bf16x2(__uint_as_float(raw[8 * j + 4]) * inv_l, __uint_as_float(raw[8 * j + 5]) * inv_l);
                u.w = pack_bf16x2(__uint_as_float(raw[8 * j + 6]) * inv_l, __uint_as_float(raw[8 * j + 7]) * inv_l);
                *reinterpret_cast<uint4*>(out + 8 * j) = u;
            }
        }
    }
    if (half == 0 && q < S && p.lse != nullptr) p.lse[(int64_t)head * p.total_rows + row0 + q] = m_run + log2f(l_tot);
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<256>(tmem);
    }
}

int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1);

int launch_attn_fwd_tc(const mmb_attn_args* a, cudaStream_t stream) {
    CUtensorMap tm;
    int rc = make_tmap_bf16(&tm, a->qkv, (uint64_t)3 * a->H, (uint64_t)a->total_rows, (uint64_t)3 * a->H, 64, 128);
    if (rc != MMB_OK) return rc;
    AttnTcParams p;
    p.ctx = (__nv_bfloat16*)a->ctx;
    p.lse = a->lse;
    p.keybias = a->keybias;
    p.cu_seqlens = a->cu_seqlens;
    p.kv_end = a->kv_end;
    p.H = a->H;
    p.nheads = a->nheads;
    p.total_rows = a->total_rows;
    p.scale_log2 = kLog2eTc / sqrtf((float)kHD);
    p.thresh = dropout_threshold(a->p_drop);
    p.inv_keep = dropout_inv_keep(a->p_drop);
    p.seed = a->seed;
    p.rng_stream = a->rng_stream;
    MMB_ENSURE_SMEM(kTcSmem, attn_fwd_tc_kernel);
    dim3 grid((a->max_seqlen + kTQ - 1) / kTQ, a->nheads, a->nseq);
    attn_fwd_tc_kernel<<<grid, kTcThreads, kTcSmem, stream>>>(tm, p);
    return check_launch("attn_fwd_tc_kernel");
}

}  // namespace mmb
