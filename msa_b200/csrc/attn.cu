// Masked self-attention over packed variable-length sequences (text | frames), head dim 64.
//
// Round-1 implementation: flash-style (online softmax, nothing of size S x S is materialised) on the legacy
// warp-level tensor path (mma.sync m16n8k16 bf16, ldmatrix, cp.async double buffering).  Attention is
// 1.5 % (MOSI shape) to 8 % (MOSEI-unaligned) of the path's FLOPs; the GEMMs that hold the other >90 % are
// on tcgen05 (gemm_tcgen05.cu).  A tcgen05/TMEM version of this kernel is the planned follow-up.
//
// Replaces BertSelfAttention's score/softmax/dropout/context chain:
//   modeling_bert.py:115-140 (eager_attention_forward) / sdpa_attention.py:40-102, called from :168-207,
//   with the additive mask of MMBertForPretraining.py:57-154,246-250 ((1-m) * -10000 per key, text ⊕ frames).
// Q, K, V are read in place from the fused QKV projection output [rows, 3H]; the context is written [rows, H].
//
// Backward = two kernels that never transpose through shared memory:
//   dKV kernel: rows = keys     S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q
//   dQ  kernel: rows = queries  S   = Q K^T, dP   = dO V^T, dQ += dS K
// plus a row-dot preprocess D = rowsum(dO ∘ O).
#include "common.cuh"
#include "ptx.cuh"

namespace mmb {

constexpr int kD = 64;       // head dim
constexpr int kTile = 64;    // rows per tile (queries or keys)
constexpr int kAttnThreads = 128;
constexpr float kLog2e = 1.4426950408889634f;

// ---------------------------------------------------------------- smem tile: 64 rows x 64 bf16, 16B-chunk XOR swizzle
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// cp.async a 64x64 bf16 tile: global rows [grow0, grow0+64) of a matrix with row stride ld (elements),
// rows >= nvalid are zero-filled.
__device__ __forceinline__ void tile_load_async(uint32_t smem_tile, const __nv_bfloat16* __restrict__ g, int64_t ld,
                                                int nvalid, int tid) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * kAttnThreads;  // 512 chunks of 16 B
        const int row = idx >> 3, chunk = idx & 7;
        const bool ok = row < nvalid;
        const __nv_bfloat16* src = g + (ok ? (int64_t)row * ld + chunk * 8 : 0);
        ptx::cp_async16(smem_tile + tile_off(row, chunk), src, ok);
    }
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragments of the warp's 16 rows x 64 columns (4 k-steps) from a tile
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[4][4], uint32_t smem_tile, int row0, int lane) {
    const int row = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(smem_tile + tile_off(row, ks * 2 + (lane >> 4)), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
}

// acc[16 x 64] += A[16 x 64] * Tile^T   where Tile is [64 (n)][64 (k)]   (e.g. S = Q K^T with Tile = K)
__device__ __forceinline__ void mma_a_tileT(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t smem_tile, int lane) {
    const int mi = lane >> 3;
#pragma unroll
    for (int np = 0; np < 4; ++np) {  // pairs of n-tiles (16 n)
        const int row = np * 16 + (lane & 7) + (mi >> 1) * 8;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(smem_tile + tile_off(row, ks * 2 + (mi & 1)), b0, b1, b2, b3);
            mma_bf16(acc[2 * np], a[ks], b0, b1);
            mma_bf16(acc[2 * np + 1], a[ks], b2, b3);
        }
    }
}

// acc[16 x 64] += A[16 x 64] * Tile      where Tile is [64 (k)][64 (n)]   (e.g. O = P V with Tile = V)
__device__ __forceinline__ void mma_a_tile(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t smem_tile, int lane) {
    const int mi = lane >> 3;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int row = ks * 16 + (lane & 7) + (mi & 1) * 8;
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(smem_tile + tile_off(row, np * 2 + (mi >> 1)), b0, b1, b2, b3);
            mma_bf16(acc[2 * np], a[ks], b0, b1);
            mma_bf16(acc[2 * np + 1], a[ks], b2, b3);
        }
    }
}

// accumulator tile (16 x 64 fp32, C layout) -> A fragments (bf16) for a following MMA that contracts its columns
__device__ __forceinline__ void acc_to_a(uint32_t (&a)[4][4], const float (&c)[8][4]) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        a[ks][0] = pack_bf16x2(c[2 * ks][0], c[2 * ks][1]);
        a[ks][1] = pack_bf16x2(c[2 * ks][2], c[2 * ks][3]);
        a[ks][2] = pack_bf16x2(c[2 * ks + 1][0], c[2 * ks + 1][1]);
        a[ks][3] = pack_bf16x2(c[2 * ks + 1][2], c[2 * ks + 1][3]);
    }
}

__device__ __forceinline__ void zero_acc(float (&c)[8][4]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
}

struct AttnParams {
    const __nv_bfloat16* qkv;   // [rows, 3H]
    __nv_bfloat16* ctx;         // [rows, H]
    float* lse;                 // [heads, rows]  log2-domain logsumexp of the scaled+biased scores
    const float* keybias;       // [rows] additive mask per key ((1-m) * -10000)
    const int* cu_seqlens;      // [nseq + 1]
    const int* kv_end;          // [nseq] or null (see mmb_attn_args)
    // backward
    const __nv_bfloat16* dctx;  // [rows, H]
    __nv_bfloat16* dqkv;        // [rows, 3H]
    float* dsum;                // [heads, rows]  D = rowsum(dO * O)
    int H, nheads, total_rows;
    float scale_log2;           // d^-1/2 * log2(e)
    float scale;                // d^-1/2
    uint32_t thresh;            // dropout threshold (0 = off)
    float inv_keep;
    uint64_t seed;
    uint32_t rng_stream;
};

// dropout generator row id of query q of (seq, head): the probability row; columns are the keys
// number of leading keys that can carry probability mass (everything behind is masked in whole tiles)
__device__ __forceinline__ int effective_keys(const int* kv_end, int seq, int S) {
    if (kv_end == nullptr) return S;
    const int e = kv_end[seq];
    return (e <= 0 || e > S) ? S : e;
}
__device__ __forceinline__ uint32_t prob_row_id(int seq, int head, int nheads, int S, int q) {
    return ((uint32_t)seq * (uint32_t)nheads + (uint32_t)head) * (uint32_t)S + (uint32_t)q;
}

// ================================================================== forward
__global__ void __launch_bounds__(kAttnThreads)
attn_fwd_kernel(const AttnParams p) {
    __shared__ __align__(128) uint8_t sQ[kTile * 128];
    __shared__ __align__(128) uint8_t sK[2][kTile * 128];
    __shared__ __align__(128) uint8_t sV[2][kTile * 128];
    __shared__ float sBias[2][kTile];

    const int seq = blockIdx.z, head = blockIdx.y, qt = blockIdx.x;
    const int row0 = p.cu_seqlens[seq];
    const int S = p.cu_seqlens[seq + 1] - row0;
    if (qt * kTile >= S) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int64_t ld = 3 * (int64_t)p.H;
    const __nv_bfloat16* Qg = p.qkv + (int64_t)(row0 + qt * kTile) * ld + head * kD;
    const __nv_bfloat16* Kg = p.qkv + (int64_t)row0 * ld + p.H + head * kD;
    const __nv_bfloat16* Vg = Kg + p.H;
    const uint32_t sQa = ptx::smem_u32(sQ), sKa = ptx::smem_u32(sK), sVa = ptx::smem_u32(sV);
    const int nkv = (effective_keys(p.kv_end, seq, S) + kTile - 1) / kTile;

    tile_load_async(sQa, Qg, ld, S - qt * kTile, tid);
    tile_load_async(sKa, Kg, ld, S, tid);
    tile_load_async(sVa, Vg, ld, S, tid);
    if (tid < kTile) sBias[0][tid] = tid < S ? p.keybias[row0 + tid] * kLog2e : -INFINITY;
    ptx::cp_async_commit();

    uint32_t qa[4][4];
    float o[8][4];
    zero_acc(o);
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};

    for (int kt = 0; kt < nkv; ++kt) {
        const int st = kt & 1;
        if (kt + 1 < nkv) {
            const int k0 = (kt + 1) * kTile;
            tile_load_async(sKa + (st ^ 1) * kTile * 128, Kg + (int64_t)k0 * ld, ld, S - k0, tid);
            tile_load_async(sVa + (st ^ 1) * kTile * 128, Vg + (int64_t)k0 * ld, ld, S - k0, tid);
            if (tid < kTile) sBias[st ^ 1][tid] = (k0 + tid) < S ? p.keybias[row0 + k0 + tid] * kLog2e : -INFINITY;
            ptx::cp_async_commit();
            ptx::cp_async_wait<1>();
        } else {
            ptx::cp_async_wait<0>();
        }
        __syncthreads();
        if (kt == 0) load_a_frags(qa, sQa, warp * 16, lane);

        float s[8][4];
        zero_acc(s);
        mma_a_tileT(s, qa, sKa + st * kTile * 128, lane);
        // scale + additive key mask (log2 domain); padded keys are -inf
        float mt[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float b0 = sBias[st][nt * 8 + 2 * t], b1 = sBias[st][nt * 8 + 2 * t + 1];
            s[nt][0] = s[nt][0] * p.scale_log2 + b0;
            s[nt][1] = s[nt][1] * p.scale_log2 + b1;
            s[nt][2] = s[nt][2] * p.scale_log2 + b0;
            s[nt][3] = s[nt][3] * p.scale_log2 + b1;
            mt[0] = fmaxf(mt[0], fmaxf(s[nt][0], s[nt][1]));
            mt[1] = fmaxf(mt[1], fmaxf(s[nt][2], s[nt][3]));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mt[r] = fmaxf(mt[r], __shfl_xor_sync(0xffffffffu, mt[r], 1));
            mt[r] = fmaxf(mt[r], __shfl_xor_sync(0xffffffffu, mt[r], 2));
        }
        float corr[2], lsum[2] = {0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float mn = fmaxf(m[r], mt[r]);
            corr[r] = exp2f(m[r] - mn);
            m[r] = mn;
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = exp2f(s[nt][0] - m[0]);
            s[nt][1] = exp2f(s[nt][1] - m[0]);
            s[nt][2] = exp2f(s[nt][2] - m[1]);
            s[nt][3] = exp2f(s[nt][3] - m[1]);
            lsum[0] += s[nt][0] + s[nt][1];
            lsum[1] += s[nt][2] + s[nt][3];
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l[r] = l[r] * corr[r] + lsum[r];  // per-thread partial; reduced at the end
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            o[nt][0] *= corr[0];
            o[nt][1] *= corr[0];
            o[nt][2] *= corr[1];
            o[nt][3] *= corr[1];
        }
        if (p.thresh != 0u) {
            const int q0 = qt * kTile + warp * 16 + g;
            const uint32_t key_lo = rng_row_key(p.seed, p.rng_stream, prob_row_id(seq, head, p.nheads, S, q0));
            const uint32_t key_hi = rng_row_key(p.seed, p.rng_stream, prob_row_id(seq, head, p.nheads, S, q0 + 8));
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const uint32_t pair = (uint32_t)(kt * kTile + nt * 8 + 2 * t) >> 1;   // this thread's two adjacent keys
                const uint32_t b_lo = rng_pair(key_lo, pair), b_hi = rng_pair(key_hi, pair);
                s[nt][0] = rng_keep_lo(b_lo, p.thresh) ? s[nt][0] * p.inv_keep : 0.f;
                s[nt][1] = rng_keep_hi(b_lo, p.thresh) ? s[nt][1] * p.inv_keep : 0.f;
                s[nt][2] = rng_keep_lo(b_hi, p.thresh) ? s[nt][2] * p.inv_keep : 0.f;
                s[nt][3] = rng_keep_hi(b_hi, p.thresh) ? s[nt][3] * p.inv_keep : 0.f;
            }
        }
        uint32_t pa[4][4];
        acc_to_a(pa, s);
        mma_a_tile(o, pa, sVa + st * kTile * 128, lane);
        __syncthreads();  // everyone done with stage st before it is refilled
    }

#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
    }
    const float inv0 = 1.f / l[0], inv1 = 1.f / l[1];
    const int q_lo = qt * kTile + warp * 16 + g, q_hi = q_lo + 8;
    __nv_bfloat16* O = p.ctx + (int64_t)row0 * p.H + head * kD;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int c = nt * 8 + 2 * t;
        if (q_lo < S) *reinterpret_cast<uint32_t*>(O + (int64_t)q_lo * p.H + c) = pack_bf16x2(o[nt][0] * inv0, o[nt][1] * inv0);
        if (q_hi < S) *reinterpret_cast<uint32_t*>(O + (int64_t)q_hi * p.H + c) = pack_bf16x2(o[nt][2] * inv1, o[nt][3] * inv1);
    }
    if (t == 0 && p.lse != nullptr) {
        float* L = p.lse + (int64_t)head * p.total_rows + row0;
        if (q_lo < S) L[q_lo] = m[0] + log2f(l[0]);
        if (q_hi < S) L[q_hi] = m[1] + log2f(l[1]);
    }
}

// ================================================================== backward preprocess: D = rowsum(dO * O)
__global__ void __launch_bounds__(256)
attn_bwd_dsum_kernel(const __nv_bfloat16* __restrict__ dctx, const __nv_bfloat16* __restrict__ ctx, float* __restrict__ dsum,
                     int total_rows, int H, int nheads) {
    // one 8-lane group per (row, head): 8 lanes x 8 bf16 = 64
    const int gidx = (blockIdx.x * 256 + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    if (gidx >= total_rows * nheads) return;
    const int row = gidx / nheads, head = gidx - row * nheads;
    const int64_t off = (int64_t)row * H + head * kD + sub * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(dctx + off);
    const uint4 b = *reinterpret_cast<const uint4*>(ctx + off);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 x = unpack_bf16x2(aw[i]), y = unpack_bf16x2(bw[i]);
        s += x.x * y.x + x.y * y.y;
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (sub == 0) dsum[(int64_t)head * total_rows + row] = s;
}

// ================================================================== backward dK, dV   (rows = keys)
__global__ void __launch_bounds__(kAttnThreads)
attn_bwd_dkv_kernel(const AttnParams p) {
    extern __shared__ __align__(128) uint8_t dyn_smem[];
    uint8_t* sK = dyn_smem;                       // [64][128]
    uint8_t* sV = sK + kTile * 128;               // [64][128]
    uint8_t* sQ = sV + kTile * 128;               // [2][64][128]
    uint8_t* sdO = sQ + 2 * kTile * 128;          // [2][64][128]
    float (*sL)[kTile] = reinterpret_cast<float (*)[kTile]>(sdO + 2 * kTile * 128);
    float (*sD)[kTile] = sL + 2;
    uint32_t (*sKey)[kTile] = reinterpret_cast<uint32_t (*)[kTile]>(sD + 2);   // dropout row keys of the q tile

    const int seq = blockIdx.z, head = blockIdx.y, kt = blockIdx.x;
    const int row0 = p.cu_seqlens[seq];
    const int S = p.cu_seqlens[seq + 1] - row0;
    if (kt * kTile >= S) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int64_t ld = 3 * (int64_t)p.H;
    if (kt * kTile >= effective_keys(p.kv_end, seq, S)) {
        // every key of this tile is masked: P == 0 exactly, so dK = dV = 0
        __nv_bfloat16* dKz = p.dqkv + (int64_t)(row0 + kt * kTile) * ld + p.H + head * kD;
        const int nrows = min(kTile, S - kt * kTile);
        for (int i = tid; i < nrows * 16; i += kAttnThreads) {      // 16 x 16 B = dK row (128 B) + dV row (128 B)
            const int r = i >> 4, c = i & 15;
            *reinterpret_cast<uint4*>(dKz + (int64_t)r * ld + (c >> 3) * p.H + (c & 7) * 8) = make_uint4(0, 0, 0, 0);
        }
        return;
    }
    const __nv_bfloat16* Qg = p.qkv + (int64_t)row0 * ld + head * kD;
    const __nv_bfloat16* Kg = Qg + p.H + (int64_t)kt * kTile * ld;
    const __nv_bfloat16* Vg = Kg + p.H;
    const __nv_bfloat16* dOg = p.dctx + (int64_t)row0 * p.H + head * kD;
    const float* Lg = p.lse + (int64_t)head * p.total_rows + row0;
    const float* Dg = p.dsum + (int64_t)head * p.total_rows + row0;
    const uint32_t sKa = ptx::smem_u32(sK), sVa = ptx::smem_u32(sV), sQa = ptx::smem_u32(sQ), sdOa = ptx::smem_u32(sdO);
    const int nq = (S + kTile - 1) / kTile;

    tile_load_async(sKa, Kg, ld, S - kt * kTile, tid);
    tile_load_async(sVa, Vg, ld, S - kt * kTile, tid);
    tile_load_async(sQa, Qg, ld, S, tid);
    tile_load_async(sdOa, dOg, p.H, S, tid);
    if (tid < kTile) {
        sL[0][tid] = tid < S ? Lg[tid] : INFINITY;  // padded queries: P = exp2(-inf) = 0
        sD[0][tid] = tid < S ? Dg[tid] : 0.f;
        if (p.thresh != 0u) sKey[0][tid] = rng_row_key(p.seed, p.rng_stream, prob_row_id(seq, head, p.nheads, S, tid));
    }
    ptx::cp_async_commit();

    // additive bias of this warp's two key rows
    const int k_lo = kt * kTile + warp * 16 + g, k_hi = k_lo + 8;
    const float bias_lo = k_lo < S ? p.keybias[row0 + k_lo] * kLog2e : -INFINITY;
    const float bias_hi = k_hi < S ? p.keybias[row0 + k_hi] * kLog2e : -INFINITY;

    uint32_t ka[4][4], va[4][4];
    float dk[8][4], dv[8][4];
    zero_acc(dk);
    zero_acc(dv);

    for (int qt = 0; qt < nq; ++qt) {
        const int st = qt & 1;
        if (qt + 1 < nq) {
            const int q0 = (qt + 1) * kTile;
            tile_load_async(sQa + (st ^ 1) * kTile * 128, Qg + (int64_t)q0 * ld, ld, S - q0, tid);
            tile_load_async(sdOa + (st ^ 1) * kTile * 128, dOg + (int64_t)q0 * p.H, p.H, S - q0, tid);
            if (tid < kTile) {
                sL[st ^ 1][tid] = (q0 + tid) < S ? Lg[q0 + tid] : INFINITY;
                sD[st ^ 1][tid] = (q0 + tid) < S ? Dg[q0 + tid] : 0.f;
                if (p.thresh != 0u)
                    sKey[st ^ 1][tid] = rng_row_key(p.seed, p.rng_stream, prob_row_id(seq, head, p.nheads, S, q0 + tid));
            }
            ptx::cp_async_commit();
            ptx::cp_async_wait<1>();
        } else {
            ptx::cp_async_wait<0>();
        }
        __syncthreads();
        if (qt == 0) {
            load_a_frags(ka, sKa, warp * 16, lane);
            load_a_frags(va, sVa, warp * 16, lane);
        }
        // S^T[key][q] and dP^T[key][q]
        float s[8][4], dp[8][4];
        zero_acc(s);
        zero_acc(dp);
        mma_a_tileT(s, ka, sQa + st * kTile * 128, lane);
        mma_a_tileT(dp, va, sdOa + st * kTile * 128, lane);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int qc = nt * 8 + 2 * t;
            const float L0 = sL[st][qc], L1 = sL[st][qc + 1];
            const float D0 = sD[st][qc], D1 = sD[st][qc + 1];
            float pr[4];
            pr[0] = exp2f(s[nt][0] * p.scale_log2 + bias_lo - L0);
            pr[1] = exp2f(s[nt][1] * p.scale_log2 + bias_lo - L1);
            pr[2] = exp2f(s[nt][2] * p.scale_log2 + bias_hi - L0);
            pr[3] = exp2f(s[nt][3] * p.scale_log2 + bias_hi - L1);
            float pd[4] = {pr[0], pr[1], pr[2], pr[3]};   // dropped probabilities (feed dV)
            float dpj[4] = {dp[nt][0], dp[nt][1], dp[nt][2], dp[nt][3]};
            if (p.thresh != 0u) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = (j >> 1) ? k_hi : k_lo;
                    const bool keep = rng_keep_col(sKey[st][qc + (j & 1)], (uint32_t)k, p.thresh);
                    pd[j] = keep ? pr[j] * p.inv_keep : 0.f;
                    dpj[j] = keep ? dpj[j] * p.inv_keep : 0.f;
                }
            }
            // dS^T = P * (dP - D[q])
            dp[nt][0] = pr[0] * (dpj[0] - D0);
            dp[nt][1] = pr[1] * (dpj[1] - D1);
            dp[nt][2] = pr[2] * (dpj[2] - D0);
            dp[nt][3] = pr[3] * (dpj[3] - D1);
            s[nt][0] = pd[0];
            s[nt][1] = pd[1];
            s[nt][2] = pd[2];
            s[nt][3] = pd[3];
        }
        uint32_t fa[4][4];
        acc_to_a(fa, s);
        mma_a_tile(dv, fa, sdOa + st * kTile * 128, lane);   // dV += P^T dO
        acc_to_a(fa, dp);
        mma_a_tile(dk, fa, sQa + st * kTile * 128, lane);    // dK += dS^T Q
        __syncthreads();
    }
    __nv_bfloat16* dK = p.dqkv + (int64_t)row0 * ld + p.H + head * kD;
    __nv_bfloat16* dV = dK + p.H;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int c = nt * 8 + 2 * t;
        if (k_lo < S) {
            *reinterpret_cast<uint32_t*>(dK + (int64_t)k_lo * ld + c) = pack_bf16x2(dk[nt][0] * p.scale, dk[nt][1] * p.scale);
            *reinterpret_cast<uint32_t*>(dV + (int64_t)k_lo * ld + c) = pack_bf16x2(dv[nt][0], dv[nt][1]);
        }
        if (k_hi < S) {
            *reinterpret_cast<uint32_t*>(dK + (int64_t)k_hi * ld + c) = pack_bf16x2(dk[nt][2] * p.scale, dk[nt][3] * p.scale);
            *reinterpret_cast<uint32_t*>(dV + (int64_t)k_hi * ld + c) = pack_bf16x2(dv[nt][2], dv[nt][3]);
        }
    }
}

// ================================================================== backward dQ   (rows = queries)
__global__ void __launch_bounds__(kAttnThreads)
attn_bwd_dq_kernel(const AttnParams p) {
    extern __shared__ __align__(128) uint8_t dyn_smem[];
    uint8_t* sQ = dyn_smem;                       // [64][128]
    uint8_t* sdO = sQ + kTile * 128;              // [64][128]
    uint8_t* sK = sdO + kTile * 128;              // [2][64][128]
    uint8_t* sV = sK + 2 * kTile * 128;           // [2][64][128]
    float (*sBias)[kTile] = reinterpret_cast<float (*)[kTile]>(sV + 2 * kTile * 128);

    const int seq = blockIdx.z, head = blockIdx.y, qt = blockIdx.x;
    const int row0 = p.cu_seqlens[seq];
    const int S = p.cu_seqlens[seq + 1] - row0;
    if (qt * kTile >= S) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int64_t ld = 3 * (int64_t)p.H;
    const __nv_bfloat16* Qg = p.qkv + (int64_t)(row0 + qt * kTile) * ld + head * kD;
    const __nv_bfloat16* Kg = p.qkv + (int64_t)row0 * ld + p.H + head * kD;
    const __nv_bfloat16* Vg = Kg + p.H;
    const __nv_bfloat16* dOg = p.dctx + (int64_t)(row0 + qt * kTile) * p.H + head * kD;
    const uint32_t sQa = ptx::smem_u32(sQ), sdOa = ptx::smem_u32(sdO), sKa = ptx::smem_u32(sK), sVa = ptx::smem_u32(sV);
    const int nkv = (effective_keys(p.kv_end, seq, S) + kTile - 1) / kTile;

    tile_load_async(sQa, Qg, ld, S - qt * kTile, tid);
    tile_load_async(sdOa, dOg, p.H, S - qt * kTile, tid);
    tile_load_async(sKa, Kg, ld, S, tid);
    tile_load_async(sVa, Vg, ld, S, tid);
    if (tid < kTile) sBias[0][tid] = tid < S ? p.keybias[row0 + tid] * kLog2e : -INFINITY;
    ptx::cp_async_commit();

    const int q_lo = qt * kTile + warp * 16 + g, q_hi = q_lo + 8;
    const float* Lg = p.lse + (int64_t)head * p.total_rows + row0;
    const float* Dg = p.dsum + (int64_t)head * p.total_rows + row0;
    const float L_lo = q_lo < S ? Lg[q_lo] : INFINITY, L_hi = q_hi < S ? Lg[q_hi] : INFINITY;
    const float D_lo = q_lo < S ? Dg[q_lo] : 0.f, D_hi = q_hi < S ? Dg[q_hi] : 0.f;
    const uint32_t rkey_lo = rng_row_key(p.seed, p.rng_stream, prob_row_id(seq, head, p.nheads, S, q_lo));
    const uint32_t rkey_hi = rng_row_key(p.seed, p.rng_stream, prob_row_id(seq, head, p.nheads, S, q_hi));

    uint32_t qa[4][4], doa[4][4];
    float dq[8][4];
    zero_acc(dq);

    for (int kt = 0; kt < nkv; ++kt) {
        const int st = kt & 1;
        if (kt + 1 < nkv) {
            const int k0 = (kt + 1) * kTile;
            tile_load_async(sKa + (st ^ 1) * kTile * 128, Kg + (int64_t)k0 * ld, ld, S - k0, tid);
            tile_load_async(sVa + (st ^ 1) * kTile * 128, Vg + (int64_t)k0 * ld, ld, S - k0, tid);
            if (tid < kTile) sBias[st ^ 1][tid] = (k0 + tid) < S ? p.keybias[row0 + k0 + tid] * kLog2e : -INFINITY;
            ptx::cp_async_commit();
            ptx::cp_async_wait<1>();
        } else {
            ptx::cp_async_wait<0>();
        }
        __syncthreads();
        if (kt == 0) {
            load_a_frags(qa, sQa, warp * 16, lane);
            load_a_frags(doa, sdOa, warp * 16, lane);
        }
        float s[8][4], dp[8][4];
        zero_acc(s);
        zero_acc(dp);
        mma_a_tileT(s, qa, sKa + st * kTile * 128, lane);    // S = Q K^T
        mma_a_tileT(dp, doa, sVa + st * kTile * 128, lane);  // dP = dO V^T
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int kc = nt * 8 + 2 * t;
            const float b0 = sBias[st][kc], b1 = sBias[st][kc + 1];
            float pr[4];
            pr[0] = exp2f(s[nt][0] * p.scale_log2 + b0 - L_lo);
            pr[1] = exp2f(s[nt][1] * p.scale_log2 + b1 - L_lo);
            pr[2] = exp2f(s[nt][2] * p.scale_log2 + b0 - L_hi);
            pr[3] = exp2f(s[nt][3] * p.scale_log2 + b1 - L_hi);
            float dpj[4] = {dp[nt][0], dp[nt][1], dp[nt][2], dp[nt][3]};
            if (p.thresh != 0u) {
                const uint32_t pair = (uint32_t)(kt * kTile + kc) >> 1;
                const uint32_t b_lo = rng_pair(rkey_lo, pair), b_hi = rng_pair(rkey_hi, pair);
                dpj[0] = rng_keep_lo(b_lo, p.thresh) ? dpj[0] * p.inv_keep : 0.f;
                dpj[1] = rng_keep_hi(b_lo, p.thresh) ? dpj[1] * p.inv_keep : 0.f;
                dpj[2] = rng_keep_lo(b_hi, p.thresh) ? dpj[2] * p.inv_keep : 0.f;
                dpj[3] = rng_keep_hi(b_hi, p.thresh) ? dpj[3] * p.inv_keep : 0.f;
            }
            dp[nt][0] = pr[0] * (dpj[0] - D_lo);
            dp[nt][1] = pr[1] * (dpj[1] - D_lo);
            dp[nt][2] = pr[2] * (dpj[2] - D_hi);
            dp[nt][3] = pr[3] * (dpj[3] - D_hi);
        }
        uint32_t fa[4][4];
        acc_to_a(fa, dp);
        mma_a_tile(dq, fa, sKa + st * kTile * 128, lane);    // dQ += dS K
        __syncthreads();
    }
    __nv_bfloat16* dQ = p.dqkv + (int64_t)row0 * ld + head * kD;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int c = nt * 8 + 2 * t;
        if (q_lo < S) *reinterpret_cast<uint32_t*>(dQ + (int64_t)q_lo * ld + c) = pack_bf16x2(dq[nt][0] * p.scale, dq[nt][1] * p.scale);
        if (q_hi < S) *reinterpret_cast<uint32_t*>(dQ + (int64_t)q_hi * ld + c) = pack_bf16x2(dq[nt][2] * p.scale, dq[nt][3] * p.scale);
    }
}

static int fill_params(AttnParams& p, const mmb_attn_args* a) {
    MMB_REQUIRE(a && a->qkv && a->keybias && a->cu_seqlens, "attn: null pointer");
    MMB_REQUIRE(a->H > 0 && a->nheads > 0 && a->H == a->nheads * kD, "attn: head dim must be 64 (H=%d heads=%d)", a->H,
                a->nheads);
    MMB_REQUIRE(a->nseq > 0 && a->max_seqlen > 0 && a->total_rows > 0, "attn: empty batch");
    p.qkv = (const __nv_bfloat16*)a->qkv;
    p.ctx = (__nv_bfloat16*)a->ctx;
    p.lse = a->lse;
    p.keybias = a->keybias;
    p.cu_seqlens = a->cu_seqlens;
    p.kv_end = a->kv_end;
    p.dctx = (const __nv_bfloat16*)a->dctx;
    p.dqkv = (__nv_bfloat16*)a->dqkv;
    p.dsum = a->dsum;
    p.H = a->H;
    p.nheads = a->nheads;
    p.total_rows = a->total_rows;
    p.scale = 1.0f / sqrtf((float)kD);
    p.scale_log2 = p.scale * kLog2e;
    p.thresh = dropout_threshold(a->p_drop);
    p.inv_keep = dropout_inv_keep(a->p_drop);
    p.seed = a->seed;
    p.rng_stream = a->rng_stream;
    return MMB_OK;
}

int launch_attn_fwd_tc(const mmb_attn_args* a, cudaStream_t stream);   // attn_tc.cu
int launch_attn_bwd_tc(const mmb_attn_args* a, cudaStream_t stream);   // attn_bwd_tc.cu

}  // namespace mmb

using namespace mmb;

extern "C" int mmb_attn_fwd(const mmb_attn_args* a, void* stream) {
    AttnParams p;
    int rc = fill_params(p, a);
    if (rc != MMB_OK) return rc;
    MMB_REQUIRE(a->ctx != nullptr, "attn_fwd: null ctx");
    // tcgen05 / TMEM kernel by default; flags bit 0 selects the legacy mma.sync kernel (kept for A/B checks)
    if (!(a->flags & 1)) return launch_attn_fwd_tc(a, (cudaStream_t)stream);
    dim3 grid((a->max_seqlen + kTile - 1) / kTile, a->nheads, a->nseq);
    attn_fwd_kernel<<<grid, kAttnThreads, 0, (cudaStream_t)stream>>>(p);
    return check_launch("attn_fwd_kernel");
}

extern "C" int mmb_attn_bwd(const mmb_attn_args* a, void* stream) {
    AttnParams p;
    int rc = fill_params(p, a);
    if (rc != MMB_OK) return rc;
    MMB_REQUIRE(a->ctx && a->lse && a->dctx && a->dqkv && a->dsum, "attn_bwd: null pointer");
    const int groups = a->total_rows * a->nheads;
    attn_bwd_dsum_kernel<<<(groups * 8 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p.dctx, p.ctx, p.dsum, a->total_rows,
                                                                                     a->H, a->nheads);
    rc = check_launch("attn_bwd_dsum_kernel");
    if (rc != MMB_OK) return rc;
    if (!(a->flags & 1)) return launch_attn_bwd_tc(a, (cudaStream_t)stream);
    dim3 grid((a->max_seqlen + kTile - 1) / kTile, a->nheads, a->nseq);
    constexpr int kBwdSmem = 6 * kTile * 128 + 6 * kTile * (int)sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        MMB_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
        MMB_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
        attr_set = true;
    }
    attn_bwd_dkv_kernel<<<grid, kAttnThreads, kBwdSmem, (cudaStream_t)stream>>>(p);
    rc = check_launch("attn_bwd_dkv_kernel");
    if (rc != MMB_OK) return rc;
    attn_bwd_dq_kernel<<<grid, kAttnThreads, kBwdSmem, (cudaStream_t)stream>>>(p);
    return check_launch("attn_bwd_dq_kernel");
}
