// Masked self-attention BACKWARD on 5th-generation tensor cores (tcgen05 + TMEM + TMA), head dim 64.
// Autograd of modeling_bert.py:115-140 (eager_attention_forward) for the packed batch; same contract as
// mmb_attn_bwd in include/mmbert_sm100.h (dQ|dK|dV written into dqkv [rows, 3H], D = rowsum(dO ∘ O) precomputed).
//
// Two launches of ONE persistent, warp-specialised kernel template (no transposes, no atomics):
//   dQ  pass  rows (TMEM lanes) = 128 queries, steps of 64 keys:
//        S  = Q K^T,  dP  = dO V^T   -> TMEM;  dS  = P ∘ (dP - D) -> bf16 -> swizzled smem;  dQ += dS K
//   dKV pass  rows (TMEM lanes) = 128 keys,    steps of 64 queries:
//        S^T = K Q^T, dP^T = V dO^T  -> TMEM;  P^T, dS^T -> smem;  dV += P^T dO,  dK += dS^T Q
// Both are "row operands (A0, A1) x step operands (B0, B1)":  scores = A0 B0^T and A1 B1^T, then accumulate with
// the staged tiles against B0 / B1 read in place as MN-major operands.
//
// One CTA per SM (all 512 TMEM columns, ~165 / 197 KB shared memory), 18 warps:
//   warp 16   TMA producer: row operands (double-buffered across work items), step operands (4-stage ring), and the
//             per-column vectors (key bias / -LSE / -D / dropout keys) of each stage
//   warp 17   one thread issues every tcgen05.mma: the scores of step g+1 are issued BEFORE the accumulating MMAs
//             of step g, so the tensor pipe computes the next scores while the compute warps work on this step
//   warps 0-15 element-wise stage: every thread owns one TMEM lane (row) and 16 of the step's 64 columns; the
//             probabilities are rebuilt from the forward's log-sum-exp, ~7 instructions per probability (packed
//             FFMA2 / FADD2 / FMUL2, one MUFU.EX2, one-IMAD dropout decision — common.cuh: attn_keep)
// Rings: scores 2 slots x (64 + 64) TMEM columns, staged tiles 2 buffers, accumulators 2 buffers (so the epilogue
// of one work item overlaps the next item's MMAs).  The CTA walks work items (sequence, head, 128-row tile) with a
// grid stride; every role evaluates the same item list, and all ring positions are carried across items.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace mmb {

int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1);

namespace {

constexpr int kRows = 128;           // TMEM lanes per work item
constexpr int kStep = 64;            // columns per step
constexpr int kHd = 64;              // head dim
constexpr int kComputeWarps = 16;
constexpr int kProducerWarp = 16, kScoreWarp = 17, kAccWarp = 18;
constexpr int kWsThreads = 19 * 32;
constexpr int kStages = 4;           // step-operand ring
constexpr float kLog2eB = 1.4426950408889634f;
constexpr int kBig = kRows * kHd * 2;     // 16 KB: 128-row operand tile (128 B rows, SWIZZLE_128B)
constexpr int kSmall = kStep * kHd * 2;   // 8 KB: 64-row operand tile

// shared-memory map (offsets from a 1024-byte aligned base)
constexpr int kOffRow = 0;                               // [2 buffers][A0 | A1] x 16 KB
constexpr int kOffStep = 4 * kBig;                       // [kStages][B0 | B1] x 8 KB
constexpr int kOffStaged = kOffStep + kStages * 2 * kSmall;
template <bool kIsDq> constexpr int staged_bytes() { return kIsDq ? kBig : 2 * kBig; }     // dS  |  P^T, dS^T
template <bool kIsDq> constexpr int off_cols() { return kOffStaged + 2 * staged_bytes<kIsDq>(); }
constexpr int kColBytes = 3 * kStep * 4;                 // per stage: add[64] f32, negD[64] f32, key[64] u32
constexpr int kRowcBytes = 3 * kRows * 4;                // per row buffer: add[128] f32, negD[128] f32, key[128] u32
template <bool kIsDq> constexpr int off_rowc() { return off_cols<kIsDq>() + kStages * kColBytes; }
template <bool kIsDq> constexpr int off_bars() { return off_rowc<kIsDq>() + 2 * kRowcBytes; }
template <bool kIsDq> constexpr int smem_bytes() { return off_bars<kIsDq>() + 256 + 1024; }
// barrier slots (8 bytes each)
enum { B_ROW_FULL = 0, B_ROW_EMPTY = 2, B_STEP_FULL = 4, B_STEP_EMPTY = 8, B_SC_FULL = 12, B_SC_EMPTY = 14,
       B_ST_FULL = 16, B_ST_EMPTY = 18, B_ACC_FULL = 20, B_ACC_EMPTY = 22, B_COUNT = 24 };

struct BwdTcParams {
    __nv_bfloat16* dqkv;
    const float* lse;
    const float* dsum;
    const float* keybias;
    const int* cu_seqlens;
    const int* kv_end;
    int H, nheads, nseq, tiles, total_rows;
    float scale_log2, scale;
    uint32_t thresh32;
    float inv_keep;
    uint64_t seed;
    uint32_t rng_stream;
};

__device__ __forceinline__ void sts128_b(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts32_b(uint32_t addr, uint32_t x) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(x) : "memory");
}
__device__ __forceinline__ uint32_t lds32_b(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
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
// D[128 x 64] (+)= A[128 x 64] * B^T,  A and B both K-major tiles with 128-byte rows (contraction over the 64 columns).
// Operand addresses are shared-memory byte addresses (< 256 KB, 1024-byte aligned tiles): one 32-bit add per descriptor.
__device__ __forceinline__ void mma_kk(uint32_t tmem_d, uint32_t sA, uint32_t sB, bool accumulate) {
    const uint32_t a16 = sA >> 4, b16 = sB >> 4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        ptx::umma_bf16(tmem_d, ptx::umma_desc_from_halves(a16 + 2 * k, ptx::kDescHiSw128),
                       ptx::umma_desc_from_halves(b16 + 2 * k, ptx::kDescHiSw128), idesc_m128_n64(false),
                       (accumulate || k > 0) ? 1u : 0u);
}
// D[128 x 64] (+)= A[128 x 64] * B,  A K-major, B = [64 (contraction rows)][64] read in place as an MN-major operand
__device__ __forceinline__ void mma_kmn(uint32_t tmem_d, uint32_t sA, uint32_t sB, bool accumulate) {
    const uint32_t a16 = sA >> 4, b16 = (sB >> 4) | ((8192u >> 4) << 16);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        ptx::umma_bf16(tmem_d, ptx::umma_desc_from_halves(a16 + 2 * k, ptx::kDescHiSw128),
                       ptx::umma_desc_from_halves(b16 + 128 * k, ptx::kDescHiSw128), idesc_m128_n64(true),
                       (accumulate || k > 0) ? 1u : 0u);
}

struct Item {
    int seq, head, tile, row0, S, nsteps;   // nsteps == 0: nothing to run (invalid tile, or a fully masked key tile)
    bool valid;
};
template <bool kIsDq>
__device__ __forceinline__ Item get_item(const BwdTcParams& p, int idx, int total) {
    Item it;
    if (idx >= total) {
        it.seq = it.head = it.tile = it.row0 = it.S = it.nsteps = 0;
        it.valid = false;
        return it;
    }
    it.tile = idx % p.tiles;
    const int sh = idx / p.tiles;
    it.head = sh % p.nheads;
    it.seq = sh / p.nheads;
    it.row0 = p.cu_seqlens[it.seq];
    it.S = p.cu_seqlens[it.seq + 1] - it.row0;
    it.valid = it.tile * kRows < it.S;
    int eff = it.S;                         // keys at index >= eff are all masked: P == 0 exactly (see mmb_attn_args)
    if (p.kv_end != nullptr) {
        const int e = p.kv_end[it.seq];
        if (e > 0 && e < it.S) eff = e;
    }
    if (!it.valid) it.nsteps = 0;
    else if (kIsDq) it.nsteps = (eff + kStep - 1) / kStep;
    else it.nsteps = (it.tile * kRows >= eff) ? 0 : (it.S + kStep - 1) / kStep;
    return it;
}

template <bool kIsDq, bool kDrop>
__global__ void __launch_bounds__(kWsThreads, 1)
attn_bwd_ws_kernel(const __grid_constant__ CUtensorMap tm_qkv128, const __grid_constant__ CUtensorMap tm_qkv64,
                   const __grid_constant__ CUtensorMap tm_do, const BwdTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = sbase + off_bars<kIsDq>();
    const uint32_t tmem_slot = bars + B_COUNT * 8;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto bar = [&](int slot) { return bars + (uint32_t)slot * 8u; };
    auto row_tile = [&](int buf, int which) { return sbase + kOffRow + (uint32_t)(buf * 2 + which) * kBig; };
    auto step_tile = [&](int st, int which) { return sbase + kOffStep + (uint32_t)(st * 2 + which) * kSmall; };
    auto staged_tile = [&](int buf, int which) {
        return sbase + kOffStaged + (uint32_t)buf * staged_bytes<kIsDq>() + (uint32_t)which * kBig;
    };
    auto cols = [&](int st) { return sbase + off_cols<kIsDq>() + (uint32_t)st * kColBytes; };
    auto rowc = [&](int buf) { return sbase + off_rowc<kIsDq>() + (uint32_t)buf * kRowcBytes; };

    if (tid == 0) {
        ptx::prefetch_tensormap(&tm_qkv128);
        ptx::prefetch_tensormap(&tm_qkv64);
        ptx::prefetch_tensormap(&tm_do);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(bar(B_ROW_FULL + i), 2);                    // expect_tx arrive + row-constant arrive
            ptx::mbar_init(bar(B_ROW_EMPTY + i), 1 + kComputeWarps);   // last scores MMA + row constants read
            ptx::mbar_init(bar(B_SC_FULL + i), 1);
            ptx::mbar_init(bar(B_SC_EMPTY + i), kComputeWarps);
            ptx::mbar_init(bar(B_ST_FULL + i), kComputeWarps);
            ptx::mbar_init(bar(B_ST_EMPTY + i), 1);
            ptx::mbar_init(bar(B_ACC_FULL + i), 1);
            ptx::mbar_init(bar(B_ACC_EMPTY + i), kComputeWarps);
        }
        for (int i = 0; i < kStages; ++i) {
            ptx::mbar_init(bar(B_STEP_FULL + i), 2);                   // expect_tx arrive + column-vector arrive
            ptx::mbar_init(bar(B_STEP_EMPTY + i), 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == kScoreWarp) ptx::tmem_alloc<512>(tmem_slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = lds32_b(tmem_slot);
    const int total_items = p.tiles * p.nheads * p.nseq;
    const int stride = gridDim.x;
    // TMEM columns: scores slot s at s * 128 (S | dP), accumulators from 256
    constexpr int kAccCols = kIsDq ? 64 : 128;

    if (warp == kProducerWarp) {
        // ================================================================ TMA producer
        uint32_t g = 0, n = 0;
        Item it = get_item<kIsDq>(p, blockIdx.x, total_items);
        for (int idx = blockIdx.x; idx < total_items; idx += stride) {
            const Item cur = it;
            it = get_item<kIsDq>(p, idx + stride, total_items);          // next item's metadata loads fly under this item
            if (cur.nsteps == 0) continue;
            const int rb = n & 1;
            const int col_q = cur.head * kHd, col_k = p.H + cur.head * kHd, col_v = 2 * p.H + cur.head * kHd;
            const uint32_t prob_base = ((uint32_t)cur.seq * (uint32_t)p.nheads + (uint32_t)cur.head) * (uint32_t)cur.S;
            const int64_t hrow = (int64_t)cur.head * p.total_rows + cur.row0;
            // column values of step s for this lane's two columns (global loads; stored one step later)
            float va[2], vd[2];
            auto load_cols = [&](int s) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = s * kStep + lane + 32 * h;
                    if (kIsDq) {
                        va[h] = c < cur.S ? p.keybias[cur.row0 + c] * kLog2eB : -INFINITY;   // keys beyond the sequence
                        vd[h] = 0.f;
                    } else {
                        va[h] = c < cur.S ? -p.lse[hrow + c] : -INFINITY;                    // padded queries: P = 0
                        vd[h] = c < cur.S ? -p.dsum[hrow + c] : 0.f;
                    }
                }
            };
            load_cols(0);
            // ---- row operands + row constants
            ptx::mbar_wait(bar(B_ROW_EMPTY + rb), ((n >> 1) & 1) ^ 1);
            if (lane == 0) {
                ptx::mbar_expect_tx(bar(B_ROW_FULL + rb), 2 * kBig);
                const int r0 = cur.row0 + cur.tile * kRows;
                if (kIsDq) {
                    ptx::tma_load_2d(row_tile(rb, 0), &tm_qkv128, bar(B_ROW_FULL + rb), col_q, r0);       // Q
                    ptx::tma_load_2d(row_tile(rb, 1), &tm_do, bar(B_ROW_FULL + rb), cur.head * kHd, r0);  // dO
                } else {
                    ptx::tma_load_2d(row_tile(rb, 0), &tm_qkv128, bar(B_ROW_FULL + rb), col_k, r0);       // K
                    ptx::tma_load_2d(row_tile(rb, 1), &tm_qkv128, bar(B_ROW_FULL + rb), col_v, r0);       // V
                }
            }
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int i = lane + 32 * h;
                const int rr = cur.tile * kRows + i;           // query (dQ pass) / key (dKV pass)
                const uint32_t a = rowc(rb) + (uint32_t)i * 4;
                if (kIsDq) {
                    sts32_b(a, __float_as_uint(rr < cur.S ? -p.lse[hrow + rr] : -INFINITY));             // padded rows: P = 0
                    sts32_b(a + kRows * 4, __float_as_uint(rr < cur.S ? -p.dsum[hrow + rr] : 0.f));
                    if (kDrop) sts32_b(a + 2 * kRows * 4, attn_drop_qkey(p.seed, p.rng_stream, prob_base + (uint32_t)rr));
                } else {
                    sts32_b(a, __float_as_uint(rr < cur.S ? p.keybias[cur.row0 + rr] * kLog2eB : -INFINITY));
                    if (kDrop) sts32_b(a + 2 * kRows * 4, attn_drop_kkey(p.seed, p.rng_stream, prob_base + (uint32_t)rr));
                }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar(B_ROW_FULL + rb));
            // ---- step operands + column vectors
            for (int s = 0; s < cur.nsteps; ++s, ++g) {
                const int st = g % kStages;
                ptx::mbar_wait(bar(B_STEP_EMPTY + st), ((g / kStages) & 1) ^ 1);
                if (lane == 0) {
                    ptx::mbar_expect_tx(bar(B_STEP_FULL + st), 2 * kSmall);
                    const int r0 = cur.row0 + s * kStep;
                    if (kIsDq) {
                        ptx::tma_load_2d(step_tile(st, 0), &tm_qkv64, bar(B_STEP_FULL + st), col_k, r0);       // K
                        ptx::tma_load_2d(step_tile(st, 1), &tm_qkv64, bar(B_STEP_FULL + st), col_v, r0);       // V
                    } else {
                        ptx::tma_load_2d(step_tile(st, 0), &tm_qkv64, bar(B_STEP_FULL + st), col_q, r0);       // Q
                        ptx::tma_load_2d(step_tile(st, 1), &tm_do, bar(B_STEP_FULL + st), cur.head * kHd, r0); // dO
                    }
                }
                const float sa[2] = {va[0], va[1]}, sd[2] = {vd[0], vd[1]};
                if (s + 1 < cur.nsteps) load_cols(s + 1);      // next step's global loads fly under this step's stores
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = lane + 32 * h;
                    const int c = s * kStep + i;
                    const uint32_t a = cols(st) + (uint32_t)i * 4;
                    sts32_b(a, __float_as_uint(sa[h]));
                    if (!kIsDq) sts32_b(a + kStep * 4, __float_as_uint(sd[h]));
                    if (kDrop)
                        sts32_b(a + 2 * kStep * 4, kIsDq ? attn_drop_kkey(p.seed, p.rng_stream, prob_base + (uint32_t)c)
                                                         : attn_drop_qkey(p.seed, p.rng_stream, prob_base + (uint32_t)c));
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar(B_STEP_FULL + st));
            }
            ++n;
        }
    } else if (warp == kScoreWarp) {
        // ================================================================ score MMAs: S = A0 B0^T, dP = A1 B1^T
        uint32_t g = 0, n = 0;
        Item it = get_item<kIsDq>(p, blockIdx.x, total_items);
        for (int idx = blockIdx.x; idx < total_items; idx += stride) {
            const Item cur = it;
            it = get_item<kIsDq>(p, idx + stride, total_items);
            if (cur.nsteps == 0) continue;
            const int rb = n & 1;
            ptx::mbar_wait(bar(B_ROW_FULL + rb), (n >> 1) & 1);
            for (int s = 0; s < cur.nsteps; ++s, ++g) {
                const int st = g % kStages, sl = g & 1;
                ptx::mbar_wait(bar(B_STEP_FULL + st), (g / kStages) & 1);
                ptx::mbar_wait(bar(B_SC_EMPTY + sl), ((g >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    mma_kk(tmem + sl * 128, row_tile(rb, 0), step_tile(st, 0), false);
                    mma_kk(tmem + sl * 128 + 64, row_tile(rb, 1), step_tile(st, 1), false);
                    ptx::umma_commit(bar(B_SC_FULL + sl));
                    if (s == cur.nsteps - 1) ptx::umma_commit(bar(B_ROW_EMPTY + rb));   // row operands no longer read
                }
                __syncwarp();
            }
            ++n;
        }
    } else if (warp == kAccWarp) {
        // ================================================================ accumulating MMAs
        uint32_t g = 0, n = 0;
        Item it = get_item<kIsDq>(p, blockIdx.x, total_items);
        for (int idx = blockIdx.x; idx < total_items; idx += stride) {
            const Item cur = it;
            it = get_item<kIsDq>(p, idx + stride, total_items);
            if (cur.nsteps == 0) continue;
            const int a = n & 1;
            ptx::mbar_wait(bar(B_ACC_EMPTY + a), ((n >> 1) & 1) ^ 1);
            const uint32_t acc = tmem + 256 + (uint32_t)a * kAccCols;
            for (int s = 0; s < cur.nsteps; ++s, ++g) {
                const int st = g % kStages, b = g & 1;
                ptx::mbar_wait(bar(B_STEP_FULL + st), (g / kStages) & 1);     // (long complete) acquire the TMA tiles
                ptx::mbar_wait(bar(B_ST_FULL + b), (g >> 1) & 1);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    if (kIsDq) {
                        mma_kmn(acc, staged_tile(b, 0), step_tile(st, 0), s > 0);          // dQ += dS   K
                    } else {
                        mma_kmn(acc, staged_tile(b, 0), step_tile(st, 1), s > 0);          // dV += P^T  dO
                        mma_kmn(acc + 64, staged_tile(b, 1), step_tile(st, 0), s > 0);     // dK += dS^T Q
                    }
                    ptx::umma_commit(bar(B_ST_EMPTY + b));
                    ptx::umma_commit(bar(B_STEP_EMPTY + st));
                    if (s == cur.nsteps - 1) ptx::umma_commit(bar(B_ACC_FULL + a));
                }
                __syncwarp();
            }
            ++n;
        }
    } else {
        // ================================================================ element-wise stage (16 warps)
        const int r = (warp & 3) * 32 + lane;          // row of the work item's tile = TMEM lane
        const int cq = warp >> 2;                      // which 16 of the step's 64 columns
        const uint32_t lane_bits = (uint32_t)((warp & 3) * 32) << 16;
        const int r7 = r & 7;
        const int64_t ld = 3 * (int64_t)p.H;
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), ik2 = make_float2(p.inv_keep, p.inv_keep);
        uint32_t g = 0, n = 0;
        // epilogue of a finished work item (accumulators -> bf16 rows of dqkv), run one step into the NEXT item so
        // that nobody waits for the item's last accumulating MMA
        bool pend = false;
        uint32_t pend_n = 0;
        int64_t pend_row = 0;      // packed row of this thread's output, or -1
        int pend_head = 0;
        auto epilogue = [&]() {
            const int a = pend_n & 1;
            ptx::mbar_wait(bar(B_ACC_FULL + a), (pend_n >> 1) & 1);
            ptx::tc_fence_after();
            const uint32_t acc = tmem + 256 + (uint32_t)a * kAccCols + lane_bits + cq * 16;
            uint32_t r0[16], r1[16];
            ptx::tmem_ld_32x32_x16(acc, r0);
            if (!kIsDq) ptx::tmem_ld_32x32_x16(acc + 64, r1);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar(B_ACC_EMPTY + a));
            if (pend_row >= 0) {
                if (kIsDq) {
                    __nv_bfloat16* o = p.dqkv + pend_row * ld + pend_head * kHd + cq * 16;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        uint4 u;
                        u.x = pack_bf16x2(__uint_as_float(r0[8 * j + 0]) * p.scale, __uint_as_float(r0[8 * j + 1]) * p.scale);
                        u.y = pack_bf16x2(__uint_as_float(r0[8 * j + 2]) * p.scale, __uint_as_float(r0[8 * j + 3]) * p.scale);
                        u.z = pack_bf16x2(__uint_as_float(r0[8 * j + 4]) * p.scale, __uint_as_float(r0[8 * j + 5]) * p.scale);
                        u.w = pack_bf16x2(__uint_as_float(r0[8 * j + 6]) * p.scale, __uint_as_float(r0[8 * j + 7]) * p.scale);
                        *reinterpret_cast<uint4*>(o + 8 * j) = u;
                    }
                } else {
                    __nv_bfloat16* oK = p.dqkv + pend_row * ld + p.H + pend_head * kHd + cq * 16;
                    __nv_bfloat16* oV = oK + p.H;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        uint4 u, v;
                        v.x = pack_bf16x2(__uint_as_float(r0[8 * j + 0]) * p.inv_keep, __uint_as_float(r0[8 * j + 1]) * p.inv_keep);
                        v.y = pack_bf16x2(__uint_as_float(r0[8 * j + 2]) * p.inv_keep, __uint_as_float(r0[8 * j + 3]) * p.inv_keep);
                        v.z = pack_bf16x2(__uint_as_float(r0[8 * j + 4]) * p.inv_keep, __uint_as_float(r0[8 * j + 5]) * p.inv_keep);
                        v.w = pack_bf16x2(__uint_as_float(r0[8 * j + 6]) * p.inv_keep, __uint_as_float(r0[8 * j + 7]) * p.inv_keep);
                        u.x = pack_bf16x2(__uint_as_float(r1[8 * j + 0]) * p.scale, __uint_as_float(r1[8 * j + 1]) * p.scale);
                        u.y = pack_bf16x2(__uint_as_float(r1[8 * j + 2]) * p.scale, __uint_as_float(r1[8 * j + 3]) * p.scale);
                        u.z = pack_bf16x2(__uint_as_float(r1[8 * j + 4]) * p.scale, __uint_as_float(r1[8 * j + 5]) * p.scale);
                        u.w = pack_bf16x2(__uint_as_float(r1[8 * j + 6]) * p.scale, __uint_as_float(r1[8 * j + 7]) * p.scale);
                        *reinterpret_cast<uint4*>(oV + 8 * j) = v;
                        *reinterpret_cast<uint4*>(oK + 8 * j) = u;
                    }
                }
            }
            pend = false;
        };
        Item it = get_item<kIsDq>(p, blockIdx.x, total_items);
        for (int idx = blockIdx.x; idx < total_items; idx += stride) {
            const Item cur = it;
            it = get_item<kIsDq>(p, idx + stride, total_items);
            if (!cur.valid) continue;
            const int rr = cur.tile * kRows + r;        // this thread's query (dQ pass) / key (dKV pass)
            if (cur.nsteps == 0) {
                // dKV pass, every key of the tile masked: P == 0 exactly, so dK = dV = 0
                if (!kIsDq && rr < cur.S) {
                    __nv_bfloat16* o = p.dqkv + (int64_t)(cur.row0 + rr) * ld + p.H + cur.head * kHd + cq * 16;
                    *reinterpret_cast<uint4*>(o) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4*>(o + 8) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4*>(o + p.H) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4*>(o + p.H + 8) = make_uint4(0, 0, 0, 0);
                }
                continue;
            }
            // row constants (staged by the producer next to the row operands)
            const int rb = n & 1;
            ptx::mbar_wait(bar(B_ROW_FULL + rb), (n >> 1) & 1);
            const uint32_t rc = rowc(rb) + (uint32_t)r * 4;
            const float row_add = __uint_as_float(lds32_b(rc));
            const float row_negD = kIsDq ? __uint_as_float(lds32_b(rc + kRows * 4)) : 0.f;
            const uint32_t rkey = kDrop ? lds32_b(rc + 2 * kRows * 4) : 0u;
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar(B_ROW_EMPTY + rb));
            const float2 radd2 = make_float2(row_add, row_add), rnegD2 = make_float2(row_negD, row_negD);

            for (int s = 0; s < cur.nsteps; ++s, ++g) {
                const int sl = g & 1, st = g % kStages, b = g & 1;
                ptx::mbar_wait(bar(B_SC_FULL + sl), (g >> 1) & 1);
                ptx::tc_fence_after();
                uint32_t s_raw[16], dp_raw[16];
                ptx::tmem_ld_32x32_x16(tmem + sl * 128 + lane_bits + cq * 16, s_raw);
                ptx::tmem_ld_32x32_x16(tmem + sl * 128 + 64 + lane_bits + cq * 16, dp_raw);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar(B_SC_EMPTY + sl));      // the scores slot is free again
                ptx::mbar_wait(bar(B_STEP_FULL + st), (g / kStages) & 1);    // (already complete) acquire the column vectors
                const uint32_t cv = cols(st) + (uint32_t)cq * 64;
                ptx::mbar_wait(bar(B_ST_EMPTY + b), ((g >> 1) & 1) ^ 1);     // staged tiles of step g-2 consumed
#pragma unroll
                for (int j = 0; j < 2; ++j) {          // 8 columns -> one 16-byte chunk of the staged rows
                    uint32_t pk[4], dk[4];
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const int i = 2 * j + h2;      // float4 index: columns 4i .. 4i+3
                        const float4 ca = lds_f4(cv + i * 16);
                        float2 x0 = fma2(make_float2(__uint_as_float(s_raw[4 * i]), __uint_as_float(s_raw[4 * i + 1])), sc2, radd2);
                        float2 x1 = fma2(make_float2(__uint_as_float(s_raw[4 * i + 2]), __uint_as_float(s_raw[4 * i + 3])), sc2, radd2);
                        x0 = add2(x0, make_float2(ca.x, ca.y));
                        x1 = add2(x1, make_float2(ca.z, ca.w));
                        const float2 p0 = make_float2(ex2_approx(x0.x), ex2_approx(x0.y));
                        const float2 p1 = make_float2(ex2_approx(x1.x), ex2_approx(x1.y));
                        float2 d0 = make_float2(__uint_as_float(dp_raw[4 * i]), __uint_as_float(dp_raw[4 * i + 1]));
                        float2 d1 = make_float2(__uint_as_float(dp_raw[4 * i + 2]), __uint_as_float(dp_raw[4 * i + 3]));
                        float2 pd0 = p0, pd1 = p1;     // dropped probabilities feed dV (rescaled by 1/keep once, on dV)
                        if (kDrop) {
                            const uint4 ck = lds_u4(cv + 2 * kStep * 4 + i * 16);
                            const bool k0 = attn_keep(rkey, ck.x, p.thresh32), k1 = attn_keep(rkey, ck.y, p.thresh32),
                                       k2 = attn_keep(rkey, ck.z, p.thresh32), k3 = attn_keep(rkey, ck.w, p.thresh32);
                            d0.x = k0 ? d0.x : 0.f;
                            d0.y = k1 ? d0.y : 0.f;
                            d1.x = k2 ? d1.x : 0.f;
                            d1.y = k3 ? d1.y : 0.f;
                            if (!kIsDq) {
                                pd0.x = k0 ? p0.x : 0.f;
                                pd0.y = k1 ? p0.y : 0.f;
                                pd1.x = k2 ? p1.x : 0.f;
                                pd1.y = k3 ? p1.y : 0.f;
                            }
                        }
                        // dS = P ∘ (dP_dropped / keep - D)
                        float2 nd0 = rnegD2, nd1 = rnegD2;
                        if (!kIsDq) {
                            const float4 cd = lds_f4(cv + kStep * 4 + i * 16);
                            nd0 = make_float2(cd.x, cd.y);
                            nd1 = make_float2(cd.z, cd.w);
                        }
                        const float2 e0 = mul2(p0, fma2(d0, ik2, nd0));
                        const float2 e1 = mul2(p1, fma2(d1, ik2, nd1));
                        dk[2 * h2] = pack_bf16x2(e0.x, e0.y);
                        dk[2 * h2 + 1] = pack_bf16x2(e1.x, e1.y);
                        if (!kIsDq) {
                            pk[2 * h2] = pack_bf16x2(pd0.x, pd0.y);
                            pk[2 * h2 + 1] = pack_bf16x2(pd1.x, pd1.y);
                        }
                    }
                    const uint32_t off = (uint32_t)(r * 128 + (((cq * 2 + j) ^ r7) << 4));
                    if (kIsDq) {
                        sts128_b(staged_tile(b, 0) + off, dk[0], dk[1], dk[2], dk[3]);
                    } else {
                        sts128_b(staged_tile(b, 0) + off, pk[0], pk[1], pk[2], pk[3]);
                        sts128_b(staged_tile(b, 1) + off, dk[0], dk[1], dk[2], dk[3]);
                    }
                }
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar(B_ST_FULL + b));
                if (s == 0 && pend) epilogue();        // the previous item's accumulators are long complete by now
            }
            pend = true;
            pend_n = n;
            pend_row = rr < cur.S ? (int64_t)(cur.row0 + rr) : -1;
            pend_head = cur.head;
            ++n;
        }
        if (pend) epilogue();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kScoreWarp) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<512>(tmem);
    }
}

template <bool kIsDq, bool kDrop>
int launch_pass(const CUtensorMap& q128, const CUtensorMap& q64, const CUtensorMap& dmap, const BwdTcParams& p,
                cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        MMB_CUDA(cudaFuncSetAttribute(attn_bwd_ws_kernel<kIsDq, kDrop>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      smem_bytes<kIsDq>()));
        attr_set = true;
    }
    const int items = p.tiles * p.nheads * p.nseq;
    const int grid = items < num_sms() ? items : num_sms();
    attn_bwd_ws_kernel<kIsDq, kDrop><<<grid, kWsThreads, smem_bytes<kIsDq>(), stream>>>(q128, q64, dmap, p);
    return check_launch(kIsDq ? "attn_bwd_ws_kernel<dQ>" : "attn_bwd_ws_kernel<dKV>");
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
    p.nseq = a->nseq;
    p.tiles = (a->max_seqlen + kRows - 1) / kRows;
    p.total_rows = a->total_rows;
    p.scale = 1.0f / sqrtf((float)kHd);
    p.scale_log2 = p.scale * kLog2eB;
    p.thresh32 = dropout_threshold(a->p_drop) << 16;
    p.inv_keep = dropout_inv_keep(a->p_drop);
    p.seed = a->seed;
    p.rng_stream = a->rng_stream;
    if (p.thresh32) {
        rc = launch_pass<false, true>(q128, q64, do64, p, stream);
        if (rc != MMB_OK) return rc;
        return launch_pass<true, true>(q128, q64, do128, p, stream);
    }
    rc = launch_pass<false, false>(q128, q64, do64, p, stream);
    if (rc != MMB_OK) return rc;
    return launch_pass<true, false>(q128, q64, do128, p, stream);
}

}  // namespace mmb
