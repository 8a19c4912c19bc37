// Masked self-attention FORWARD, persistent warp-specialised form (same design as attn_bwd_tc.cu), head dim 64.
// Same contract as attn_tc.cu (mmb_attn_fwd): Q|K|V read in place from [rows, 3H], context written to [rows, H],
// log2-domain LSE to [heads, rows].  Selected by mmb_attn_args.flags bit 1 (see attn.cu).
//
// One CTA per SM, persistent over work items (sequence, head, 128-query tile), steps of 64 keys, 19 warps:
//   warp 16   producer: Q tile per item (double-buffered across items), K / V tiles per step (6-stage ring) by TMA, plus
//             the step's per-key vectors (key bias in the log2 domain, dropout key) which it computes itself, the next
//             step's bias loads already in flight
//   warp 17   score MMAs  S = Q K^T (M 128 x N 64) into one of two TMEM slots
//   warp 18   accumulating MMAs  O += P V with the A operand (P, bf16 pairs) read from TENSOR MEMORY, where the
//             element-wise warps wrote it over the score columns; V is read in place as an MN-major operand
//   warps 0-15 online softmax: every thread owns one query row (TMEM lane) and 16 of the step's 64 keys; the row maximum
//             is combined across the four warps that share a lane quarter through shared memory and one 128-thread named
//             barrier per step; the running reference maximum only moves (and O, which lives in TMEM, is only rescaled)
//             when a row exceeds it by more than 2^8; the row sums stay per-thread partials until the item's epilogue
// The epilogue of an item (O / l -> bf16, coalesced through a swizzled staging tile; LSE) runs one step into the next item.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace mmb {

int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1);

namespace {

constexpr int kRows = 128, kStep = 64, kHd = 64;
constexpr int kComputeWarps = 16;
constexpr int kProducerWarp = 16, kScoreWarp = 17, kAccWarp = 18;
constexpr int kFwThreads = 19 * 32;
constexpr int kStages = 6;
constexpr float kLog2eF = 1.4426950408889634f;
constexpr int kBig = kRows * kHd * 2;     // 16 KB
constexpr int kSmall = kStep * kHd * 2;   // 8 KB

// shared-memory map (offsets from a 1024-byte aligned base)
constexpr int kOffQ = 0;                                   // [2] x 16 KB
constexpr int kOffStep = 2 * kBig;                         // [kStages][K | V] x 8 KB
constexpr int kOffCols = kOffStep + kStages * 2 * kSmall;  // [kStages][bias f32 [64] | key u32 [64]]
constexpr int kColBytes = 2 * kStep * 4;
constexpr int kOffMax = kOffCols + kStages * kColBytes;    // [2 parities][4 column groups][128 rows] f32 partial maxima
constexpr int kOffSum = kOffMax + 2 * 4 * kRows * 4;       // [4][128] f32 partial row sums (epilogue)
constexpr int kOffStage = kOffSum + 4 * kRows * 4;         // [4 quarters] x 4 KB epilogue staging
constexpr int kMetaSeqs = 2048;
constexpr int kOffMeta = kOffStage + 4 * 4096;
constexpr int kOffBars = kOffMeta + (2 * kMetaSeqs + 4) * 4;
constexpr int kFwSmem = kOffBars + 256 + 1024;
enum { B_ROW_FULL = 0, B_ROW_EMPTY = 2, B_STEP_FULL = 4, B_STEP_EMPTY = 10, B_SC_FULL = 16, B_SC_EMPTY = 18, B_ST_FULL = 20,
       B_ACC_FULL = 22, B_ACC_EMPTY = 24, B_COUNT = 26 };

struct FwParams {
    __nv_bfloat16* ctx;
    float* lse;
    const float* keybias;
    const int* cu_seqlens;
    const int* kv_end;
    const int4* work;          // mmb_attn_schedule's lists or null (items in index order)
    int H, nheads, nseq, tiles, total_rows;
    float scale_log2;
    uint32_t thresh32;
    float inv_keep;
    uint64_t seed;
    uint32_t rng_stream;
    int skip_tail;             // flags bit 3: query tiles entirely behind kv_end may stay unwritten (see mmb_attn_args.flags)
};

__device__ __forceinline__ void sts32(uint32_t addr, uint32_t x) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(x) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// three-input maximum (FMNMX3, sm_100)
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
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

__device__ __forceinline__ constexpr uint32_t idesc(bool b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | ((uint32_t)(kStep >> 3) << 17) |
           ((uint32_t)(kRows >> 4) << 24);
}
// S[128 x 64] = Q[128 x 64] K^T : both operands K-major SWIZZLE_128B tiles
__device__ __forceinline__ void mma_scores(uint32_t tmem_d, uint32_t sA, uint32_t sB) {
    const uint32_t a16 = sA >> 4, b16 = sB >> 4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        ptx::umma_bf16(tmem_d, ptx::umma_desc_from_halves(a16 + 2 * k, ptx::kDescHiSw128),
                       ptx::umma_desc_from_halves(b16 + 2 * k, ptx::kDescHiSw128), idesc(false), k > 0 ? 1u : 0u);
}
// O[128 x 64] (+)= P[128 x 64 keys] V : P from tensor memory (K-slice k in the 8 columns at tA + 16 k), V MN-major in smem
__device__ __forceinline__ void mma_pv(uint32_t tmem_d, uint32_t tA, uint32_t sB, bool accumulate) {
    const uint32_t b16 = (sB >> 4) | ((8192u >> 4) << 16);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        ptx::umma_bf16_ts(tmem_d, tA + 16 * k, ptx::umma_desc_from_halves(b16 + 128 * k, ptx::kDescHiSw128), idesc(true),
                          (accumulate || k > 0) ? 1u : 0u);
}

struct Item {
    int head, tile, row0, S, nsteps;
    bool valid;
};
// Items come either from mmb_attn_schedule's list (wl != null: one 16-byte record {first row, length, effective keys,
// head << 16 | tile} per item, longest first; the record of the NEXT item is loaded while the current one runs) or, in
// index order, from (sequence, head, tile) = decode(idx) and the sequence metadata in shared memory.
__device__ __forceinline__ int4 fetch_work(const int4* wl, int idx, int total) {
    return (wl != nullptr && idx < total) ? __ldg(wl + idx) : make_int4(0, 0, 0, 0);
}
__device__ __forceinline__ Item make_item(const FwParams& p, const int* cu, const int* kvend, int idx, int total,
                                          const int4* wl, const int4& w) {
    Item it;
    it.tile = it.head = it.row0 = it.S = it.nsteps = 0;
    it.valid = false;
    if (idx >= total) return it;
    int eff;                                // keys at index >= eff are all masked: P == 0 exactly (see mmb_attn_args)
    if (wl != nullptr) {
        it.tile = w.w & 0xffff;
        it.head = w.w >> 16;
        it.row0 = w.x;
        it.S = w.y;
        eff = w.z;
    } else {
        it.tile = idx % p.tiles;
        const int sh = idx / p.tiles;
        it.head = sh % p.nheads;
        const int seq = sh / p.nheads;
        it.row0 = cu[seq];
        it.S = cu[seq + 1] - it.row0;
        eff = it.S;
        if (kvend != nullptr) {
            const int e = kvend[seq];
            if (e > 0 && e < it.S) eff = e;
        }
    }
    it.valid = it.tile * kRows < it.S;
    it.nsteps = it.valid ? (eff + kStep - 1) / kStep : 0;
    return it;
}

template <bool kDrop>
__global__ void __launch_bounds__(kFwThreads, 1)
attn_fwd_ws_kernel(const __grid_constant__ CUtensorMap tm_q128, const __grid_constant__ CUtensorMap tm_q64, const FwParams p) {
    pdl_trigger();     // programmatic dependent launch (common.cuh)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - ptx::smem_u32(smem_raw));
    const uint32_t bars = sbase + kOffBars;
    const uint32_t tmem_slot = bars + B_COUNT * 8;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto bar = [&](int slot) { return bars + (uint32_t)slot * 8u; };
    auto q_tile = [&](int buf) { return sbase + kOffQ + (uint32_t)buf * kBig; };
    auto step_tile = [&](int st, int which) { return sbase + kOffStep + (uint32_t)(st * 2 + which) * kSmall; };
    auto cols = [&](int st) { return sbase + kOffCols + (uint32_t)st * kColBytes; };

    if (tid == 0) {
        ptx::prefetch_tensormap(&tm_q128);
        ptx::prefetch_tensormap(&tm_q64);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(bar(B_ROW_FULL + i), 1);
            ptx::mbar_init(bar(B_ROW_EMPTY + i), 1);                   // last score MMA of the item retired
            ptx::mbar_init(bar(B_SC_FULL + i), 1);
            ptx::mbar_init(bar(B_SC_EMPTY + i), 1);                    // the P V MMAs that read the slot retired
            ptx::mbar_init(bar(B_ST_FULL + i), kComputeWarps);
            ptx::mbar_init(bar(B_ACC_FULL + i), 1);
            ptx::mbar_init(bar(B_ACC_EMPTY + i), kComputeWarps);
        }
        for (int i = 0; i < kStages; ++i) {
            ptx::mbar_init(bar(B_STEP_FULL + i), 2);                   // expect_tx arrive + key-vector arrive
            ptx::mbar_init(bar(B_STEP_EMPTY + i), 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == kScoreWarp) ptx::tmem_alloc<256>(tmem_slot);
    pdl_wait();        // barriers and tensor memory are set up; everything below reads what earlier kernels wrote
    int* meta = reinterpret_cast<int*>(sgen + kOffMeta);
    const bool meta_in_smem = p.nseq <= kMetaSeqs;
    if (meta_in_smem) {
        for (int i = tid; i <= p.nseq; i += kFwThreads) meta[i] = p.cu_seqlens[i];
        if (p.kv_end != nullptr)
            for (int i = tid; i < p.nseq; i += kFwThreads) meta[kMetaSeqs + 2 + i] = p.kv_end[i];
    }
    const int* m_cu = meta_in_smem ? meta : p.cu_seqlens;
    const int* m_kv = p.kv_end == nullptr ? nullptr : (meta_in_smem ? meta + kMetaSeqs + 2 : p.kv_end);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = lds32(tmem_slot);
    int total_items = p.tiles * p.nheads * p.nseq;
    const int4* wl = nullptr;
    if (p.work != nullptr) {
        const int4 hdr = __ldg(p.work);      // {items of this list, of the dK/dV list, capacity | qskip << 30, unmasked dK/dV items}
        total_items = hdr.x;
        wl = p.work + 1;
        if (p.skip_tail && ((hdr.z >> 30) & 1)) {
            // The schedule verified that no row at or behind kv_end carries a label: those rows are masked keys for every
            // query and feed no loss, so a 128-query tile that lies entirely behind kv_end is unobservable.  Walk the
            // unmasked part of the dK/dV list instead: one record per 128-row tile that starts before kv_end.
            total_items = hdr.w;
            wl = p.work + 1 + (hdr.z & 0x3fffffff);
        }
    }
    const int stride = gridDim.x;
    // TMEM columns: score slot s at s * 64, O accumulator a at 128 + a * 64

    if (warp == kProducerWarp) {
        // ================================================================ producer
        uint32_t g = 0, n = 0;
        int4 wnext = fetch_work(wl, blockIdx.x, total_items);
        for (int idx = blockIdx.x; idx < total_items; idx += stride) {
            const Item cur = make_item(p, m_cu, m_kv, idx, total_items, wl, wnext);
            wnext = fetch_work(wl, idx + stride, total_items);
            if (cur.nsteps == 0) continue;
            const int rb = n & 1;
            const int col_q = cur.head * kHd, col_k = p.H + cur.head * kHd, col_v = 2 * p.H + cur.head * kHd;
            const uint32_t prob_base = (uint32_t)cur.head * (uint32_t)p.total_rows + (uint32_t)cur.row0;
            float vb[2];
            auto load_bias = [&](int s) {      // this lane's two key columns of step s (global loads; stored one step later)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = s * kStep + lane + 32 * h;
                    vb[h] = c < cur.S ? p.keybias[cur.row0 + c] * kLog2eF : -INFINITY;   // keys beyond the sequence: P = 0
                }
            };
            load_bias(0);
            ptx::mbar_wait(bar(B_ROW_EMPTY + rb), ((n >> 1) & 1) ^ 1);
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx(bar(B_ROW_FULL + rb), kBig);
                ptx::tma_load_2d(q_tile(rb), &tm_q128, bar(B_ROW_FULL + rb), col_q, cur.row0 + cur.tile * kRows);
            }
            __syncwarp();
            for (int s = 0; s < cur.nsteps; ++s, ++g) {
                const int st = g % kStages;
                ptx::mbar_wait(bar(B_STEP_EMPTY + st), ((g / kStages) & 1) ^ 1);
                if (ptx::elect_one()) {
                    const uint32_t fb = bar(B_STEP_FULL + st);
                    ptx::mbar_expect_tx(fb, 2 * kSmall);
                    ptx::tma_load_2d(step_tile(st, 0), &tm_q64, fb, col_k, cur.row0 + s * kStep);
                    ptx::tma_load_2d(step_tile(st, 1), &tm_q64, fb, col_v, cur.row0 + s * kStep);
                }
                __syncwarp();
                const float sb[2] = {vb[0], vb[1]};
                if (s + 1 < cur.nsteps) load_bias(s + 1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = lane + 32 * h;
                    sts32(cols(st) + (uint32_t)i * 4, __float_as_uint(sb[h]));
                    if (kDrop)
                        sts32(cols(st) + kStep * 4 + (uint32_t)i * 4,
                              attn_drop_kkey(p.seed, p.rng_stream, prob_base + (uint32_t)(s * kStep + i)));
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar(B_STEP_FULL + st));
            }
            ++n;
        }
    } else if (warp == kScoreWarp) {
        // ================================================================ score MMAs
        uint32_t g = 0, n = 0;
        int4 wnext = fetch_work(wl, blockIdx.x, total_items);
        for (int idx = blockIdx.x; idx < total_items; idx += stride) {
            const Item cur = make_item(p, m_cu, m_kv, idx, total_items, wl, wnext);
            wnext = fetch_work(wl, idx + stride, total_items);
            if (cur.nsteps == 0) continue;
            const int rb = n & 1;
            ptx::mbar_wait(bar(B_ROW_FULL + rb), (n >> 1) & 1);
            for (int s = 0; s < cur.nsteps; ++s, ++g) {
                const int st = g % kStages, sl = g & 1;
                ptx::mbar_wait(bar(B_STEP_FULL + st), (g / kStages) & 1);
                ptx::mbar_wait(bar(B_SC_EMPTY + sl), ((g >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    mma_scores(tmem + sl * 64, q_tile(rb), step_tile(st, 0));
                    ptx::umma_commit(bar(B_SC_FULL + sl));
                    if (s == cur.nsteps - 1) ptx::umma_commit(bar(B_ROW_EMPTY + rb));
                }
                __syncwarp();
            }
            ++n;
        }
    } else if (warp == kAccWarp) {
        // ================================================================ O += P V
        uint32_t g = 0, n = 0;
        int4 wnext = fetch_work(wl, blockIdx.x, total_items);
        for (int idx = blockIdx.x; idx < total_items; idx += stride) {
            const Item cur = make_item(p, m_cu, m_kv, idx, total_items, wl, wnext);
            wnext = fetch_work(wl, idx + stride, total_items);
            if (cur.nsteps == 0) continue;
            const int a = n & 1;
            ptx::mbar_wait(bar(B_ACC_EMPTY + a), ((n >> 1) & 1) ^ 1);
            for (int s = 0; s < cur.nsteps; ++s, ++g) {
                const int st = g % kStages, sl = g & 1;
                ptx::mbar_wait(bar(B_STEP_FULL + st), (g / kStages) & 1);
                ptx::mbar_wait(bar(B_ST_FULL + sl), (g >> 1) & 1);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    mma_pv(tmem + 128 + a * 64, tmem + sl * 64, step_tile(st, 1), s > 0);
                    ptx::umma_commit(bar(B_SC_EMPTY + sl));
                    ptx::umma_commit(bar(B_STEP_EMPTY + st));
                    if (s == cur.nsteps - 1) ptx::umma_commit(bar(B_ACC_FULL + a));
                }
                __syncwarp();
            }
            ++n;
        }
    } else {
        // ================================================================ online softmax (16 warps)
        const int quarter = warp & 3, cq = warp >> 2;
        const int r = quarter * 32 + lane;             // query row of the tile = TMEM lane
        const uint32_t lane_bits = (uint32_t)(quarter * 32) << 16;
        const uint32_t smax = sbase + kOffMax, ssum = sbase + kOffSum;
        const uint32_t stg = sbase + kOffStage + (uint32_t)quarter * 4096;
        auto quarter_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory"); };
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
        uint32_t g = 0, n = 0;
        // pending epilogue (previous item)
        bool pend = false;
        uint32_t pend_n = 0;
        int pend_base = 0, pend_valid = 0, pend_head = 0;
        float pend_l = 0.f, pend_m = 0.f;
        auto epilogue = [&]() {
            const int a = pend_n & 1;
            ptx::mbar_wait(bar(B_ACC_FULL + a), (pend_n >> 1) & 1);
            ptx::tc_fence_after();
            uint32_t o[16];
            ptx::tmem_ld_32x32_x16(tmem + 128 + a * 64 + lane_bits + cq * 16, o);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar(B_ACC_EMPTY + a));
            quarter_sync();                                     // the previous item's staging / sums have been consumed
            sts32(ssum + (uint32_t)(cq * kRows + r) * 4, __float_as_uint(pend_l));
            quarter_sync();
            const float l_tot = __uint_as_float(lds32(ssum + (uint32_t)r * 4)) + __uint_as_float(lds32(ssum + (uint32_t)(kRows + r) * 4)) +
                                __uint_as_float(lds32(ssum + (uint32_t)(2 * kRows + r) * 4)) +
                                __uint_as_float(lds32(ssum + (uint32_t)(3 * kRows + r) * 4));
            const float inv_l = p.inv_keep / l_tot;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t off = (uint32_t)(lane * 128 + (((cq * 2 + j) ^ (lane & 7)) << 4));
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + off),
                             "r"(pack_bf16x2(__uint_as_float(o[8 * j + 0]) * inv_l, __uint_as_float(o[8 * j + 1]) * inv_l)),
                             "r"(pack_bf16x2(__uint_as_float(o[8 * j + 2]) * inv_l, __uint_as_float(o[8 * j + 3]) * inv_l)),
                             "r"(pack_bf16x2(__uint_as_float(o[8 * j + 4]) * inv_l, __uint_as_float(o[8 * j + 5]) * inv_l)),
                             "r"(pack_bf16x2(__uint_as_float(o[8 * j + 6]) * inv_l, __uint_as_float(o[8 * j + 7]) * inv_l))
                             : "memory");
            }
            if (cq == 0 && r < pend_valid && p.lse != nullptr)
                p.lse[(int64_t)pend_head * p.total_rows + pend_base + r] = pend_m + log2f(l_tot);
            quarter_sync();
#pragma unroll
            for (int it = 0; it < 2; ++it) {               // this warp: rows cq*8 .. cq*8+7 of the quarter's 32, whole 128-byte rows
                const int i = cq * 8 + it * 4 + (lane >> 3), c = lane & 7;
                const uint4 v = lds_u4(stg + i * 128 + ((c ^ (i & 7)) << 4));
                const int trow = quarter * 32 + i;
                if (trow < pend_valid)
                    *reinterpret_cast<uint4*>(p.ctx + (int64_t)(pend_base + trow) * p.H + pend_head * kHd + c * 8) = v;
            }
            pend = false;
        };
        int4 wnext = fetch_work(wl, blockIdx.x, total_items);
        for (int idx = blockIdx.x; idx < total_items; idx += stride) {
            const Item cur = make_item(p, m_cu, m_kv, idx, total_items, wl, wnext);
            wnext = fetch_work(wl, idx + stride, total_items);
            if (cur.nsteps == 0) continue;
            const int a = n & 1;
            const uint32_t prob_base = (uint32_t)cur.head * (uint32_t)p.total_rows + (uint32_t)cur.row0;
            const uint32_t qkey = kDrop ? attn_drop_qkey(p.seed, p.rng_stream, prob_base + (uint32_t)(cur.tile * kRows + r)) : 0u;
            float m_ref = -INFINITY, l_part = 0.f;
            for (int s = 0; s < cur.nsteps; ++s, ++g) {
                const int sl = g & 1, st = g % kStages;
                ptx::mbar_wait(bar(B_SC_FULL + sl), (g >> 1) & 1);
                ptx::tc_fence_after();
                const uint32_t t_s = tmem + sl * 64 + lane_bits + cq * 16;
                uint32_t s_raw[16];
                ptx::tmem_ld_32x32_x16(t_s, s_raw);
                ptx::tmem_ld_wait();
                ptx::mbar_wait(bar(B_STEP_FULL + st), (g / kStages) & 1);    // (already complete) acquire the key vectors
                const uint32_t cv = cols(st) + (uint32_t)cq * 64;
                // x = s * scale + bias (log2 domain), partial row maximum over this thread's 16 keys
                float x[16];
                float m_part = -INFINITY;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 b = lds_f4(cv + i * 16);
                    const float2 x0 = fma2(make_float2(__uint_as_float(s_raw[4 * i]), __uint_as_float(s_raw[4 * i + 1])), sc2,
                                           make_float2(b.x, b.y));
                    const float2 x1 = fma2(make_float2(__uint_as_float(s_raw[4 * i + 2]), __uint_as_float(s_raw[4 * i + 3])), sc2,
                                           make_float2(b.z, b.w));
                    x[4 * i] = x0.x; x[4 * i + 1] = x0.y; x[4 * i + 2] = x1.x; x[4 * i + 3] = x1.y;
                    m_part = max3(max3(m_part, x0.x, x0.y), x1.x, x1.y);
                }
                // row maximum across the four warps of this lane quarter
                const uint32_t mx = smax + (uint32_t)((g & 1) * 4 * kRows) * 4;
                sts32(mx + (uint32_t)(cq * kRows + r) * 4, __float_as_uint(m_part));
                quarter_sync();
                const float m_tile = max3(fmaxf(__uint_as_float(lds32(mx + (uint32_t)r * 4)), __uint_as_float(lds32(mx + (uint32_t)(kRows + r) * 4))),
                                          __uint_as_float(lds32(mx + (uint32_t)(2 * kRows + r) * 4)),
                                          __uint_as_float(lds32(mx + (uint32_t)(3 * kRows + r) * 4)));
                // lazy reference maximum: move it (and rescale O and the partial sum) only when exceeded by more than 2^8
                const bool need = m_tile > m_ref + 8.f;     // identical for the four threads of a row
                if (__any_sync(0xffffffffu, need)) {
                    const float m_new = need ? m_tile : m_ref;
                    const float corr = (need && m_ref != -INFINITY) ? ex2_approx(m_ref - m_new) : 1.f;
                    if (s > 0) {
                        // the previous P V must have retired before O is touched
                        ptx::mbar_wait(bar(B_SC_EMPTY + (sl ^ 1)), ((g - 1) >> 1) & 1);
                        ptx::tc_fence_after();
                        uint32_t o[16];
                        const uint32_t t_o = tmem + 128 + a * 64 + lane_bits + cq * 16;
                        ptx::tmem_ld_32x32_x16(t_o, o);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
                        asm volatile(
                            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                            "%15, %16};" ::"r"(t_o), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]),
                            "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15])
                            : "memory");
                    }
                    l_part *= corr;
                    m_ref = m_new;
                }
                const float2 nm2 = make_float2(-m_ref, -m_ref);
                uint32_t pk[8];
                float2 lacc = make_float2(0.f, 0.f);
                uint32_t kk[16];       // this thread's 16 column dropout keys: four 16-byte shared loads
                if (kDrop) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint4 q = lds_u4(cv + kStep * 4 + i * 16);
                        kk[4 * i] = q.x; kk[4 * i + 1] = q.y; kk[4 * i + 2] = q.z; kk[4 * i + 3] = q.w;
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 xs = add2(make_float2(x[2 * i], x[2 * i + 1]), nm2);
                    float2 pr = make_float2(ex2_approx(xs.x), ex2_approx(xs.y));
                    lacc = add2(lacc, pr);
                    if (kDrop) {       // dropped entries become 0; the 1/(1-p) rescale is applied once in the epilogue
                        pr.x = attn_keep(qkey, kk[2 * i], p.thresh32) ? pr.x : 0.f;
                        pr.y = attn_keep(qkey, kk[2 * i + 1], p.thresh32) ? pr.y : 0.f;
                    }
                    pk[i] = pack_bf16x2(pr.x, pr.y);
                }
                l_part += lacc.x + lacc.y;
                // P (bf16 pairs) over the score columns this thread just read: the A operand of the P V MMA (K-slice cq)
                ptx::tmem_st_32x32_x8(t_s, pk);
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar(B_ST_FULL + sl));
                if (s == 0 && pend) epilogue();
            }
            pend = true;
            pend_n = n;
            pend_base = cur.row0 + cur.tile * kRows;
            pend_valid = min(kRows, cur.S - cur.tile * kRows);
            pend_head = cur.head;
            pend_l = l_part;
            pend_m = m_ref;
            ++n;
        }
        if (pend) epilogue();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kScoreWarp) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<256>(tmem);
    }
}

template <bool kDrop>
int launch(const CUtensorMap& q128, const CUtensorMap& q64, const FwParams& p, cudaStream_t stream) {
    MMB_ENSURE_SMEM(kFwSmem, attn_fwd_ws_kernel<kDrop>);
    const int items = p.tiles * p.nheads * p.nseq;
    const int grid = items < persistent_sms() ? items : persistent_sms();
    launch_pdl(attn_fwd_ws_kernel<kDrop>, dim3(grid), dim3(kFwThreads), kFwSmem, stream, q128, q64, p);
    return check_launch("attn_fwd_ws_kernel");
}

}  // namespace

int launch_attn_fwd_ws(const mmb_attn_args* a, cudaStream_t stream) {
    CUtensorMap q128, q64;
    const uint64_t rows = (uint64_t)a->total_rows, H = (uint64_t)a->H;
    int rc = make_tmap_bf16(&q128, a->qkv, 3 * H, rows, 3 * H, 64, 128);
    if (rc == MMB_OK) rc = make_tmap_bf16(&q64, a->qkv, 3 * H, rows, 3 * H, 64, 64);
    if (rc != MMB_OK) return rc;
    FwParams p;
    p.ctx = (__nv_bfloat16*)a->ctx;
    p.lse = a->lse;
    p.keybias = a->keybias;
    p.cu_seqlens = a->cu_seqlens;
    p.kv_end = a->kv_end;
    p.work = (const int4*)a->work;
    p.H = a->H;
    p.nheads = a->nheads;
    p.nseq = a->nseq;
    p.tiles = (a->max_seqlen + kRows - 1) / kRows;
    p.total_rows = a->total_rows;
    p.scale_log2 = kLog2eF / sqrtf((float)kHd);
    p.thresh32 = dropout_threshold(a->p_drop) << 16;
    p.inv_keep = dropout_inv_keep(a->p_drop);
    p.seed = a->seed;
    p.rng_stream = a->rng_stream;
    p.skip_tail = (a->flags & 8) ? 1 : 0;
    return p.thresh32 ? launch<true>(q128, q64, p, stream) : launch<false>(q128, q64, p, stream);
}

}  // namespace mmb
