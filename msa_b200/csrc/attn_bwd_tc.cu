// Masked self-attention BACKWARD on 5th-generation tensor cores (tcgen05 + TMEM + TMA), head dim 64.
// Autograd of modeling_bert.py:115-140 (eager_attention_forward) for the packed batch; same contract as
// mmb_attn_bwd in include/mmbert_sm100.h (dQ|dK|dV written into dqkv [rows, 3H], D = rowsum(dO ∘ O) precomputed).
//
// Two launches of ONE persistent, warp-specialised kernel template (no transposes, no atomics):
//   dQ  pass  rows (TMEM lanes) = 128 queries, steps of 64 keys:
//        S  = Q K^T,  dP  = dO V^T   -> TMEM;  dS  = P ∘ (dP - D) -> bf16 -> back into TMEM;  dQ += dS K
//   dKV pass  rows (TMEM lanes) = 128 keys,    steps of 64 queries:
//        S^T = K Q^T, dP^T = V dO^T  -> TMEM;  P^T, dS^T -> TMEM;  dV += P^T dO,  dK += dS^T Q
// Both are "row operands (A0, A1) x step operands (B0, B1)":  scores = A0 B0^T and A1 B1^T from shared memory, then
// the accumulating MMAs take their A operand (P^T / dS) straight from TENSOR MEMORY — the element-wise warps write
// the bf16 values over the score columns they just read (tcgen05.st), so nothing is staged through shared memory
// (no swizzled stores, no generic->async proxy fence) — against B0 / B1 read in place as MN-major operands.
//
// One CTA per SM (all 512 TMEM columns, ~200 KB shared memory), 19 warps:
//   warp 16   TMA producer: row operands (double-buffered across work items), step operands (6-stage ring), and the
//             per-row / per-column vectors (key bias, -LSE, -D, dropout keys: prepared by attn.cu's prep kernel) next
//             to them — copies only, the producer computes nothing
//   warp 17   score MMAs (one elected thread; warp-uniform control flow so descriptors live in uniform registers)
//   warp 18   accumulating MMAs
//   warps 0-15 element-wise stage: every thread owns one TMEM lane (row) and 16 of the step's 64 columns (= one
//             K = 16 slice of the accumulating MMA); the probabilities are rebuilt from the forward's log-sum-exp,
//             ~7 instructions per probability (packed FFMA2 / FADD2 / FMUL2, one MUFU.EX2, one-IMAD dropout
//             decision — common.cuh: attn_keep)
// Rings: scores 2 slots x (64 + 64) TMEM columns (a slot is free again when the accumulating MMAs that read it have
// retired), accumulators 2 buffers (the epilogue of one work item runs one step into the next item).  The CTA walks
// work items (sequence, head, 128-row tile) with a grid stride — in index order, or longest first through the work
// lists of mmb_attn_schedule (attn.cu) when the caller passes them; every role evaluates the same item list, and
// all ring positions are carried across items.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace mmb {

int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1);
int make_tmap_rec16(CUtensorMap* out, const void* ptr, uint64_t records, uint64_t planes, uint32_t nbox);

#ifdef MMB_ATTN_TRACE
// bring-up aid (never compiled into the product library): per-role event timestamps of CTA 0
__device__ unsigned long long g_attn_trace[4][1024][4];
#endif

namespace {

constexpr int kRows = 128;           // TMEM lanes per work item
constexpr int kStep = 64;            // columns per step
constexpr int kHd = 64;              // head dim
constexpr int kComputeWarps = 16;
constexpr int kProducerWarp = 16, kScoreWarp = 17, kAccWarp = 18;
constexpr int kWsThreads = 19 * 32;
constexpr int kStages = 6;           // step-operand ring
constexpr float kLog2eB = 1.4426950408889634f;
constexpr int kBig = kRows * kHd * 2;     // 16 KB: 128-row operand tile (128 B rows, SWIZZLE_128B)
constexpr int kSmall = kStep * kHd * 2;   // 8 KB: 64-row operand tile

// shared-memory map (offsets from a 1024-byte aligned base)
constexpr int kOffRow = 0;                               // [2 buffers][A0 | A1] x 16 KB
constexpr int kOffStep = 4 * kBig;                       // [kStages][B0 | B1] x 8 KB
template <bool kIsDq> constexpr int off_cols() { return kOffStep + kStages * 2 * kSmall; }
constexpr int kColBytes = kStep * 16;                    // per stage: one 16-byte record per column (attn.cu: Rq / Rk)
constexpr int kRowcBytes = kRows * 16;                   // per row buffer: one record per row
template <bool kIsDq> constexpr int off_rowc() { return off_cols<kIsDq>() + kStages * kColBytes; }
// epilogue staging: per TMEM lane quarter a 32-row x 128-byte tile per output matrix (dQ | dV, dK), XOR-swizzled
template <bool kIsDq> constexpr int stage_bytes() { return (kIsDq ? 1 : 2) * 4 * 4096; }
template <bool kIsDq> constexpr int off_stage() { return off_rowc<kIsDq>() + 2 * kRowcBytes; }
// sequence metadata (cu_seqlens, kv_end) of up to kMetaSeqs sequences, copied once per CTA
constexpr int kMetaSeqs = 2048;
constexpr int kMetaBytes = (2 * kMetaSeqs + 4) * 4;
template <bool kIsDq> constexpr int off_meta() { return off_stage<kIsDq>() + stage_bytes<kIsDq>(); }
template <bool kIsDq> constexpr int off_bars() { return off_meta<kIsDq>() + kMetaBytes; }
template <bool kIsDq> constexpr int smem_bytes() { return off_bars<kIsDq>() + 256 + 1024; }
// barrier slots (8 bytes each)
// (the step ring may grow to 8 stages without renumbering)
enum { B_ROW_FULL = 0, B_ROW_EMPTY = 2, B_STEP_FULL = 4, B_STEP_EMPTY = 12, B_SC_FULL = 20, B_SC_EMPTY = 22,
       B_ST_FULL = 24, B_ACC_FULL = 26, B_ACC_EMPTY = 28, B_COUNT = 30 };

struct BwdTcParams {
    __nv_bfloat16* dqkv;
    const int* cu_seqlens;
    const int* kv_end;
    const int4* work;          // mmb_attn_schedule's lists or null (items in index order)
    int zero_tiles_done;       // flags bit 4: the zero-fill items of the dK/dV list may be left out (see mmb_attn_args.flags)
    int H, nheads, nseq, tiles, total_rows;
    float scale_log2, scale;
    uint32_t thresh32;
    float inv_keep;
};

__device__ __forceinline__ uint32_t lds32_b(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
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
// D[128 x 64] (+)= A[128 x 64] * B,  A in tensor memory (bf16 pairs; K-slice k of 16 values lives in the 8 columns
// at tA + 16 k — where the element-wise warp of column group k wrote it), B = [64 (contraction rows)][64] in shared
// memory read in place as an MN-major operand
__device__ __forceinline__ void mma_tmn(uint32_t tmem_d, uint32_t tA, uint32_t sB, bool accumulate) {
    const uint32_t b16 = (sB >> 4) | ((8192u >> 4) << 16);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        ptx::umma_bf16_ts(tmem_d, tA + 16 * k, ptx::umma_desc_from_halves(b16 + 128 * k, ptx::kDescHiSw128),
                          idesc_m128_n64(true), (accumulate || k > 0) ? 1u : 0u);
}

#ifdef MMB_ATTN_TRACE
#define TRACE(role, step, ev)                                                                  \
    do {                                                                                       \
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (step) < 1024u) g_attn_trace[role][step][ev] = clock64(); \
    } while (0)
#else
#define TRACE(role, step, ev) do {} while (0)
#endif

struct Item {
    int head, tile, row0, S, nsteps;        // nsteps == 0: nothing to run (invalid tile, or a fully masked key tile)
    bool valid;
};
// Work-item metadata is fetched one item ahead: fetch_item only ISSUES the loads (nothing depends on them until
// make_item runs at the top of the next iteration), so their latency hides under the current item's steps.
// Items come either from mmb_attn_schedule's list (wl != null: one 16-byte record {first row, length, effective keys,
// head << 16 | tile} per item, longest first) or, in index order, from (sequence, head, tile) = decode(idx).
struct ItemRaw {
    int idx, row0, row1, kvend, ht;
};
__device__ __forceinline__ ItemRaw fetch_item(const BwdTcParams& p, const int4* wl, const int* cu, const int* kvend,
                                              int idx, int total) {
    ItemRaw r;
    r.idx = idx;
    r.row0 = r.row1 = r.kvend = r.ht = 0;
    if (idx < total) {
        if (wl != nullptr) {
            const int4 w = __ldg(wl + idx);
            r.row0 = w.x;
            r.row1 = w.y;       // the length, not the end row
            r.kvend = w.z;
            r.ht = w.w;
        } else {
            const int seq = idx / (p.tiles * p.nheads);
            r.row0 = cu[seq];
            r.row1 = cu[seq + 1];
            if (kvend != nullptr) r.kvend = kvend[seq];
        }
    }
    return r;
}
// qskip (mmb_attn_schedule verified it): dctx is exactly zero on the rows >= kv_end of every sequence, so those QUERY rows
// add nothing to dK / dV and their dQ is zero — the dK/dV pass stops at ceil(eff / 64) query steps and a dQ tile behind
// kv_end only stores zeros
template <bool kIsDq>
__device__ __forceinline__ Item make_item(const BwdTcParams& p, const ItemRaw& r, int total, bool listed, bool qskip) {
    Item it;
    if (listed) {
        it.tile = r.ht & 0xffff;
        it.head = r.ht >> 16;
        it.S = r.row1;
    } else {
        it.tile = r.idx % p.tiles;
        it.head = (r.idx / p.tiles) % p.nheads;
        it.S = r.row1 - r.row0;
    }
    it.row0 = r.row0;
    it.valid = r.idx < total && it.tile * kRows < it.S;
    int eff = it.S;                         // keys at index >= eff are all masked: P == 0 exactly (see mmb_attn_args)
    if (r.kvend > 0 && r.kvend < it.S) eff = r.kvend;
    if (!it.valid) it.nsteps = 0;
    else if (kIsDq) it.nsteps = (qskip && it.tile * kRows >= eff) ? 0 : (eff + kStep - 1) / kStep;
    else it.nsteps = (it.tile * kRows >= eff) ? 0 : ((qskip ? eff : it.S) + kStep - 1) / kStep;
    return it;
}

// Element-wise stage of one step for one thread: 16 scores + 16 dP values of its row -> 16 probabilities (dKV pass only)
// and 16 dS values, as bf16 pairs.   P = 2^(s * scale + row_add + col_add),   dS = P ∘ (keep ? dP / (1-p) : 0  -  D).
// Column records (16 bytes each, attn.cu): dQ pass Rk {bias, key}; dKV pass Rq {-LSE, -D, key}.  kTail: some of the
// columns lie beyond the sequence (only in an item's last step) and must give P = 0.
template <bool kIsDq, bool kDrop, bool kTail>
__device__ __forceinline__ void ew_step(const uint32_t (&s_raw)[16], const uint32_t (&dp_raw)[16], uint32_t cv, int ncol,
                                        float scale_log2, float row_add, float row_negD, float inv_keep, uint32_t rkey,
                                        uint32_t thresh32, uint32_t (&pk)[8], uint32_t (&dk)[8]) {
    const float2 sc2 = make_float2(scale_log2, scale_log2), radd2 = make_float2(row_add, row_add);
#pragma unroll
    for (int c = 0; c < 16; c += 2) {          // two columns at a time
        const uint4 r0 = lds_u4(cv + c * 16), r1 = lds_u4(cv + c * 16 + 16);
        float a0 = __uint_as_float(r0.x), a1 = __uint_as_float(r1.x);
        float nd0 = kIsDq ? row_negD : __uint_as_float(r0.y), nd1 = kIsDq ? row_negD : __uint_as_float(r1.y);
        if (kTail) {
            a0 = c < ncol ? a0 : -INFINITY;
            a1 = c + 1 < ncol ? a1 : -INFINITY;
            if (!kIsDq) {
                nd0 = c < ncol ? nd0 : 0.f;
                nd1 = c + 1 < ncol ? nd1 : 0.f;
            }
        }
        const float2 t = fma2(make_float2(__uint_as_float(s_raw[c]), __uint_as_float(s_raw[c + 1])), sc2, radd2);
        const float p0 = ex2_approx(t.x + a0), p1 = ex2_approx(t.y + a1);
        float d0 = __uint_as_float(dp_raw[c]), d1 = __uint_as_float(dp_raw[c + 1]);
        float pd0 = p0, pd1 = p1;              // dropped probabilities feed dV (rescaled by 1/keep once, on dV)
        if (kDrop) {
            const bool k0 = attn_keep(rkey, kIsDq ? r0.y : r0.z, thresh32), k1 = attn_keep(rkey, kIsDq ? r1.y : r1.z, thresh32);
            d0 = k0 ? d0 : 0.f;
            d1 = k1 ? d1 : 0.f;
            if (!kIsDq) {
                pd0 = k0 ? p0 : 0.f;
                pd1 = k1 ? p1 : 0.f;
            }
        }
        float2 e;
        if (kIsDq) {
            e = mul2(make_float2(p0, p1), fma2(make_float2(d0, d1), make_float2(inv_keep, inv_keep), make_float2(nd0, nd1)));
        } else {
            e.x = p0 * fmaf(d0, inv_keep, nd0);
            e.y = p1 * fmaf(d1, inv_keep, nd1);
        }
        dk[c >> 1] = pack_bf16x2(e.x, e.y);
        if (!kIsDq) pk[c >> 1] = pack_bf16x2(pd0, pd1);
    }
}

template <bool kIsDq, bool kDrop>
__global__ void __launch_bounds__(kWsThreads, 1)
attn_bwd_ws_kernel(const __grid_constant__ CUtensorMap tm_qkv128, const __grid_constant__ CUtensorMap tm_qkv64,
                   const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_aux128,
                   const __grid_constant__ CUtensorMap tm_aux64, const BwdTcParams p) {
    pdl_trigger();     // programmatic dependent launch (common.cuh)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = sbase + off_bars<kIsDq>();
    const uint32_t tmem_slot = bars + B_COUNT * 8;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto bar = [&](int slot) { return bars + (uint32_t)slot * 8u; };
    auto row_tile = [&](int buf, int which) { return sbase + kOffRow + (uint32_t)(buf * 2 + which) * kBig; };
    auto step_tile = [&](int st, int which) { return sbase + kOffStep + (uint32_t)(st * 2 + which) * kSmall; };
    auto cols = [&](int st) { return sbase + off_cols<kIsDq>() + (uint32_t)st * kColBytes; };
    auto rowc = [&](int buf) { return sbase + off_rowc<kIsDq>() + (uint32_t)buf * kRowcBytes; };

    if (tid == 0) {
        ptx::prefetch_tensormap(&tm_qkv128);
        ptx::prefetch_tensormap(&tm_qkv64);
        ptx::prefetch_tensormap(&tm_do);
        ptx::prefetch_tensormap(&tm_aux128);
        ptx::prefetch_tensormap(&tm_aux64);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(bar(B_ROW_FULL + i), 1);
            ptx::mbar_init(bar(B_ROW_EMPTY + i), 1 + kComputeWarps);   // last scores MMA + row constants read
            ptx::mbar_init(bar(B_SC_FULL + i), 1);
            ptx::mbar_init(bar(B_SC_EMPTY + i), 1);                    // accumulating MMAs that read the slot retired
            ptx::mbar_init(bar(B_ST_FULL + i), kComputeWarps);
            ptx::mbar_init(bar(B_ACC_FULL + i), 1);
            ptx::mbar_init(bar(B_ACC_EMPTY + i), kComputeWarps);
        }
        for (int i = 0; i < kStages; ++i) {
            ptx::mbar_init(bar(B_STEP_FULL + i), 1);
            ptx::mbar_init(bar(B_STEP_EMPTY + i), 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == kScoreWarp) ptx::tmem_alloc<512>(tmem_slot);
    pdl_wait();        // barriers and tensor memory are set up; everything below reads what earlier kernels wrote
    // sequence metadata -> shared memory (items are decoded from it by every role; an item whose tile lies beyond its
    // sequence is skipped for the price of two shared loads instead of a global round trip)
    extern __shared__ uint8_t smem_generic[];
    int* meta = reinterpret_cast<int*>(smem_generic + (sbase - ptx::smem_u32(smem_generic)) + off_meta<kIsDq>());
    const bool meta_in_smem = p.nseq <= kMetaSeqs;
    if (meta_in_smem) {
        for (int i = tid; i <= p.nseq; i += kWsThreads) meta[i] = p.cu_seqlens[i];
        if (p.kv_end != nullptr)
            for (int i = tid; i < p.nseq; i += kWsThreads) meta[kMetaSeqs + 2 + i] = p.kv_end[i];
    }
    const int* m_cu = meta_in_smem ? meta : p.cu_seqlens;
    const int* m_kv = p.kv_end == nullptr ? nullptr : (meta_in_smem ? meta + kMetaSeqs + 2 : p.kv_end);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = lds32_b(tmem_slot);
    int total_items = p.tiles * p.nheads * p.nseq;
    const int4* wl = nullptr;
    bool qskip = false;
    if (p.work != nullptr) {
        const int4 hdr = __ldg(p.work);      // {items of the forward / dQ list, of the dK/dV list, capacity | qskip << 30}
        qskip = (hdr.z >> 30) & 1;
        const int cap = hdr.z & 0x3fffffff;
        // with the query tail skipped both passes walk the dK/dV list: live tiles by effective length, zero-fill tiles last
        total_items = (kIsDq && !qskip) ? hdr.x : hdr.y;
        if (qskip && p.zero_tiles_done) total_items = hdr.w;      // only the tiles that start before kv_end; the rest is zero already
        wl = p.work + 1 + ((kIsDq && !qskip) ? 0 : cap);
    }
    const bool listed = wl != nullptr;
    const int stride = gridDim.x;
    // TMEM columns: scores slot s at s * 128 (S | dP), accumulators from 256
    constexpr int kAccCols = kIsDq ? 64 : 128;

    if (warp == kProducerWarp) {
        // ================================================================ TMA producer (one elected lane issues)
        // per-row records (attn.cu): planes [0, nh) = Rq {-LSE, -D, row dropout key}, [nh, 2 nh) = Rk {bias, column key};
        // the dQ pass wants Rq for its rows and Rk for its columns, the dKV pass the other way round
        uint32_t g = 0, n = 0;
        ItemRaw raw = fetch_item(p, wl, m_cu, m_kv, blockIdx.x, total_items);
        for (int idx = blockIdx.x; idx < total_items; idx += stride) {
            const Item cur = make_item<kIsDq>(p, raw, total_items, listed, qskip);
            raw = fetch_item(p, wl, m_cu, m_kv, idx + stride, total_items);
            if (cur.nsteps == 0) continue;
            const int rb = n & 1;
            const int col_q = cur.head * kHd, col_k = p.H + cur.head * kHd, col_v = 2 * p.H + cur.head * kHd;
            const int nh = p.nheads;
            // ---- row operands + row vectors
            ptx::mbar_wait(bar(B_ROW_EMPTY + rb), ((n >> 1) & 1) ^ 1);
            if (ptx::elect_one()) {
                const uint32_t fb = bar(B_ROW_FULL + rb);
                const int r0 = cur.row0 + cur.tile * kRows;
                ptx::mbar_expect_tx(fb, 2 * kBig + kRowcBytes);
                if (kIsDq) {
                    ptx::tma_load_2d(row_tile(rb, 0), &tm_qkv128, fb, col_q, r0);                   // Q
                    ptx::tma_load_2d(row_tile(rb, 1), &tm_do, fb, cur.head * kHd, r0);              // dO
                    ptx::tma_load_2d(rowc(rb), &tm_aux128, fb, 2 * r0, cur.head);                   // Rq
                } else {
                    ptx::tma_load_2d(row_tile(rb, 0), &tm_qkv128, fb, col_k, r0);                   // K
                    ptx::tma_load_2d(row_tile(rb, 1), &tm_qkv128, fb, col_v, r0);                   // V
                    ptx::tma_load_2d(rowc(rb), &tm_aux128, fb, 2 * r0, nh + cur.head);              // Rk
                }
            }
            __syncwarp();
            // ---- step operands + column vectors
            for (int s = 0; s < cur.nsteps; ++s, ++g) {
                const int st = g % kStages;
                ptx::mbar_wait(bar(B_STEP_EMPTY + st), ((g / kStages) & 1) ^ 1);
                if (ptx::elect_one()) {
                    const uint32_t fb = bar(B_STEP_FULL + st);
                    const int r0 = cur.row0 + s * kStep;
                    ptx::mbar_expect_tx(fb, 2 * kSmall + kColBytes);
                    if (kIsDq) {
                        ptx::tma_load_2d(step_tile(st, 0), &tm_qkv64, fb, col_k, r0);               // K
                        ptx::tma_load_2d(step_tile(st, 1), &tm_qkv64, fb, col_v, r0);               // V
                        ptx::tma_load_2d(cols(st), &tm_aux64, fb, 2 * r0, nh + cur.head);           // Rk
                    } else {
                        ptx::tma_load_2d(step_tile(st, 0), &tm_qkv64, fb, col_q, r0);               // Q
                        ptx::tma_load_2d(step_tile(st, 1), &tm_do, fb, cur.head * kHd, r0);         // dO
                        ptx::tma_load_2d(cols(st), &tm_aux64, fb, 2 * r0, cur.head);                // Rq
                    }
                }
                __syncwarp();
            }
            ++n;
        }
    } else if (warp == kScoreWarp) {
        // ================================================================ score MMAs: S = A0 B0^T, dP = A1 B1^T
        uint32_t g = 0, n = 0;
        ItemRaw raw = fetch_item(p, wl, m_cu, m_kv, blockIdx.x, total_items);
        for (int idx = blockIdx.x; idx < total_items; idx += stride) {
            const Item cur = make_item<kIsDq>(p, raw, total_items, listed, qskip);
            raw = fetch_item(p, wl, m_cu, m_kv, idx + stride, total_items);
            if (cur.nsteps == 0) continue;
            const int rb = n & 1;
            ptx::mbar_wait(bar(B_ROW_FULL + rb), (n >> 1) & 1);
            for (int s = 0; s < cur.nsteps; ++s, ++g) {
                const int st = g % kStages, sl = g & 1;
                TRACE(1, g, 0);
                ptx::mbar_wait(bar(B_STEP_FULL + st), (g / kStages) & 1);
                TRACE(1, g, 1);
                ptx::mbar_wait(bar(B_SC_EMPTY + sl), ((g >> 1) & 1) ^ 1);
                TRACE(1, g, 2);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    mma_kk(tmem + sl * 128, row_tile(rb, 0), step_tile(st, 0), false);
                    mma_kk(tmem + sl * 128 + 64, row_tile(rb, 1), step_tile(st, 1), false);
                    ptx::umma_commit(bar(B_SC_FULL + sl));
                    if (s == cur.nsteps - 1) ptx::umma_commit(bar(B_ROW_EMPTY + rb));   // row operands no longer read
                }
                __syncwarp();
                TRACE(1, g, 3);
            }
            ++n;
        }
    } else if (warp == kAccWarp) {
        // ================================================================ accumulating MMAs
        uint32_t g = 0, n = 0;
        ItemRaw raw = fetch_item(p, wl, m_cu, m_kv, blockIdx.x, total_items);
        for (int idx = blockIdx.x; idx < total_items; idx += stride) {
            const Item cur = make_item<kIsDq>(p, raw, total_items, listed, qskip);
            raw = fetch_item(p, wl, m_cu, m_kv, idx + stride, total_items);
            if (cur.nsteps == 0) continue;
            const int a = n & 1;
            ptx::mbar_wait(bar(B_ACC_EMPTY + a), ((n >> 1) & 1) ^ 1);
            const uint32_t acc = tmem + 256 + (uint32_t)a * kAccCols;
            for (int s = 0; s < cur.nsteps; ++s, ++g) {
                const int st = g % kStages, sl = g & 1;
                TRACE(2, g, 0);
                ptx::mbar_wait(bar(B_STEP_FULL + st), (g / kStages) & 1);     // (long complete) acquire the TMA tiles
                ptx::mbar_wait(bar(B_ST_FULL + sl), (g >> 1) & 1);
                TRACE(2, g, 1);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    if (kIsDq) {
                        mma_tmn(acc, tmem + sl * 128 + 64, step_tile(st, 0), s > 0);       // dQ += dS   K
                    } else {
                        mma_tmn(acc, tmem + sl * 128, step_tile(st, 1), s > 0);            // dV += P^T  dO
                        mma_tmn(acc + 64, tmem + sl * 128 + 64, step_tile(st, 0), s > 0);  // dK += dS^T Q
                    }
                    ptx::umma_commit(bar(B_SC_EMPTY + sl));
                    ptx::umma_commit(bar(B_STEP_EMPTY + st));
                    if (s == cur.nsteps - 1) ptx::umma_commit(bar(B_ACC_FULL + a));
                }
                __syncwarp();
                TRACE(2, g, 2);
            }
            ++n;
        }
    } else {
        // ================================================================ element-wise stage (16 warps)
        const int r = (warp & 3) * 32 + lane;          // row of the work item's tile = TMEM lane
        const int cq = warp >> 2;                      // which 16 of the step's 64 columns
        const uint32_t lane_bits = (uint32_t)((warp & 3) * 32) << 16;
        const int64_t ld = 3 * (int64_t)p.H;
        uint32_t g = 0, n = 0;
        // epilogue of a finished work item (accumulators -> bf16 rows of dqkv), run one step into the NEXT item so
        // that nobody waits for the item's last accumulating MMA
        bool pend = false;
        uint32_t pend_n = 0;
        int pend_base = 0;         // packed row of the tile's first row
        int pend_valid = 0;        // rows of the tile that belong to the sequence
        int pend_head = 0;
        // The four warps of a TMEM lane quarter own the same 32 rows (16 columns each).  Writing them straight to global
        // memory costs one 32-byte wavefront per row and instruction (~1000-2000 LSU cycles per work item); instead they
        // meet in a swizzled shared-memory tile and write whole 128-byte rows, 4 rows per instruction.
        const int quarter = warp & 3;
        const uint32_t stg = sbase + off_stage<kIsDq>() + (uint32_t)quarter * 4096;
        auto quarter_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory"); };
        auto flush = [&](uint32_t tile, int col_off) {     // this warp: rows cq*8 .. cq*8+7 of the quarter's 32
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int i = cq * 8 + it * 4 + (lane >> 3), c = lane & 7;
                const uint4 v = lds_u4(tile + i * 128 + ((c ^ (i & 7)) << 4));
                const int trow = quarter * 32 + i;
                if (trow < pend_valid)
                    *reinterpret_cast<uint4*>(p.dqkv + (int64_t)(pend_base + trow) * ld + col_off + c * 8) = v;
            }
        };
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
            quarter_sync();                                 // the previous item's staging tile has been read by all four
            const float s0 = kIsDq ? p.scale : p.inv_keep;  // dQ * scale | dV * 1/keep
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t off = (uint32_t)(lane * 128 + (((cq * 2 + j) ^ (lane & 7)) << 4));
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + off),
                             "r"(pack_bf16x2(__uint_as_float(r0[8 * j + 0]) * s0, __uint_as_float(r0[8 * j + 1]) * s0)),
                             "r"(pack_bf16x2(__uint_as_float(r0[8 * j + 2]) * s0, __uint_as_float(r0[8 * j + 3]) * s0)),
                             "r"(pack_bf16x2(__uint_as_float(r0[8 * j + 4]) * s0, __uint_as_float(r0[8 * j + 5]) * s0)),
                             "r"(pack_bf16x2(__uint_as_float(r0[8 * j + 6]) * s0, __uint_as_float(r0[8 * j + 7]) * s0))
                             : "memory");
                if (!kIsDq)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + 4 * 4096 + off),
                                 "r"(pack_bf16x2(__uint_as_float(r1[8 * j + 0]) * p.scale, __uint_as_float(r1[8 * j + 1]) * p.scale)),
                                 "r"(pack_bf16x2(__uint_as_float(r1[8 * j + 2]) * p.scale, __uint_as_float(r1[8 * j + 3]) * p.scale)),
                                 "r"(pack_bf16x2(__uint_as_float(r1[8 * j + 4]) * p.scale, __uint_as_float(r1[8 * j + 5]) * p.scale)),
                                 "r"(pack_bf16x2(__uint_as_float(r1[8 * j + 6]) * p.scale, __uint_as_float(r1[8 * j + 7]) * p.scale))
                                 : "memory");
            }
            quarter_sync();
            if (kIsDq) {
                flush(stg, pend_head * kHd);                                // dQ
            } else {
                flush(stg, 2 * p.H + pend_head * kHd);                      // dV
                flush(stg + 4 * 4096, p.H + pend_head * kHd);               // dK
            }
            pend = false;
        };
        ItemRaw raw = fetch_item(p, wl, m_cu, m_kv, blockIdx.x, total_items);
        for (int idx = blockIdx.x; idx < total_items; idx += stride) {
            const Item cur = make_item<kIsDq>(p, raw, total_items, listed, qskip);
            raw = fetch_item(p, wl, m_cu, m_kv, idx + stride, total_items);
            if (!cur.valid) continue;
            const int rr = cur.tile * kRows + r;        // this thread's query (dQ pass) / key (dKV pass)
            if (cur.nsteps == 0) {
                if (kIsDq && rr < cur.S) {
                    // dQ pass, query tile behind kv_end with the zero-gradient tail verified: dQ = 0
                    __nv_bfloat16* o = p.dqkv + (int64_t)(cur.row0 + rr) * ld + cur.head * kHd + cq * 16;
                    *reinterpret_cast<uint4*>(o) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4*>(o + 8) = make_uint4(0, 0, 0, 0);
                }
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
            const uint4 rrec = lds_u4(rowc(rb) + (uint32_t)r * 16);     // dQ pass: Rq {-LSE, -D, key}; dKV pass: Rk {bias, key}
            // rows beyond the sequence (they belong to the next one, or are TMA zero fill): P = 0
            const float row_add = rr < cur.S ? __uint_as_float(rrec.x) : -INFINITY;
            const float row_negD = (kIsDq && rr < cur.S) ? __uint_as_float(rrec.y) : 0.f;
            const uint32_t rkey = kIsDq ? rrec.z : rrec.y;
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar(B_ROW_EMPTY + rb));

            for (int s = 0; s < cur.nsteps; ++s, ++g) {
                const int sl = g & 1, st = g % kStages;
                if (warp == 0) TRACE(3, g, 0);
                ptx::mbar_wait(bar(B_SC_FULL + sl), (g >> 1) & 1);
                if (warp == 0) TRACE(3, g, 1);
                ptx::tc_fence_after();
                const uint32_t t_s = tmem + sl * 128 + lane_bits + cq * 16, t_dp = t_s + 64;
                uint32_t s_raw[16], dp_raw[16];
                ptx::tmem_ld_32x32_x16(t_s, s_raw);
                ptx::tmem_ld_32x32_x16(t_dp, dp_raw);
                ptx::tmem_ld_wait();
                if (warp == 0) TRACE(3, g, 2);
                ptx::mbar_wait(bar(B_STEP_FULL + st), (g / kStages) & 1);    // (already complete) acquire the column vectors
                const uint32_t cv = cols(st) + (uint32_t)cq * 256;    // this thread's 16 column records
                // columns beyond the sequence (only possible in an item's last step): P = 0
                const int ncol = cur.S - s * kStep - cq * 16;       // this thread's columns 0 .. 15 are valid while < ncol
                const bool tail = ncol < 16;
                uint32_t pk_all[8], dk_all[8];         // this thread's 16 probabilities / dS values as bf16 pairs
                if (tail)
                    ew_step<kIsDq, kDrop, true>(s_raw, dp_raw, cv, ncol, p.scale_log2, row_add, row_negD, p.inv_keep, rkey,
                                                p.thresh32, pk_all, dk_all);
                else
                    ew_step<kIsDq, kDrop, false>(s_raw, dp_raw, cv, ncol, p.scale_log2, row_add, row_negD, p.inv_keep, rkey,
                                                 p.thresh32, pk_all, dk_all);
                // back into tensor memory, over the columns this thread just read: the A operand of the accumulating
                // MMAs (K-slice cq).  dQ pass: dS over dP;  dKV pass: P^T over S^T and dS^T over dP^T.
                if (!kIsDq) ptx::tmem_st_32x32_x8(t_s, pk_all);
                ptx::tmem_st_32x32_x8(t_dp, dk_all);
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bar(B_ST_FULL + sl));
                if (warp == 0) TRACE(3, g, 3);
                if (s == 0 && pend) epilogue();        // the previous item's accumulators are long complete by now
            }
            pend = true;
            pend_n = n;
            pend_base = cur.row0 + cur.tile * kRows;
            pend_valid = min(kRows, cur.S - cur.tile * kRows);
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
int launch_pass(const CUtensorMap& q128, const CUtensorMap& q64, const CUtensorMap& dmap, const CUtensorMap& aux128,
                const CUtensorMap& aux64, const BwdTcParams& p, cudaStream_t stream) {
    MMB_ENSURE_SMEM(smem_bytes<kIsDq>(), attn_bwd_ws_kernel<kIsDq, kDrop>);
    const int items = p.tiles * p.nheads * p.nseq;
    const int grid = items < persistent_sms() ? items : persistent_sms();
    launch_pdl(attn_bwd_ws_kernel<kIsDq, kDrop>, dim3(grid), dim3(kWsThreads), smem_bytes<kIsDq>(), stream, q128, q64, dmap, aux128,
               aux64, p);
    return check_launch(kIsDq ? "attn_bwd_ws_kernel<dQ>" : "attn_bwd_ws_kernel<dKV>");
}

}  // namespace

// a->bwd_ws must already hold the records written by attn.cu's attn_bwd_prep_kernel.
int launch_attn_bwd_tc(const mmb_attn_args* a, cudaStream_t stream) {
    CUtensorMap q128, q64, do128, do64, aux128, aux64;
    const uint64_t rows = (uint64_t)a->total_rows, H = (uint64_t)a->H;
    int rc = make_tmap_bf16(&q128, a->qkv, 3 * H, rows, 3 * H, 64, 128);
    if (rc == MMB_OK) rc = make_tmap_bf16(&q64, a->qkv, 3 * H, rows, 3 * H, 64, 64);
    if (rc == MMB_OK) rc = make_tmap_bf16(&do128, a->dctx, H, rows, H, 64, 128);
    if (rc == MMB_OK) rc = make_tmap_bf16(&do64, a->dctx, H, rows, H, 64, 64);
    if (rc == MMB_OK) rc = make_tmap_rec16(&aux128, a->bwd_ws, rows, 2 * (uint64_t)a->nheads, kRows);
    if (rc == MMB_OK) rc = make_tmap_rec16(&aux64, a->bwd_ws, rows, 2 * (uint64_t)a->nheads, kStep);
    if (rc != MMB_OK) return rc;
    BwdTcParams p;
    p.dqkv = (__nv_bfloat16*)a->dqkv;
    p.cu_seqlens = a->cu_seqlens;
    p.kv_end = a->kv_end;
    p.work = (const int4*)a->work;
    p.zero_tiles_done = (a->flags & 16) ? 1 : 0;
    p.H = a->H;
    p.nheads = a->nheads;
    p.nseq = a->nseq;
    p.tiles = (a->max_seqlen + kRows - 1) / kRows;
    p.total_rows = a->total_rows;
    p.scale = 1.0f / sqrtf((float)kHd);
    p.scale_log2 = p.scale * kLog2eB;
    p.thresh32 = dropout_threshold(a->p_drop) << 16;
    p.inv_keep = dropout_inv_keep(a->p_drop);
    if (p.thresh32) {
        rc = launch_pass<false, true>(q128, q64, do64, aux128, aux64, p, stream);
        if (rc != MMB_OK) return rc;
        return launch_pass<true, true>(q128, q64, do128, aux128, aux64, p, stream);
    }
    rc = launch_pass<false, false>(q128, q64, do64, aux128, aux64, p, stream);
    if (rc != MMB_OK) return rc;
    return launch_pass<true, false>(q128, q64, do128, aux128, aux64, p, stream);
}

}  // namespace mmb

#ifdef MMB_ATTN_TRACE
extern "C" int mmb_debug_attn_trace(void* host_dst) {
    return cudaMemcpyFromSymbol(host_dst, mmb::g_attn_trace, sizeof(mmb::g_attn_trace)) == cudaSuccess ? 0 : -3;
}
#endif
