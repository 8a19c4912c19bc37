// Persistent warp-specialised bf16 GEMM for sm_100a.
//
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> smem ring (kStages) -> tcgen05.mma (cta_group::1, 128 x BN x 16)
//   -> fp32 accumulators in TMEM (double buffered: 2 x BN columns) -> tcgen05.ld -> epilogue -> global.
//
// Warp roles (384 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer, warp 2 = TMEM
// allocator, warp 3 idle, warps 4..11 = epilogue (warp w may only touch TMEM lanes 32*(w%4)..+31, so the
// two epilogue warpgroups split the BN columns in halves).
//
// Replaces the cuBLASLt dispatches behind every nn.Linear of the reference path and their autograd
// backward (see include/mmbert_sm100.h: mmb_gemm).  Both operands may be K-major (activations, weights)
// or MN-major (dY^T / X^T read in place for wgrad), selected in the UMMA instruction descriptor.
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "ptx.cuh"

namespace mmb {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int kGemmThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiWarps = 8;

template <int BN>
struct GemmCfg {
    static constexpr int kStages = (BN == 256) ? 4 : 6;
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = 2 * BN;  // 512 or 256: power of two
    static constexpr int kBarBytes = 256;
    static constexpr int kEpiStageBytes = 8 * 4096;  // per-warp epilogue staging (kNumEpiWarps * kStageBytesPerWarp)
    static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + kEpiStageBytes + 1024;  // + alignment slack
};

struct GemmParams {
    void* C;
    void* aux;
    void* aux2;
    const float* bias;
    int64_t ldc, ldaux;
    int M, N, K;
    int a_mn, b_mn;  // 1 = MN-major operand
    int epilogue;
    int split_k;
    int kb_per_split;  // k-blocks (of BK) per split
    int m_tiles, n_tiles, total_tiles;
    float alpha;
    // UMMA smem-descriptor parameters per operand (bytes): leading/stride byte offsets, per-UMMA_K advance
    uint32_t a_lbo, a_sbo, a_step, b_lbo, b_sbo, b_step;
    int dbg;  // bring-up only: bit 4 = epilogue skips global stores, bit 5 = epilogue skips the TMEM loads too
    int tma_store;  // CTA-pair kernel: bf16 outputs leave through cp.async.bulk.tensor stores (tmC / tmAux)
    int stages;     // CTA-pair kernel: depth of the operand ring (6, or 5 to make room for the column-sum array)
    float* colsum;  // CTA-pair kernel, 8 epilogue warps: [N] column sums of the bf16-rounded output, or null
    const int* row_live;   // CTA-pair kernel, GELU_GRAD / MUL_AUX TMA-store epilogues: per-row flags (mmb_gemm_args.row_live)
    int dead_zeroed;       // MUL_AUX with row_live: all-dead slices of C already hold zeros, skip them
};

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32], int ncols_valid) {
    if (ncols_valid >= 32) {
        uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
            u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
            u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
            u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
            d4[i] = u;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < ncols_valid) dst[i] = __float2bfloat16_rn(v[i]);
    }
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6), a=bf16 [7,10), b=bf16 [10,13),
// a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------- epilogue
// Each epilogue thread owns one accumulator row (its TMEM lane).  Writing that row straight to global memory
// makes every warp store touch 32 different 128-byte lines with 16 bytes each; measured on B200 this store
// pattern, not the math, bounded the K = 768 GEMMs (FFN1+GELU: 166 us with stores, 64 us without).  bf16
// outputs are therefore transposed through a per-warp shared-memory staging tile (32 rows x 64 B, XOR
// swizzled: conflict-free both ways) and written as full 64-byte row segments, 8 rows per warp instruction.
constexpr int kStageBytesPerWarp = 4096;   // [0, 2048): C chunk, [2048, 4096): aux chunk (GELU out / DGELU in)

__device__ __forceinline__ uint32_t stage_off(int r, int j) { return (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }
// explicit shared-space accesses (the staging pointer is a 32-bit shared address: guarantees STS/LDS, no generic LD/ST)
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void lds128(uint32_t addr, uint32_t& x, uint32_t& y, uint32_t& z, uint32_t& w) {
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(addr) : "memory");
}

// this lane's 32 values (one row) -> staging tile
__device__ __forceinline__ void stage_row(uint32_t st, const float (&v)[32], int lane) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
        sts128(st + stage_off(lane, j), pack_bf16x2(v[8 * j + 0], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
               pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
}
// staging tile -> global, coalesced: instruction i writes rows 8i..8i+7, 4 lanes x 16 B per row
__device__ __forceinline__ void flush_stage(uint32_t st, __nv_bfloat16* base, int64_t ld, int row_base, int M, int col0,
                                            int N, int lane) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + (lane >> 2), j = lane & 3;
        uint32_t w0, w1, w2, w3;
        lds128(st + stage_off(r, j), w0, w1, w2, w3);
        const int grow = row_base + r, gcol = col0 + j * 8;
        if (grow < M && gcol < N) {
            __nv_bfloat16* dst = base + (size_t)grow * ld + gcol;
            if (gcol + 8 <= N) {
                *reinterpret_cast<uint4*>(dst) = make_uint4(w0, w1, w2, w3);
            } else {
                unsigned short* d16 = reinterpret_cast<unsigned short*>(dst);
                const int n = N - gcol;
                if (0 < n) d16[0] = (unsigned short)(w0 & 0xFFFFu);
                if (1 < n) d16[1] = (unsigned short)(w0 >> 16);
                if (2 < n) d16[2] = (unsigned short)(w1 & 0xFFFFu);
                if (3 < n) d16[3] = (unsigned short)(w1 >> 16);
                if (4 < n) d16[4] = (unsigned short)(w2 & 0xFFFFu);
                if (5 < n) d16[5] = (unsigned short)(w2 >> 16);
                if (6 < n) d16[6] = (unsigned short)(w3 & 0xFFFFu);
            }
        }
    }
    __syncwarp();
}
// global -> staging tile, coalesced (DGELU reads the saved pre-activation)
__device__ __forceinline__ void fill_stage(uint32_t st, const __nv_bfloat16* base, int64_t ld, int row_base, int M, int col0,
                                           int N, int lane) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + (lane >> 2), j = lane & 3;
        const int grow = row_base + r, gcol = col0 + j * 8;
        uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
        if (grow < M && gcol < N) {
            const __nv_bfloat16* src = base + (size_t)grow * ld + gcol;
            if (gcol + 8 <= N) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(src));
                w0 = q.x; w1 = q.y; w2 = q.z; w3 = q.w;
            } else {
                const unsigned short* s16 = reinterpret_cast<const unsigned short*>(src);
                const int n = N - gcol;
                if (0 < n) w0 |= (uint32_t)s16[0];
                if (1 < n) w0 |= (uint32_t)s16[1] << 16;
                if (2 < n) w1 |= (uint32_t)s16[2];
                if (3 < n) w1 |= (uint32_t)s16[3] << 16;
                if (4 < n) w2 |= (uint32_t)s16[4];
                if (5 < n) w2 |= (uint32_t)s16[5] << 16;
                if (6 < n) w3 |= (uint32_t)s16[6];
            }
        }
        sts128(st + stage_off(r, j), w0, w1, w2, w3);
    }
    __syncwarp();
}

// One 32-column chunk of the warp's 32 accumulator rows.  Warp-collective: every lane must call it.
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, float (&v)[32], int row_base, int lane, int col0,
                                               uint32_t st) {
    const int nvalid = min(32, p.N - col0);
    const int row = row_base + lane;
    const bool row_ok = row < p.M;
    if (p.bias != nullptr && p.epilogue != MMB_EPI_ATOMIC_ADD_F32 && p.epilogue != MMB_EPI_DGELU_BF16 &&
        p.epilogue != MMB_EPI_MUL_AUX_BF16) {
        if (nvalid == 32) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + i);
                const float2 lo = add2(make_float2(v[4 * i], v[4 * i + 1]), make_float2(b.x, b.y));
                const float2 hi = add2(make_float2(v[4 * i + 2], v[4 * i + 3]), make_float2(b.z, b.w));
                v[4 * i] = lo.x;
                v[4 * i + 1] = lo.y;
                v[4 * i + 2] = hi.x;
                v[4 * i + 3] = hi.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i < nvalid) v[i] += __ldg(p.bias + col0 + i);
        }
    }
    __nv_bfloat16* Cb = reinterpret_cast<__nv_bfloat16*>(p.C);
    switch (p.epilogue) {
        case MMB_EPI_STORE_BF16: {
            stage_row(st, v, lane);
            flush_stage(st, Cb, p.ldc, row_base, p.M, col0, p.N, lane);
        } break;
        case MMB_EPI_GELU_BF16: {
            if (p.aux != nullptr) stage_row(st + 2048, v, lane);   // pre-activation, rounded to bf16 by the pack
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                float2 g, dg;
                gelu_fast2(make_float2(v[i], v[i + 1]), g, dg);
                v[i] = g.x;
                v[i + 1] = g.y;
            }
            stage_row(st, v, lane);
            if (p.aux != nullptr)
                flush_stage(st + 2048, reinterpret_cast<__nv_bfloat16*>(p.aux), p.ldaux, row_base, p.M, col0, p.N, lane);
            flush_stage(st, Cb, p.ldc, row_base, p.M, col0, p.N, lane);
        } break;
        case MMB_EPI_GELU_GRAD_BF16: {
#pragma unroll
            for (int j = 0; j < 4; ++j) {     // 8 columns -> one 16-byte chunk of each staging row (keeps few values live)
                uint32_t pg[4], pd[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float2 g, dg;
                    gelu_fast2(make_float2(v[8 * j + 2 * i], v[8 * j + 2 * i + 1]), g, dg);
                    pg[i] = pack_bf16x2(g.x, g.y);
                    pd[i] = pack_bf16x2(dg.x, dg.y);
                }
                sts128(st + stage_off(lane, j), pg[0], pg[1], pg[2], pg[3]);
                sts128(st + 2048 + stage_off(lane, j), pd[0], pd[1], pd[2], pd[3]);
            }
            flush_stage(st + 2048, reinterpret_cast<__nv_bfloat16*>(p.aux), p.ldaux, row_base, p.M, col0, p.N, lane);
            flush_stage(st, Cb, p.ldc, row_base, p.M, col0, p.N, lane);
        } break;
        case MMB_EPI_MUL_AUX_BF16: {
            fill_stage(st + 2048, reinterpret_cast<const __nv_bfloat16*>(p.aux), p.ldaux, row_base, p.M, col0, p.N, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t q0, q1, q2, q3;
                lds128(st + 2048 + stage_off(lane, j), q0, q1, q2, q3);
                const float2 f0 = unpack_bf16x2(q0), f1 = unpack_bf16x2(q1), f2 = unpack_bf16x2(q2), f3 = unpack_bf16x2(q3);
                v[8 * j + 0] *= f0.x;
                v[8 * j + 1] *= f0.y;
                v[8 * j + 2] *= f1.x;
                v[8 * j + 3] *= f1.y;
                v[8 * j + 4] *= f2.x;
                v[8 * j + 5] *= f2.y;
                v[8 * j + 6] *= f3.x;
                v[8 * j + 7] *= f3.y;
            }
            stage_row(st, v, lane);
            flush_stage(st, Cb, p.ldc, row_base, p.M, col0, p.N, lane);
        } break;
        case MMB_EPI_RELU_BF16: {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            stage_row(st, v, lane);
            flush_stage(st, Cb, p.ldc, row_base, p.M, col0, p.N, lane);
        } break;
        case MMB_EPI_DGELU_BF16: {
            fill_stage(st + 2048, reinterpret_cast<const __nv_bfloat16*>(p.aux), p.ldaux, row_base, p.M, col0, p.N, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t q0, q1, q2, q3;
                lds128(st + 2048 + stage_off(lane, j), q0, q1, q2, q3);
                const float2 f0 = unpack_bf16x2(q0), f1 = unpack_bf16x2(q1), f2 = unpack_bf16x2(q2), f3 = unpack_bf16x2(q3);
                v[8 * j + 0] *= gelu_fast_grad(f0.x);
                v[8 * j + 1] *= gelu_fast_grad(f0.y);
                v[8 * j + 2] *= gelu_fast_grad(f1.x);
                v[8 * j + 3] *= gelu_fast_grad(f1.y);
                v[8 * j + 4] *= gelu_fast_grad(f2.x);
                v[8 * j + 5] *= gelu_fast_grad(f2.y);
                v[8 * j + 6] *= gelu_fast_grad(f3.x);
                v[8 * j + 7] *= gelu_fast_grad(f3.y);
            }
            stage_row(st, v, lane);
            flush_stage(st, Cb, p.ldc, row_base, p.M, col0, p.N, lane);
        } break;
        case MMB_EPI_STORE_F32: {
            if (row_ok) {
                float* dst = reinterpret_cast<float*>(p.C) + (size_t)row * p.ldc + col0;
                if (nvalid == 32) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i < nvalid) dst[i] = v[i];
                }
            }
        } break;
        case MMB_EPI_ATOMIC_ADD_F32: {
            if (row_ok) {
                float* dst = reinterpret_cast<float*>(p.C) + (size_t)row * p.ldc + col0;
                if (nvalid == 32) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        ptx::red_add_v4(dst + 4 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i < nvalid) atomicAdd(dst + i, v[i]);
                }
            }
        } break;
        default: break;
    }
}

// ---------------------------------------------------------------------------------------------- TMA-store epilogue
// CTA-pair kernel, bf16 outputs.  The per-warp staging tile is laid out exactly as a swizzled TMA box, so that ONE elected
// lane hands the whole tile to the TMA unit (cp.async.bulk.tensor.2d.global.shared::cta): full-line writes, no LDS / STG
// instructions, rows / columns beyond the matrix clipped by the tensor map.  Measured before (B200, K = 768 GEMMs of the
// MOSEI shape): with the st.global epilogue the tile period followed the ~1.5 TB/s the 64-byte row segments reached
// (QKV 1 198 TF/s against 1 575 with the epilogue's stores removed).
//   8 epilogue warps : staging = 32 rows x 64 columns (128-byte rows, SWIZZLE_128B), two stores per 128-column strip
//   16 epilogue warps: staging = 2 streams x (32 rows x 32 columns) (64-byte rows, SWIZZLE_64B: stage_off above)
__device__ __forceinline__ uint32_t stage_off128(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

__device__ __forceinline__ void bias_add_chunk(const GemmParams& p, float (&v)[32], int col0) {
    const int nvalid = min(32, p.N - col0);
    if (p.bias == nullptr) return;
    if (nvalid == 32) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + i);
            const float2 lo = add2(make_float2(v[4 * i], v[4 * i + 1]), make_float2(b.x, b.y));
            const float2 hi = add2(make_float2(v[4 * i + 2], v[4 * i + 3]), make_float2(b.z, b.w));
            v[4 * i] = lo.x;
            v[4 * i + 1] = lo.y;
            v[4 * i + 2] = hi.x;
            v[4 * i + 3] = hi.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < nvalid) v[i] += __ldg(p.bias + col0 + i);
    }
}

// the staging tile may be rewritten once the TMA unit has read the previous store out of it (lane 0 issued it)
__device__ __forceinline__ void stage_acquire(int lane) {
    if (lane == 0) ptx::bulk_wait_read<0>();
    __syncwarp();
}
// 8-warp layout: this lane's 32 values -> columns [32 * half, 32 * half + 32) of its row in the 64-column tile
__device__ __forceinline__ void stage_row128(uint32_t st, const float (&v)[32], int lane, int half) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
        sts128(st + stage_off128(lane, 4 * half + j), pack_bf16x2(v[8 * j + 0], v[8 * j + 1]),
               pack_bf16x2(v[8 * j + 2], v[8 * j + 3]), pack_bf16x2(v[8 * j + 4], v[8 * j + 5]),
               pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
}
// all lanes' shared-memory writes -> visible to the async proxy -> one lane issues the store(s)
__device__ __forceinline__ void stage_release(const CUtensorMap* tm0, uint32_t st0, const CUtensorMap* tm1, uint32_t st1,
                                              int col, int row, int lane) {
    ptx::fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        ptx::tma_store_2d(tm0, st0, col, row);
        if (tm1 != nullptr) ptx::tma_store_2d(tm1, st1, col, row);
        ptx::bulk_commit();
    }
}

// element-wise part of the bf16 epilogues on one 32-column chunk (bias already added); MUL_AUX is handled by the caller.
// Returns through v (the C values) and, for GELU_GRAD / GELU with aux, through w (the aux stream).
__device__ __forceinline__ void activate_chunk(int epilogue, bool gelu_aux, float (&v)[32], float (&w)[32]) {
    if (epilogue == MMB_EPI_GELU_GRAD_BF16) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            float2 g, dg;
            gelu_fast2(make_float2(v[i], v[i + 1]), g, dg);
            v[i] = g.x;
            v[i + 1] = g.y;
            w[i] = dg.x;
            w[i + 1] = dg.y;
        }
    } else if (epilogue == MMB_EPI_GELU_BF16) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            float2 g, dg;
            if (gelu_aux) {
                w[i] = v[i];            // pre-activation (rounded to bf16 by the pack)
                w[i + 1] = v[i + 1];
            }
            gelu_fast2(make_float2(v[i], v[i + 1]), g, dg);
            v[i] = g.x;
            v[i + 1] = g.y;
        }
    } else if (epilogue == MMB_EPI_RELU_BF16) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    }
}

// MMB_EPI_CE_STATS: masked-LM cross entropy fused into the tied-decoder GEMM (MMBertForPretraining.py:293 + :381-384).
// The logits are never written.  Each epilogue thread owns one row and this warp's 128 columns (column group g =
// colw / 128): a labelled row keeps an online (max, sum exp) over acc * alpha + bias, picks the label's logit, and stores
// its bf16 logits into the dlogits buffer for the backward.  A warp without a labelled row returns without touching TMEM:
// ~99 % of the rows (frame positions, unselected tokens) cost nothing.  Workspace planes of M floats: [2g] = max,
// [2g + 1] = sum, [2G] = label logit, [2G + 1] = row-written flag, G = ceil(N / 128).  Warp-collective.
constexpr int kCeGroupCols = 128;
__device__ __forceinline__ void ce_stats_warp(const GemmParams& p, uint32_t taddr0, int row_base, int lane, int colw) {
    const int row = row_base + lane;
    const int label = row < p.M ? __ldg(reinterpret_cast<const int*>(p.aux) + row) : -100;
    const bool on = label != -100;
    if (!__any_sync(0xffffffffu, on)) return;
    constexpr float kLog2e = 1.4426950408889634f;
    float m = -INFINITY, s = 0.f, picked = 0.f;
    bool has_pick = false;
    __nv_bfloat16* crow = p.C != nullptr ? reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)(on ? row : 0) * p.ldc : nullptr;
#pragma unroll 1
    for (int c = 0; c < kCeGroupCols; c += 32) {
        const int col0 = colw + c;
        if (col0 >= p.N) break;   // warp-uniform
        uint32_t raw[32];
        ptx::tmem_ld_32x32(taddr0 + c, raw);
        ptx::tmem_ld_wait();
        if (on) {
            const int nvalid = min(32, p.N - col0);
            float v[32];
            if (nvalid == 32) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 b = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col0) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[4 * i + 0] = fmaf(__uint_as_float(raw[4 * i + 0]), p.alpha, b.x);
                    v[4 * i + 1] = fmaf(__uint_as_float(raw[4 * i + 1]), p.alpha, b.y);
                    v[4 * i + 2] = fmaf(__uint_as_float(raw[4 * i + 2]), p.alpha, b.z);
                    v[4 * i + 3] = fmaf(__uint_as_float(raw[4 * i + 3]), p.alpha, b.w);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    v[i] = i < nvalid ? fmaf(__uint_as_float(raw[i]), p.alpha, p.bias ? __ldg(p.bias + col0 + i) : 0.f) : -INFINITY;
            }
            float cm = v[0];
#pragma unroll
            for (int i = 1; i < 32; ++i) cm = fmaxf(cm, v[i]);
            const float mn = fmaxf(m, cm);
            const float off = -mn * kLog2e;
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += ex2_approx(fmaf(v[i], kLog2e, off));     // exp2(-inf) = 0 for the padded tail
            s = fmaf(s, ex2_approx((m - mn) * kLog2e), acc);
            m = mn;
            const int rel = label - col0;
            if (rel >= 0 && rel < 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i == rel) picked = v[i];
                has_pick = true;
            }
            if (crow != nullptr) {
                if (nvalid == 32) {
                    store_bf16x32(crow + col0, v, 32);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i < nvalid) crow[col0 + i] = __float2bfloat16_rn(v[i]);
                }
            }
        }
        __syncwarp();
    }
    if (on) {
        float* st = reinterpret_cast<float*>(p.aux2);
        const int g = colw / kCeGroupCols, G = (p.N + kCeGroupCols - 1) / kCeGroupCols;
        st[(size_t)(2 * g) * p.M + row] = m;
        st[(size_t)(2 * g + 1) * p.M + row] = s;
        if (has_pick) st[(size_t)(2 * G) * p.M + row] = picked;
        if (g == 0 && crow != nullptr) st[(size_t)(2 * G + 1) * p.M + row] = 1.0f;
    }
}

// ---------------------------------------------------------------------------------------------- 1-CTA kernel
template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
    using Cfg = GemmCfg<BN>;
    pdl_trigger();     // programmatic dependent launch (common.cuh): the next kernel's prologue may overlap this grid's tail
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t smem_base = ptx::smem_u32(smem);
    const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
    // barrier layout: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then the TMEM base slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + 2 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::kStages * Cfg::kStageBytes + 8 * (2 * Cfg::kStages + 4));
    uint8_t* epi_stage = smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kBarBytes;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(tfull_bar(a), 1);
            ptx::mbar_init(tempty_bar(a), kNumEpiWarps);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc<Cfg::kTmemCols>(ptx::smem_u32(tmem_slot));
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();        // everything above touched shared / tensor memory only; operands and outputs are global

    const int total_kb = (p.K + BK - 1) / BK;
    const int mn_tiles = p.m_tiles * p.n_tiles;

    if (warp == 0 && lane == 0) {
        // ================================ TMA producer ================================
        int stage = 0;
        uint32_t phase = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const int ks = t / mn_tiles;
            const int r = t - ks * mn_tiles;
            const int m_blk = r / p.n_tiles;   // n fastest: the tiles that share an A row-panel run concurrently
            const int n_blk = r - m_blk * p.n_tiles;
            const int kb0 = ks * p.kb_per_split;
            const int kb1 = min(total_kb, kb0 + p.kb_per_split);
            for (int kb = kb0; kb < kb1; ++kb) {
                ptx::mbar_wait(empty_bar(stage), phase ^ 1);
                const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
                const uint32_t sb = sa + Cfg::kABytes;
                ptx::mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
                if (!p.a_mn) {
                    ptx::tma_load_2d(sa, &tmA, full_bar(stage), kb * BK, m_blk * BM);
                } else {
#pragma unroll
                    for (int i = 0; i < BM / 64; ++i)
                        ptx::tma_load_2d(sa + i * 8192, &tmA, full_bar(stage), m_blk * BM + i * 64, kb * BK);
                }
                if (!p.b_mn) {
                    ptx::tma_load_2d(sb, &tmB, full_bar(stage), kb * BK, n_blk * BN);
                } else {
#pragma unroll
                    for (int i = 0; i < BN / 64; ++i)
                        ptx::tma_load_2d(sb + i * 8192, &tmB, full_bar(stage), n_blk * BN + i * 64, kb * BK);
                }
                if (++stage == Cfg::kStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ================================ MMA issuer ================================
        const uint32_t idesc = make_idesc(BM, BN, p.a_mn, p.b_mn);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const int ks = t / mn_tiles;
            const int kb0 = ks * p.kb_per_split;
            const int kb1 = min(total_kb, kb0 + p.kb_per_split);
            ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1);
            ptx::tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            for (int kb = kb0; kb < kb1; ++kb) {
                ptx::mbar_wait(full_bar(stage), phase);
                ptx::tc_fence_after();
                const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
                const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t da = ptx::umma_desc_sw128(sa + k * p.a_step, p.a_lbo, p.a_sbo);
                    const uint64_t db = ptx::umma_desc_sw128(sb + k * p.b_step, p.b_lbo, p.b_sbo);
                    ptx::umma_bf16(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                }
                ptx::umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
                if (++stage == Cfg::kStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            ptx::umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ================================ epilogue ================================
        const int ew = warp - kEpiWarp0;
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        const int half = ew >> 2;      // which half of the BN columns
        const uint32_t stage_buf = ptx::smem_u32(epi_stage) + ew * kStageBytesPerWarp;
        constexpr int kColsPerWarp = BN / 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const int ks = t / mn_tiles;
            const int r = t - ks * mn_tiles;
            const int m_blk = r / p.n_tiles;   // n fastest: the tiles that share an A row-panel run concurrently
            const int n_blk = r - m_blk * p.n_tiles;
            ptx::mbar_wait(tfull_bar(acc), acc_phase);
            ptx::tc_fence_after();
            const int row_base = m_blk * BM + quarter * 32;
            if constexpr (kColsPerWarp == kCeGroupCols) {
                if (p.epilogue == MMB_EPI_CE_STATS && n_blk * BN + half * kColsPerWarp < p.N)
                    ce_stats_warp(p, tmem_base + acc * BN + half * kColsPerWarp + ((uint32_t)(quarter * 32) << 16), row_base, lane,
                                  n_blk * BN + half * kColsPerWarp);
            }
#pragma unroll 1
            for (int c = p.epilogue == MMB_EPI_CE_STATS ? kColsPerWarp : 0; c < kColsPerWarp; c += 32) {
                const int col0 = n_blk * BN + half * kColsPerWarp + c;
                if (col0 >= p.N) break;  // warp-uniform
                if (p.dbg & 32) continue;
                uint32_t raw[32];
                const uint32_t taddr = tmem_base + acc * BN + half * kColsPerWarp + c + ((uint32_t)(quarter * 32) << 16);
                ptx::tmem_ld_32x32(taddr, raw);
                ptx::tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                if (p.alpha != 1.0f) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
                }
                if (!(p.dbg & 16)) epilogue_chunk(p, v, row_base, lane, col0, stage_buf);
                __syncwarp();  // reconverge before the next .sync.aligned TMEM load
            }
            // all TMEM reads of this warp are complete (tcgen05.wait::ld) -> hand the accumulator back
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    // teardown
    __syncwarp();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------- 2-CTA kernel
// cta_group::2: a cluster of two CTAs (one SM pair) computes one 256 x 256 tile.  Each CTA loads its own 128 rows
// of A and its own 128 columns of B (32 KB per stage instead of 48 KB), the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256) which reads both halves of B across the pair, and each CTA's epilogue drains
// its own 128 accumulator rows from its own TMEM.  Per SM the shared-memory traffic drops from 192 B/cycle
// (96 B/cycle TMA fill + 96 B/cycle operand reads at full MMA rate; the 1-CTA kernel measured 66-69 % tensor-pipe
// activity, i.e. the 128 B/cycle shared-memory limit) to 128 B/cycle.
// Two instantiations: 8 epilogue warps (12 warps, 6 smem stages) for the plain / atomic / multiply epilogues, and
// SIXTEEN epilogue warps (20 warps, 4 per TMEM lane quarter, 64 columns each, 5 stages) for the GELU epilogues, which
// are long latency chains (tcgen05.ld -> bias -> erf math -> two staged output streams): with 8 warps the issue slots
// were ~30 % busy and the epilogue, not the MMA, bounded FFN1 (39 % tensor-pipe active).
template <int kEpiWarps> struct Cfg2 {
    static constexpr int kThreads = (4 + kEpiWarps) * 32;
    static constexpr int kMaxSmem = 227 * 1024;
    // [operand ring: stages x 32 KB][epilogue staging: kEpiWarps x 4 KB][barriers: 256 B][column sums: colsum_floats x 4]
    static int stages(bool colsum) { return (kEpiWarps == 16 || colsum) ? 5 : 6; }
    static int smem_bytes(int stages, int colsum_floats) {
        return stages * (2 * 128 * BK * 2) + kEpiWarps * 4096 + 256 + colsum_floats * 4 + 1024;
    }
};
constexpr int k2ABytes = 128 * BK * 2;           // this CTA's half of the 256-row A tile
constexpr int k2BBytes = 128 * BK * 2;           // this CTA's half of the 256-column B tile
constexpr int k2StageBytes = k2ABytes + k2BBytes;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;      // clears the CTA-rank bit of a shared::cluster address -> leader CTA

template <int k2EpiWarps>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg2<k2EpiWarps>::kThreads, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux,
                         const GemmParams p) {
    const int k2Stages = p.stages;
    pdl_trigger();     // programmatic dependent launch (common.cuh): the next kernel's prologue may overlap this grid's tail
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t smem_base = ptx::smem_u32(smem);
    // [operand ring][epilogue staging: k2EpiWarps x 4 KB, 1024-byte aligned (swizzled TMA-store boxes)][barriers]
    const int kBarOff = k2Stages * k2StageBytes + k2EpiWarps * kStageBytesPerWarp;
    const uint32_t bar_base = smem_base + kBarOff;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (k2Stages + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * k2Stages + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * k2Stages + 2 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kBarOff + 8 * (2 * k2Stages + 4));
    uint8_t* epi_stage = smem + k2Stages * k2StageBytes;
    float* colacc = reinterpret_cast<float*>(smem + kBarOff + 256);   // [N] per-CTA column sums (p.colsum != null)
    if (p.colsum != nullptr)
        for (int i = threadIdx.x; i < p.N; i += Cfg2<k2EpiWarps>::kThreads) colacc[i] = 0.f;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();   // 0 = leader
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    constexpr int TM = 256, TN = 256;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
        if (p.tma_store) {
            ptx::prefetch_tensormap(&tmC);
            ptx::prefetch_tensormap(&tmAux);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < k2Stages; ++s) {
            ptx::mbar_init(full_bar(s), 1);    // leader's: one arrive.expect_tx covering both CTAs' TMA bytes
            ptx::mbar_init(empty_bar(s), 1);   // per CTA: multicast tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(tfull_bar(a), 1);                    // per CTA: multicast tcgen05.commit
            ptx::mbar_init(tempty_bar(a), 2 * k2EpiWarps);      // leader's: epilogue warps of both CTAs
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc_2cta<512>(ptx::smem_u32(tmem_slot));
    ptx::tc_fence_before();
    ptx::cluster_sync();        // peer barriers initialised, TMEM allocated in both CTAs
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                 // everything above touched shared / tensor memory only; operands and outputs are global

    const int total_kb = (p.K + BK - 1) / BK;
    const int mn_tiles = p.m_tiles * p.n_tiles;

    // one k-block of this CTA's operand halves -> stage ``stage``; bytes are credited to the leader's full barrier
    auto load_kblock = [&](int stage, int kb, int m0, int n0) {
        const uint32_t sa = smem_base + stage * k2StageBytes;
        const uint32_t sb = sa + k2ABytes;
        const uint32_t lead_full = full_bar(stage) & kPeerMask;
        if (rank == 0) ptx::mbar_expect_tx(full_bar(stage), 2 * k2StageBytes);
        if (!p.a_mn) {
            ptx::tma_load_2d_2cta(sa, &tmA, lead_full, kb * BK, m0);
        } else {
            ptx::tma_load_2d_2cta(sa, &tmA, lead_full, m0, kb * BK);
            ptx::tma_load_2d_2cta(sa + 8192, &tmA, lead_full, m0 + 64, kb * BK);
        }
        if (!p.b_mn) {
            ptx::tma_load_2d_2cta(sb, &tmB, lead_full, kb * BK, n0);
        } else {
            ptx::tma_load_2d_2cta(sb, &tmB, lead_full, n0, kb * BK);
            ptx::tma_load_2d_2cta(sb + 8192, &tmB, lead_full, n0 + 64, kb * BK);
        }
    };
    if (warp == 0 && !(p.dbg & 128)) {
        // ================================ TMA producer (both CTAs) ================================
        // whole warp, warp-uniform control flow, one elected lane issues (see the MMA issuer below)
        int stage = 0;
        uint32_t phase = 0;
        for (int t = cluster_id; t < p.total_tiles; t += num_clusters) {
            const int ks = t / mn_tiles;
            const int r = t - ks * mn_tiles;
            const int m_blk = r / p.n_tiles;   // n fastest: the tiles that share an A row-panel run concurrently
            const int n_blk = r - m_blk * p.n_tiles;
            const int kb0 = ks * p.kb_per_split;
            const int kb1 = min(total_kb, kb0 + p.kb_per_split);
            const int m0 = m_blk * TM + (int)rank * 128, n0 = n_blk * TN + (int)rank * 128;
            for (int kb = kb0; kb < kb1; ++kb) {
                ptx::mbar_wait(empty_bar(stage), phase ^ 1);
                if (ptx::elect_one()) load_kblock(stage, kb, m0, n0);
                __syncwarp();
                if (++stage == k2Stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 0 && lane == 0) {
        // bring-up (dbg bit 7): the same by lane 0 alone
        int stage = 0;
        uint32_t phase = 0;
        for (int t = cluster_id; t < p.total_tiles; t += num_clusters) {
            const int ks = t / mn_tiles;
            const int r = t - ks * mn_tiles;
            const int m_blk = r / p.n_tiles;
            const int n_blk = r - m_blk * p.n_tiles;
            const int kb0 = ks * p.kb_per_split;
            const int kb1 = min(total_kb, kb0 + p.kb_per_split);
            const int m0 = m_blk * TM + (int)rank * 128, n0 = n_blk * TN + (int)rank * 128;
            for (int kb = kb0; kb < kb1; ++kb) {
                ptx::mbar_wait(empty_bar(stage), phase ^ 1);
                load_kblock(stage, kb, m0, n0);
                if (++stage == k2Stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1 && rank == 0 && !(p.dbg & 128)) {
        // ================================ MMA issuer (leader CTA only) ================================
        // The whole warp runs the loop with warp-uniform control flow and one elected lane issues: the descriptors then
        // live in uniform registers (32-bit adds on the low half, the high half is loop-invariant) and the four MMAs of
        // a k-block go out back to back.  Issued from a lone lane of a diverged warp the same code costs ~35
        // instructions per tcgen05.mma (an ELECT / R2UR sequence and 64-bit descriptor arithmetic each time).
        const uint32_t idesc = make_idesc(TM, TN, p.a_mn, p.b_mn);
        const uint32_t hi_a = ((p.a_sbo >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);   // SBO | version 1 | SWIZZLE_128B
        const uint32_t hi_b = ((p.b_sbo >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
        const uint32_t lbo_a = ((p.a_lbo >> 4) & 0x3FFFu) << 16, lbo_b = ((p.b_lbo >> 4) & 0x3FFFu) << 16;
        const uint32_t step_a = p.a_step >> 4, step_b = p.b_step >> 4;   // tile addresses stay below 2^18: no carry
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = cluster_id; t < p.total_tiles; t += num_clusters) {
            const int ks = t / mn_tiles;
            const int kb0 = ks * p.kb_per_split;
            const int kb1 = min(total_kb, kb0 + p.kb_per_split);
            ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1);
            ptx::tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * TN;
            for (int kb = kb0; kb < kb1; ++kb) {
                ptx::mbar_wait(full_bar(stage), phase);
                ptx::tc_fence_after();
                const uint32_t sa = smem_base + stage * k2StageBytes;
                const uint32_t a16 = ((sa & 0x3FFFFu) >> 4) | lbo_a;
                const uint32_t b16 = (((sa + k2ABytes) & 0x3FFFFu) >> 4) | lbo_b;
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        ptx::umma_bf16_2cta(tmem_d, ptx::umma_desc_from_halves(a16 + k * step_a, hi_a),
                                            ptx::umma_desc_from_halves(b16 + k * step_b, hi_b), idesc,
                                            (kb > kb0 || k > 0) ? 1u : 0u);
                    ptx::umma_commit_2cta(empty_bar(stage));   // both CTAs' slots are free once these MMAs retire
                }
                __syncwarp();
                if (++stage == k2Stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (ptx::elect_one()) ptx::umma_commit_2cta(tfull_bar(acc));   // both CTAs' epilogues may drain their half
            __syncwarp();
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        // bring-up (dbg bit 7 / MMB_GEMM_ISSUE=lane): the same issued by lane 0 alone, for A/B runs
        const uint32_t idesc = make_idesc(TM, TN, p.a_mn, p.b_mn);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = cluster_id; t < p.total_tiles; t += num_clusters) {
            const int ks = t / mn_tiles;
            const int kb0 = ks * p.kb_per_split;
            const int kb1 = min(total_kb, kb0 + p.kb_per_split);
            ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1);
            ptx::tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * TN;
            for (int kb = kb0; kb < kb1; ++kb) {
                ptx::mbar_wait(full_bar(stage), phase);
                ptx::tc_fence_after();
                const uint32_t sa = smem_base + stage * k2StageBytes;
                const uint32_t sb = sa + k2ABytes;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t da = ptx::umma_desc_sw128(sa + k * p.a_step, p.a_lbo, p.a_sbo);
                    const uint64_t db = ptx::umma_desc_sw128(sb + k * p.b_step, p.b_lbo, p.b_sbo);
                    ptx::umma_bf16_2cta(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                }
                ptx::umma_commit_2cta(empty_bar(stage));
                if (++stage == k2Stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            ptx::umma_commit_2cta(tfull_bar(acc));
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ================================ epilogue (both CTAs, own 128 rows) ================================
        const int ew = warp - kEpiWarp0;
        const int quarter = warp & 3;
        const int half = ew >> 2;          // column group 0..3 of the tile
        const uint32_t stage_buf = ptx::smem_u32(epi_stage) + ew * kStageBytesPerWarp;
        constexpr int kColsPerWarp = TN / (k2EpiWarps / 4);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = cluster_id; t < p.total_tiles; t += num_clusters) {
            const int ks = t / mn_tiles;
            const int r = t - ks * mn_tiles;
            const int m_blk = r / p.n_tiles;   // n fastest: the tiles that share an A row-panel run concurrently
            const int n_blk = r - m_blk * p.n_tiles;
            const int row_base = m_blk * TM + (int)rank * 128 + quarter * 32;
            // MUL_AUX: this thread's 128 aux values (its row, the warp's column half) are fetched BEFORE waiting for the
            // accumulator, so the global-load latency hides under the tile's MMAs instead of stalling every chunk
            const int colw = n_blk * TN + half * kColsPerWarp;
            // padding rows (mmb_gemm_args.row_live): a 32-row slice without a live row needs no epilogue work
            bool slice_live = true;
            if (p.row_live != nullptr) {
                const int row = row_base + lane;
                slice_live = __any_sync(0xffffffffu, row < p.M && __ldg(p.row_live + row) != 0);
            }
            const bool aux_pref = p.epilogue == MMB_EPI_MUL_AUX_BF16 && colw + kColsPerWarp <= p.N && !(p.dbg & 48);
            // (32-byte loads: a lane walks along ITS row, so a 16-byte load would use half of every sector it touches — ncu
            // had the L1 at 78 % throughput with 16-byte loads, the busiest unit of this GEMM)
            uint32_t auxr[kColsPerWarp / 16][8];
            if (aux_pref && slice_live) {
                const int row = row_base + lane;
                const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(p.aux) + (size_t)(row < p.M ? row : 0) * p.ldaux + colw;
                if ((reinterpret_cast<uintptr_t>(src) & 31) == 0) {
#pragma unroll
                    for (int j = 0; j < kColsPerWarp / 16; ++j) ptx::ldg256_nc(src + 16 * j, auxr[j]);
                } else {                                   // ldaux / base not 32-byte aligned: two 16-byte loads
#pragma unroll
                    for (int j = 0; j < kColsPerWarp / 16; ++j) {
                        const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(src + 16 * j));
                        const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(src + 16 * j + 8));
                        auxr[j][0] = q0.x; auxr[j][1] = q0.y; auxr[j][2] = q0.z; auxr[j][3] = q0.w;
                        auxr[j][4] = q1.x; auxr[j][5] = q1.y; auxr[j][6] = q1.z; auxr[j][7] = q1.w;
                    }
                }
            }
            ptx::mbar_wait(tfull_bar(acc), acc_phase);
            ptx::tc_fence_after();
            if constexpr (kColsPerWarp == kCeGroupCols) {
                if (p.epilogue == MMB_EPI_CE_STATS && colw < p.N)
                    ce_stats_warp(p, tmem_base + acc * TN + half * kColsPerWarp + ((uint32_t)(quarter * 32) << 16), row_base, lane, colw);
            }
            const uint32_t tacc = tmem_base + acc * TN + half * kColsPerWarp + ((uint32_t)(quarter * 32) << 16);
            const bool has_aux = p.epilogue == MMB_EPI_GELU_GRAD_BF16 || (p.epilogue == MMB_EPI_GELU_BF16 && p.aux != nullptr);
            // 8 warps: one output stream (multiply needs its prefetched aux values); 16 warps: no multiply epilogue
            const bool tma_epi = p.tma_store && !(p.dbg & 48) &&
                                 (k2EpiWarps == 8 ? (!has_aux && (p.epilogue != MMB_EPI_MUL_AUX_BF16 || aux_pref))
                                                  : p.epilogue != MMB_EPI_MUL_AUX_BF16);
            bool done = p.epilogue == MMB_EPI_CE_STATS;
            if (tma_epi && colw < p.N) {
                done = true;
                if constexpr (k2EpiWarps == 8) {
                    // 128-column strip = two 64-column boxes (plain / ReLU / multiply epilogues; GELU only in A/B runs)
#pragma unroll
                    for (int g = 0; g < kColsPerWarp / 64; ++g) {     // unrolled: auxr[] stays in registers
                        const int colg = colw + g * 64;
                        if (colg >= p.N) break;
                        if (!slice_live) {                // an all-padding slice: unwritten (plain store) or zeros (multiply), nothing read
                            if (p.dead_zeroed || p.epilogue != MMB_EPI_MUL_AUX_BF16) continue;  // (zero already: earlier launch of this step)
                            stage_acquire(lane);
#pragma unroll
                            for (int j = 0; j < 8; ++j) sts128(stage_buf + stage_off128(lane, j), 0u, 0u, 0u, 0u);
                            stage_release(&tmC, stage_buf, nullptr, 0, colg, row_base, lane);
                            continue;
                        }
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int col0 = colg + 32 * h;
                            uint32_t raw[32];
                            ptx::tmem_ld_32x32(tacc + g * 64 + h * 32, raw);
                            ptx::tmem_ld_wait();
                            float v[32], w[32];
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                            if (p.alpha != 1.0f) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
                            }
                            if (p.epilogue == MMB_EPI_MUL_AUX_BF16) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) {      // 32 columns = two 32-byte vectors of 8 bf16 pairs
                                    const float2 f = unpack_bf16x2(auxr[4 * g + 2 * h + (j >> 3)][j & 7]);
                                    v[2 * j] *= f.x;
                                    v[2 * j + 1] *= f.y;
                                }
                            } else {
                                bias_add_chunk(p, v, col0);
                                activate_chunk(p.epilogue, false, v, w);
                            }
                            if (h == 0) stage_acquire(lane);
                            stage_row128(stage_buf, v, lane, h);
                        }
                        if (p.colsum != nullptr) {
                            // column sums of the staged (bf16-rounded) tile: lane l owns columns 2l, 2l+1 of the 64; rows
                            // beyond M hold padding (bias-only values) and are left out
                            __syncwarp();
                            const int nrow = min(32, p.M - row_base);
                            float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
                            const uint32_t wb = stage_buf + (uint32_t)((lane & 3) * 4);
                            if (nrow == 32) {
#pragma unroll
                                for (int r = 0; r < 32; r += 4) {
                                    a0 = add2(a0, unpack_bf16x2(lds32(wb + stage_off128(r, lane >> 2))));
                                    a1 = add2(a1, unpack_bf16x2(lds32(wb + stage_off128(r + 1, lane >> 2))));
                                    a2 = add2(a2, unpack_bf16x2(lds32(wb + stage_off128(r + 2, lane >> 2))));
                                    a3 = add2(a3, unpack_bf16x2(lds32(wb + stage_off128(r + 3, lane >> 2))));
                                }
                            } else {
                                for (int r = 0; r < nrow; ++r) a0 = add2(a0, unpack_bf16x2(lds32(wb + stage_off128(r, lane >> 2))));
                            }
                            const float2 t = add2(add2(a0, a1), add2(a2, a3));
                            const int c = colg + 2 * lane;
                            if (c < p.N) atomicAdd(colacc + c, t.x);
                            if (c + 1 < p.N) atomicAdd(colacc + c + 1, t.y);
                        }
                        stage_release(&tmC, stage_buf, nullptr, 0, colg, row_base, lane);
                    }
                } else {
                    // 64-column strip = two 32-column boxes per stream (GELU epilogues: C and, in training, the aux stream)
#pragma unroll 1
                    for (int c = 0; c < kColsPerWarp; c += 32) {
                        const int col0 = colw + c;
                        if (col0 >= p.N || !slice_live) break;     // (an all-padding slice stays unwritten)
                        uint32_t raw[32];
                        ptx::tmem_ld_32x32(tacc + c, raw);
                        ptx::tmem_ld_wait();
                        float v[32], w[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                        if (p.alpha != 1.0f) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
                        }
                        bias_add_chunk(p, v, col0);
                        stage_acquire(lane);
                        if (p.epilogue == MMB_EPI_GELU_GRAD_BF16 || p.epilogue == MMB_EPI_GELU_BF16) {
                            // 8 columns at a time -> one 16-byte segment of each staging row: only 4 pairs of results are live
                            // (this kernel has 96 registers per thread; 32 + 32 fp32 results at once left the erf math no ILP)
                            const bool grad = p.epilogue == MMB_EPI_GELU_GRAD_BF16;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                uint32_t pg[4], pd[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float2 x = make_float2(v[8 * j + 2 * i], v[8 * j + 2 * i + 1]);
                                    float2 g, dg;
                                    gelu_fast2(x, g, dg);
                                    pg[i] = pack_bf16x2(g.x, g.y);
                                    pd[i] = grad ? pack_bf16x2(dg.x, dg.y) : pack_bf16x2(x.x, x.y);     // gelu' | pre-activation
                                }
                                sts128(stage_buf + stage_off(lane, j), pg[0], pg[1], pg[2], pg[3]);
                                if (has_aux) sts128(stage_buf + 2048 + stage_off(lane, j), pd[0], pd[1], pd[2], pd[3]);
                            }
                        } else {
                            activate_chunk(p.epilogue, has_aux, v, w);
                            stage_row(stage_buf, v, lane);
                            if (has_aux) stage_row(stage_buf + 2048, w, lane);
                        }
                        stage_release(&tmC, stage_buf, has_aux ? &tmAux : nullptr, stage_buf + 2048, col0, row_base, lane);
                    }
                }
            } else if (p.tma_store) {
                stage_acquire(lane);       // the st.global paths below reuse the staging tile
            }
#pragma unroll 1
            for (int c = done ? kColsPerWarp : 0; c < kColsPerWarp; c += 32) {
                const int col0 = n_blk * TN + half * kColsPerWarp + c;
                if (col0 >= p.N) break;
                if (p.dbg & 32) continue;
                uint32_t raw[32];
                const uint32_t taddr = tmem_base + acc * TN + half * kColsPerWarp + c + ((uint32_t)(quarter * 32) << 16);
                ptx::tmem_ld_32x32(taddr, raw);
                ptx::tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                if (p.alpha != 1.0f) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
                }
                if (!(p.dbg & 16)) epilogue_chunk(p, v, row_base, lane, col0, stage_buf);
                __syncwarp();
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(tempty_bar(acc) & kPeerMask);   // always the leader's barrier
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    // teardown: nobody may exit while the peer can still touch its shared memory / barriers / TMEM
    if (warp >= kEpiWarp0 && lane == 0 && p.tma_store) ptx::bulk_wait_read<0>();   // staging tiles outlive their last store
    __syncwarp();
    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_2cta<512>(tmem_base);
    }
    if (p.colsum != nullptr)        // every epilogue warp's shared-memory reductions are behind the cluster barrier
        for (int i = threadIdx.x; i < p.N; i += Cfg2<k2EpiWarps>::kThreads) {
            const float v = colacc[i];
            if (v != 0.f) atomicAdd(p.colsum + i, v);
        }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    });
    return fn;
}

struct TmapKey {
    const void* ptr;
    uint64_t d0, d1, ld;
    uint32_t b0, b1;   // b1 carries the swizzle mode in its top bit (box rows are <= 256)
    bool operator==(const TmapKey& o) const {
        return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && ld == o.ld && b0 == o.b0 && b1 == o.b1;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        uint64_t h = reinterpret_cast<uint64_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
        h ^= (k.d0 + 0x7F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
        h ^= (k.d1 + 0x94D049BBull) * 0x94D049BB133111EBull;
        h ^= (k.ld << 17) ^ ((uint64_t)k.b0 << 40) ^ ((uint64_t)k.b1 << 52);
        return (size_t)(h ^ (h >> 29));
    }
};

// bf16 2-D tensor map, SWIZZLE_128B: dims {d0 (contiguous), d1 (rows)}, row stride ld elements, box {b0, b1}.
// Cached: the encode is pure host work but the engine issues ~150 GEMMs per step over a fixed buffer set.
static int make_tmap_bf16_sw(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1,
                             bool sw64);
int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1) {
    return make_tmap_bf16_sw(out, ptr, d0, d1, ld, b0, b1, false);
}
// sw64: SWIZZLE_64B (64-byte box rows: the 16-warp epilogue's store tiles) instead of SWIZZLE_128B
static int make_tmap_bf16_sw(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1,
                             bool sw64) {
    static std::mutex mu;
    static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
    TmapKey key{ptr, d0, d1, ld, b0, b1 | (sw64 ? 0x80000000u : 0u)};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return MMB_OK;
        }
    }
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
        set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
        return MMB_ECUDA;
    }
    cuuint64_t dims[2] = {d0, d1};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {b0, b1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed (%d) ptr=%p dims=(%llu,%llu) ld=%llu box=(%u,%u)", (int)r, ptr,
                       (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)ld, b0, b1);
        return MMB_ECUDA;
    }
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, *out);
    return MMB_OK;
}

static void fill_common(GemmParams& p, const mmb_gemm_args* a, int tile_m, int tile_n) {
    p.C = a->C;
    p.aux = a->aux;
    p.aux2 = a->aux2;
    p.bias = a->bias;
    p.ldc = a->ldc;
    p.ldaux = a->ldaux;
    p.M = a->M;
    p.N = a->N;
    p.K = a->K;
    p.a_mn = a->a_major == MMB_MAJOR_MN;
    p.b_mn = a->b_major == MMB_MAJOR_MN;
    p.epilogue = a->epilogue;
    const int total_kb = (a->K + BK - 1) / BK;
    int split = a->split_k < 1 ? 1 : a->split_k;
    if (split > total_kb) split = total_kb;
    p.kb_per_split = (total_kb + split - 1) / split;
    p.split_k = (total_kb + p.kb_per_split - 1) / p.kb_per_split;  // every split non-empty
    p.m_tiles = (a->M + tile_m - 1) / tile_m;
    p.n_tiles = (a->N + tile_n - 1) / tile_n;
    p.total_tiles = p.m_tiles * p.n_tiles * p.split_k;
    p.alpha = a->alpha;
    p.tma_store = 0;
    p.stages = 0;
    p.colsum = nullptr;
    p.row_live = nullptr;
    p.dead_zeroed = 0;
    p.dbg = a->dbg_flags;
    static const int lane_issue = [] {                       // MMB_GEMM_ISSUE=lane: single-lane MMA issue (A/B runs)
        const char* e = getenv("MMB_GEMM_ISSUE");
        return (e != nullptr && e[0] == 'l') ? 128 : 0;
    }();
    p.dbg |= lane_issue;
    // K-major SW128: 8-row groups 1024 B apart (SBO), LBO unused; 32 B per UMMA_K inside the swizzled row.
    // MN-major SW128: 64-element MN blocks 8192 B apart (LBO), 8-k groups 1024 B apart (SBO); two k-groups
    // (2048 B) per UMMA_K.  dbg_flags bit 1 swaps LBO/SBO of MN-major operands, bit 2 sets K-major LBO = 16 B.
    const bool swap = (a->dbg_flags & 2) != 0;
    const uint32_t k_lbo = (a->dbg_flags & 4) ? 16u : 0u;
    p.a_lbo = p.a_mn ? (swap ? 1024u : 8192u) : k_lbo;
    p.a_sbo = p.a_mn ? (swap ? 8192u : 1024u) : 1024u;
    p.a_step = p.a_mn ? 2048u : 32u;
    p.b_lbo = p.b_mn ? (swap ? 1024u : 8192u) : k_lbo;
    p.b_sbo = p.b_mn ? (swap ? 8192u : 1024u) : 1024u;
    p.b_step = p.b_mn ? 2048u : 32u;
}

template <int BN>
static int launch_gemm(const mmb_gemm_args* a, cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    CUtensorMap tmA, tmB;
    int rc;
    if (a->a_major == MMB_MAJOR_K)
        rc = make_tmap_bf16(&tmA, a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda, BK, BM);
    else
        rc = make_tmap_bf16(&tmA, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, 64, BK);
    if (rc != MMB_OK) return rc;
    if (a->b_major == MMB_MAJOR_K)
        rc = make_tmap_bf16(&tmB, a->B, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb, BK, BN);
    else
        rc = make_tmap_bf16(&tmB, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, 64, BK);
    if (rc != MMB_OK) return rc;
    GemmParams p;
    fill_common(p, a, BM, BN);
    MMB_ENSURE_SMEM(Cfg::kSmemBytes, gemm_tcgen05_kernel<BN>);
    const int grid = p.total_tiles < persistent_sms() ? p.total_tiles : persistent_sms();
    launch_pdl(gemm_tcgen05_kernel<BN>, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, stream, tmA, tmB, p);
    return check_launch("gemm_tcgen05_kernel");
}

// Planes of 16-byte records [planes][records], box = {nbox records, 1 plane}, no swizzle: the per-(head, row) records of
// the attention backward (attn.cu) copied next to the operand tiles.  Encoded as pairs of 64-bit elements so that a
// 128-record box stays within the 256-element box limit and every record offset is a 16-byte aligned address.
int make_tmap_rec16(CUtensorMap* out, const void* ptr, uint64_t records, uint64_t planes, uint32_t nbox) {
    const uint64_t d0 = 2 * records, d1 = planes, ld = 2 * records;
    const uint32_t b0 = 2 * nbox;
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
        set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
        return MMB_ECUDA;
    }
    cuuint64_t dims[2] = {d0, d1};
    cuuint64_t strides[1] = {ld * 8};
    cuuint32_t box[2] = {b0, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled (16-byte records) failed (%d) ptr=%p dims=(%llu,%llu) ld=%llu box=%u", (int)r, ptr,
                       (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)ld, b0);
        return MMB_ECUDA;
    }
    return MMB_OK;
}

// mmb_gemm_args.colsum where the epilogue cannot take it: a mmb_colsum_bf16 launch over the finished C
static int gemm_colsum_fallback(const mmb_gemm_args* a, cudaStream_t stream) {
    MMB_REQUIRE(a->N % 8 == 0 && a->ldc % 8 == 0, "mmb_gemm: colsum needs N %% 8 == 0 on this path (N=%d)", a->N);
    mmb_colsum_args c;
    c.X = a->C;
    c.out = a->colsum;
    c.ld = a->ldc;
    c.M = a->M;
    c.N = a->N;
    c.row_list = nullptr;
    return mmb_colsum_bf16(&c, stream);
}

// CTA-pair kernel: 256 x 256 tiles, one cluster of 2 per SM pair.
static int launch_gemm_2cta(const mmb_gemm_args* a, cudaStream_t stream) {
    CUtensorMap tmA, tmB;
    int rc;
    if (a->a_major == MMB_MAJOR_K)
        rc = make_tmap_bf16(&tmA, a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda, BK, 128);
    else
        rc = make_tmap_bf16(&tmA, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, 64, BK);
    if (rc != MMB_OK) return rc;
    if (a->b_major == MMB_MAJOR_K)
        rc = make_tmap_bf16(&tmB, a->B, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb, BK, 128);
    else
        rc = make_tmap_bf16(&tmB, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, 64, BK);
    if (rc != MMB_OK) return rc;
    GemmParams p;
    fill_common(p, a, 256, 256);
    MMB_ENSURE_SMEM(Cfg2<8>::kMaxSmem, gemm_tcgen05_2cta_kernel<8>);
    MMB_ENSURE_SMEM(Cfg2<16>::kMaxSmem, gemm_tcgen05_2cta_kernel<16>);
    const int pairs = persistent_sms() / 2;
    const int clusters = p.total_tiles < pairs ? p.total_tiles : pairs;
    // GELU epilogues are long latency chains: 16 epilogue warps; everything else: 8 (dbg bit 6 flips the choice, for A/B runs)
    bool wide = a->epilogue == MMB_EPI_GELU_BF16 || a->epilogue == MMB_EPI_GELU_GRAD_BF16;
    if (a->dbg_flags & 64) wide = !wide;
    // bf16 outputs leave through TMA stores (dbg bit 8 / MMB_GEMM_STORE=stg keeps the st.global epilogue, for A/B runs)
    static const bool stg_env = [] {
        const char* e = getenv("MMB_GEMM_STORE");
        return e != nullptr && e[0] == 's';
    }();
    const bool bf16_out = a->epilogue == MMB_EPI_STORE_BF16 || a->epilogue == MMB_EPI_GELU_BF16 ||
                          a->epilogue == MMB_EPI_RELU_BF16 || a->epilogue == MMB_EPI_GELU_GRAD_BF16 ||
                          a->epilogue == MMB_EPI_MUL_AUX_BF16;
    p.tma_store = (bf16_out && !stg_env && !(a->dbg_flags & 256)) ? 1 : 0;
    CUtensorMap tmC = tmA, tmAux = tmA;      // placeholders when unused (never dereferenced)
    if (p.tma_store) {
        const bool aux_out = a->epilogue == MMB_EPI_GELU_GRAD_BF16 || (a->epilogue == MMB_EPI_GELU_BF16 && a->aux != nullptr);
        rc = make_tmap_bf16_sw(&tmC, a->C, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldc, wide ? 32 : 64, 32, wide);
        if (rc == MMB_OK && aux_out) {
            MMB_REQUIRE(((uintptr_t)a->aux % 16) == 0, "mmb_gemm: aux must be 16-byte aligned");
            rc = make_tmap_bf16_sw(&tmAux, a->aux, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldaux, wide ? 32 : 64, 32, wide);
        }
        if (rc != MMB_OK) return rc;
    }
    // column sums ride on the 8-warp TMA-store epilogue (one output stream); everything else: mmb_gemm adds a launch
    static const int stages_env = [] {
        const char* e = getenv("MMB_GEMM_STAGES");
        return e != nullptr ? atoi(e) : 0;
    }();
    // (the multiply epilogue leaves the TMA-store path on a ragged last column strip: N must be a multiple of 128 there)
    const bool fused_colsum = a->colsum != nullptr && p.tma_store && !wide && (size_t)a->N * 4 <= 28 * 1024 &&
                              (a->epilogue != MMB_EPI_MUL_AUX_BF16 || a->N % 128 == 0) && !(a->dbg_flags & 48);
    p.colsum = fused_colsum ? a->colsum : nullptr;
    // padding-row hint: the GELU epilogues on their 16-warp TMA path and the plain store on the 8-warp TMA path (slices stay
    // unwritten), MUL_AUX on the 8-warp TMA path with whole 128-column strips (slices are zero-filled); every other
    // combination computes all rows
    if (a->row_live != nullptr && p.tma_store &&
        ((wide && (a->epilogue == MMB_EPI_GELU_GRAD_BF16 || a->epilogue == MMB_EPI_GELU_BF16)) ||
         (!wide && a->epilogue == MMB_EPI_MUL_AUX_BF16 && a->N % 128 == 0) ||
         (!wide && a->epilogue == MMB_EPI_STORE_BF16 && a->colsum == nullptr)))
        p.row_live = a->row_live;
    p.dead_zeroed = (p.row_live != nullptr && a->dead_rows_zeroed) ? 1 : 0;
    if (wide) {
        p.stages = Cfg2<16>::stages(false);
        launch_pdl(gemm_tcgen05_2cta_kernel<16>, dim3(2 * clusters), dim3(Cfg2<16>::kThreads), Cfg2<16>::smem_bytes(p.stages, 0), stream,
                   tmA, tmB, tmC, tmAux, p);
    } else {
        p.stages = Cfg2<8>::stages(fused_colsum);
        if (!fused_colsum && (stages_env == 5 || stages_env == 6)) p.stages = stages_env;      // A/B runs
        launch_pdl(gemm_tcgen05_2cta_kernel<8>, dim3(2 * clusters), dim3(Cfg2<8>::kThreads),
                   Cfg2<8>::smem_bytes(p.stages, fused_colsum ? a->N : 0), stream, tmA, tmB, tmC, tmAux, p);
    }
    if (a->colsum != nullptr && !fused_colsum) {
        int rc2 = check_launch("gemm_tcgen05_2cta_kernel");
        if (rc2 != MMB_OK) return rc2;
        return gemm_colsum_fallback(a, stream);
    }
    return check_launch("gemm_tcgen05_2cta_kernel");
}

}  // namespace mmb

extern "C" size_t mmb_ce_stats_floats(int M, int N) {
    if (M <= 0 || N <= 0) return 0;
    const size_t G = ((size_t)N + mmb::kCeGroupCols - 1) / mmb::kCeGroupCols;
    return (2 * G + 2) * (size_t)M;
}

extern "C" int mmb_gemm(const mmb_gemm_args* a, void* stream) {
    using namespace mmb;
    MMB_REQUIRE(a != nullptr, "mmb_gemm: null args");
    MMB_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "mmb_gemm: empty problem M=%d N=%d K=%d", a->M, a->N, a->K);
    MMB_REQUIRE((a->lda % 8) == 0 && (a->ldb % 8) == 0, "mmb_gemm: lda/ldb must be multiples of 8 (got %lld, %lld)",
                (long long)a->lda, (long long)a->ldb);
    const bool ce = a->epilogue == MMB_EPI_CE_STATS;
    MMB_REQUIRE(a->A && a->B && (a->C || ce), "mmb_gemm: null operand");
    const bool f32_out = a->epilogue == MMB_EPI_STORE_F32 || a->epilogue == MMB_EPI_ATOMIC_ADD_F32;
    MMB_REQUIRE((a->ldc % (f32_out ? 4 : 8)) == 0, "mmb_gemm: ldc=%lld misaligned", (long long)a->ldc);
    MMB_REQUIRE(((uintptr_t)a->A % 16) == 0 && ((uintptr_t)a->B % 16) == 0 && ((uintptr_t)a->C % 16) == 0,
                "mmb_gemm: operands must be 16-byte aligned");
    MMB_REQUIRE(a->epilogue >= 0 && a->epilogue <= MMB_EPI_CE_STATS, "mmb_gemm: bad epilogue %d", a->epilogue);
    if (ce) {
        MMB_REQUIRE(a->aux && a->aux2, "mmb_gemm: MMB_EPI_CE_STATS needs aux (row labels) and aux2 (stats workspace)");
        MMB_REQUIRE(a->N > 128, "mmb_gemm: MMB_EPI_CE_STATS needs N > 128 (128-column statistic groups)");
        MMB_REQUIRE(a->split_k <= 1 && !(a->dbg_flags & (1 | 64)), "mmb_gemm: MMB_EPI_CE_STATS: no split-K / narrow tiles");
    }
    MMB_REQUIRE(a->split_k <= 1 || a->epilogue == MMB_EPI_ATOMIC_ADD_F32, "mmb_gemm: split_k needs ATOMIC_ADD_F32");
    if (a->epilogue == MMB_EPI_DGELU_BF16 || a->epilogue == MMB_EPI_GELU_GRAD_BF16 || a->epilogue == MMB_EPI_MUL_AUX_BF16)
        MMB_REQUIRE(a->aux != nullptr && (a->ldaux % 8) == 0, "mmb_gemm: this epilogue needs aux with ldaux %% 8 == 0");
    if (a->epilogue == MMB_EPI_GELU_BF16 && a->aux) MMB_REQUIRE((a->ldaux % 8) == 0, "mmb_gemm: ldaux %% 8 != 0");
    const int min_lda = a->a_major == MMB_MAJOR_K ? a->K : a->M;
    const int min_ldb = a->b_major == MMB_MAJOR_K ? a->K : a->N;
    MMB_REQUIRE(a->lda >= min_lda && a->ldb >= min_ldb && (a->ldc >= a->N || (ce && !a->C)), "mmb_gemm: leading dimension too small");
    // Dispatch: the CTA-pair kernel (256 x 256 tiles) whenever the problem has more than one 128-row tile and
    // more than 128 columns; otherwise the single-CTA kernel (128 x 256, or 128 x 128 for N <= 128).
    // dbg_flags: bit 0 forces the 128 x 128 tile, bit 3 forces the single-CTA 128 x 256 kernel.
    if (a->colsum != nullptr)
        MMB_REQUIRE(a->epilogue == MMB_EPI_STORE_BF16 || a->epilogue == MMB_EPI_RELU_BF16 || a->epilogue == MMB_EPI_MUL_AUX_BF16 ||
                        a->epilogue == MMB_EPI_GELU_BF16 || a->epilogue == MMB_EPI_DGELU_BF16,
                    "mmb_gemm: colsum needs a single bf16 output (epilogue %d)", a->epilogue);
    const bool one_cta = (a->dbg_flags & (1 | 8)) || a->N <= 128 || a->M <= 128;
    if (!one_cta) return launch_gemm_2cta(a, (cudaStream_t)stream);
    int rc;
    if ((a->dbg_flags & 1) || (a->N <= 128 && !(a->dbg_flags & 8))) rc = launch_gemm<128>(a, (cudaStream_t)stream);
    else rc = launch_gemm<256>(a, (cudaStream_t)stream);
    if (rc == MMB_OK && a->colsum != nullptr) rc = gemm_colsum_fallback(a, (cudaStream_t)stream);
    return rc;
}
