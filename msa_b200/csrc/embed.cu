// Packed-batch preparation and the fused embedding kernels (forward / backward).
//
// Packed row order: the reference's three passes become one variable-length batch of 3B sequences,
//   pass 0 (text)          rows [0, B*T)                       sequence q = b,       S = T
//   pass 1 (text ⊕ visual) rows [B*T, B*T + B*(T+Lv))          sequence q = B + b,   S = T + Lv
//   pass 2 (text ⊕ speech) rows [.., .. + B*(T+La))            sequence q = 2B + b,  S = T + La
// (MMBertForPretraining.py:402-404).  Every later kernel works on these M = B*(3T+Lv+La) rows.
//
// Replaces
//   BertEmbeddings.forward        modeling_bert.py:103-112  word + token_type + position, LayerNorm(eps), dropout(.1)
//   JointEmbeddings.forward       MMBertEmbedding.py:61-70  relu(W·f32(frames)+b), cat, LayerNorm(1e-5), dropout(.5)
//   get_extended_attention_mask   MMBertForPretraining.py:57-154,246-250   (1 - mask) * -10000, frame mask = feature 0
// and their autograd backward (embedding scatter-add with padding_idx 0, LN / dropout / relu / Linear backward).
#include "common.cuh"
#include "ptx.cuh"
#include "rowops.cuh"

namespace mmb {

struct PackDims {
    int B, T, L1, L2;  // L1 = visual frames, L2 = speech frames
    __host__ __device__ int S(int pass) const { return pass == 0 ? T : (pass == 1 ? T + L1 : T + L2); }
    __host__ __device__ int base(int pass) const { return pass == 0 ? 0 : (pass == 1 ? B * T : B * T + B * (T + L1)); }
    __host__ __device__ int rows() const { return B * (3 * T + L1 + L2); }
    __host__ __device__ int frame_rows() const { return B * (L1 + L2); }
};

struct RowCoord {
    int pass, b, s;  // s = position inside the sequence (s >= T: frame s - T)
};
__device__ __forceinline__ RowCoord locate(const PackDims& d, int row) {
    RowCoord c;
    c.pass = row < d.base(1) ? 0 : (row < d.base(2) ? 1 : 2);
    const int local = row - d.base(c.pass), S = d.S(c.pass);
    c.b = local / S;
    c.s = local - c.b * S;
    return c;
}

__device__ __forceinline__ float load_as_float(const void* p, int dtype, int64_t idx) {
    switch (dtype) {
        case MMB_DT_F32: return reinterpret_cast<const float*>(p)[idx];
        case MMB_DT_F64: return (float)reinterpret_cast<const double*>(p)[idx];
        case MMB_DT_I64: return (float)reinterpret_cast<const long long*>(p)[idx];
        case MMB_DT_I32: return (float)reinterpret_cast<const int*>(p)[idx];
        default: return (float)reinterpret_cast<const unsigned char*>(p)[idx];
    }
}

// ------------------------------------------------------------------ pack_prepare
struct PrepParams {
    PackDims d;
    const void* mask_text[3];   // [B,T] per pass
    int mask_text_dt[3];
    const void* mask_frame[2];  // [B,L,D] for pass 1, 2 (feature 0 is read)
    int mask_frame_dt[2];
    int mask_frame_stride[2];   // elements per frame in mask_frame (D, or 1 for a [B,L] mask)
    const long long* labels[3]; // [B, S(pass)]
    float* keybias;             // [rows]
    int* cu_seqlens;            // [3B + 1]
    int* label_count;           // [3] (+ [3]: error count)
    int* kv_end;                // [3B] or null
    int* row_label;             // [rows] or null
    int vocab;                  // > 0: labels outside [0, vocab) are errors (counted in label_count[3], treated as -100)
};

__global__ void pack_prepare_kernel(const PrepParams p) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int rows = p.d.rows();
    int cnt[3] = {0, 0, 0};
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < rows; row += gridDim.x * blockDim.x) {
        const RowCoord c = locate(p.d, row);
        float m;
        if (c.s < p.d.T) {
            m = load_as_float(p.mask_text[c.pass], p.mask_text_dt[c.pass], (int64_t)c.b * p.d.T + c.s);
        } else {
            const int L = c.pass == 1 ? p.d.L1 : p.d.L2, D = p.mask_frame_stride[c.pass - 1];
            m = load_as_float(p.mask_frame[c.pass - 1], p.mask_frame_dt[c.pass - 1], ((int64_t)c.b * L + (c.s - p.d.T)) * D);
        }
        p.keybias[row] = (1.0f - m) * -10000.0f;
        if (p.kv_end != nullptr && m >= 0.5f) atomicMax(p.kv_end + c.pass * p.d.B + c.b, c.s + 1);
        long long lab = p.labels[c.pass] != nullptr ? p.labels[c.pass][(int64_t)c.b * p.d.S(c.pass) + c.s] : -100;
        if (lab != -100 && p.vocab > 0 && (lab < 0 || lab >= p.vocab)) {
            atomicAdd(p.label_count + 3, 1);
            lab = -100;
        }
        if (lab != -100) cnt[c.pass]++;
        if (p.row_label != nullptr) p.row_label[row] = (int)lab;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int v = cnt[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(p.label_count + i, v);
    }
    const int nseq = 3 * p.d.B;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q <= nseq; q += gridDim.x * blockDim.x) {
        const int pass = q < p.d.B ? 0 : (q < 2 * p.d.B ? 1 : 2);
        p.cu_seqlens[q] = q == nseq ? rows : p.d.base(pass) + (q - pass * p.d.B) * p.d.S(pass);
    }
}

// ------------------------------------------------------------------ embedding forward
struct EmbedParams {
    PackDims d;
    int H, V, max_pos;
    int frame_dim[2];
    const long long* ids[3];       // [B,T] per pass
    const long long* token_type;   // [B,T] for pass 0 (joint passes use type 0, :223)
    const void* frames[2];         // [B,L,D]
    int frames_dt[2];
    const float* word;             // [V,H]
    const float* pos;              // [max_pos,H]
    const float* type;             // [2,H]
    const float* ln1_g; const float* ln1_b; float eps1;
    const float* ln2_g; const float* ln2_b; float eps2;
    const float* wT[2];            // [D,H] transposed projection weights (Wv, Ws)
    const float* wb[2];            // [H]
    uint32_t thresh1, thresh2;
    float inv_keep1, inv_keep2;
    uint64_t seed;
    // outputs / saved
    __nv_bfloat16* x0;             // [rows,H]
    float* x0_f32;                 // [rows,H] fp32 residual-stream copy (optional)
    float* mean1; float* rstd1;    // [rows] (text rows)
    float* mean2; float* rstd2;    // [rows] (rows of joint passes)
    __nv_bfloat16* pframe;         // [frame_rows,H] relu(W f + b), bf16
    int exact_frames;              // fp32 validation path: no bf16 rounding of the projection
    // backward
    const __nv_bfloat16* dx0;      // [rows,H]
    const float* dx0b;             // [rows,H] optional fp32 residual-stream gradient
    __nv_bfloat16* dpre;           // [frame_rows,H] gradient of the projection pre-activation
    float* g_word; float* g_pos; float* g_type;
    float* g_ln1_g; float* g_ln1_b; float* g_ln2_g; float* g_ln2_b;
    float* g_w[2];                 // [H,D]
    float* g_wb[2];                // [H]
    __nv_bfloat16* frames_bf16[2]; // optional [B*L, ldf] bf16 copy of the frames (ldf = D rounded up to 8)
    float* gw_pad[2];              // optional [H, ldf] scratch of the tensor-core projection wgrad
    int* err_count;                // optional: count of token ids / token types outside their tables (forward only)
    const int* row_live;           // optional per-row flags (mmb_embed_args.row_live)
};

constexpr uint32_t kStreamEmb1 = 0x100, kStreamEmb2 = 0x101;
constexpr int kEmbWarps = 8;

// index of a frame row in the compact [frame_rows, H] buffers
__device__ __forceinline__ int frame_index(const PackDims& d, int pass, int b, int l) {
    return pass == 1 ? b * d.L1 + l : d.B * d.L1 + b * d.L2 + l;
}
// packed row of frame (pass, b, l)
__device__ __forceinline__ int frame_packed_row(const PackDims& d, int pass, int b, int l) {
    return d.base(pass) + b * d.S(pass) + d.T + l;
}

template <int NCH>
__device__ __forceinline__ void text_embed_row(const EmbedParams& p, const RowCoord& c, int lane, RowF<NCH>& e, long long& id,
                                               int& tt) {
    id = p.ids[c.pass][(int64_t)c.b * p.d.T + c.s];
    long long tt64 = c.pass == 0 ? p.token_type[(int64_t)c.b * p.d.T + c.s] : 0;
    // nn.Embedding device-asserts on an index outside its table; here such an index reads row 0 (and gets no gradient)
    // and is counted, which turns the step's joint loss into NaN (mmb_heads_fwd)
    if (id < 0 || id >= p.V || tt64 < 0 || tt64 > 1) {
        if (lane == 0 && p.err_count != nullptr) atomicAdd(p.err_count, 1);
        if (id < 0 || id >= p.V) id = 0;
        if (tt64 < 0 || tt64 > 1) tt64 = 0;
    }
    tt = (int)tt64;
    RowF<NCH> a;
    row_load_f32(e, p.word + (int64_t)id * p.H, p.H, lane);
    row_load_f32(a, p.type + (int64_t)tt * p.H, p.H, lane);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
        for (int i = 0; i < 8; ++i) e.v[ch][i] += a.v[ch][i];
    row_load_f32(a, p.pos + (int64_t)c.s * p.H, p.H, lane);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
        for (int i = 0; i < 8; ++i) e.v[ch][i] += a.v[ch][i];
}

template <int NCH>
__global__ void __launch_bounds__(kEmbWarps * 32)
embed_text_fwd_kernel(const EmbedParams p) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntext = 3 * p.d.B * p.d.T;
    for (int tr = blockIdx.x * kEmbWarps + warp; tr < ntext; tr += gridDim.x * kEmbWarps) {
        // text row tr -> (pass, b, s)
        RowCoord c;
        c.pass = tr / (p.d.B * p.d.T);
        const int r = tr - c.pass * p.d.B * p.d.T;
        c.b = r / p.d.T;
        c.s = r - c.b * p.d.T;
        const int row = p.d.base(c.pass) + c.b * p.d.S(c.pass) + c.s;
        RowF<NCH> e;
        long long id;
        int tt;
        text_embed_row(p, c, lane, e, id, tt);
        float mean, rstd;
        row_stats(e, p.H, lane, p.eps1, mean, rstd);
        row_affine(e, p.H, lane, mean, rstd, p.ln1_g, p.ln1_b);
        row_dropout(e, p.H, lane, p.seed, kStreamEmb1, (uint64_t)row, p.thresh1, p.inv_keep1);
        if (lane == 0) {
            p.mean1[row] = mean;
            p.rstd1[row] = rstd;
        }
        if (c.pass != 0) {
            row_stats(e, p.H, lane, p.eps2, mean, rstd);
            row_affine(e, p.H, lane, mean, rstd, p.ln2_g, p.ln2_b);
            row_dropout(e, p.H, lane, p.seed, kStreamEmb2, (uint64_t)row, p.thresh2, p.inv_keep2);
            if (lane == 0) {
                p.mean2[row] = mean;
                p.rstd2[row] = rstd;
            }
        }
        row_store_bf16(e, p.x0 + (int64_t)row * p.H, p.H, lane);
        if (p.x0_f32 != nullptr) row_store_f32(e, p.x0_f32 + (int64_t)row * p.H, p.H, lane);
    }
}

// Frame rows: one CTA projects kFrameRows frames (FFMA, weights streamed through L1/L2 as W^T so that
// consecutive threads read consecutive output columns), then each warp layer-normalises two rows.
constexpr int kFrameRows = 16;

template <int NCH>
__global__ void __launch_bounds__(256)
embed_frame_fwd_kernel(const EmbedParams p, int mod) {  // mod 0 = visual (pass 1), 1 = speech (pass 2)
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    extern __shared__ float fsm[];
    const int D = p.frame_dim[mod], L = mod == 0 ? p.d.L1 : p.d.L2, pass = mod + 1;
    const int Dp = D + 1;
    float* sF = fsm;                       // [kFrameRows][Dp]
    float* sO = fsm + kFrameRows * Dp;     // [kFrameRows][H]
    const int nrows = p.d.B * L;
    const int r0 = blockIdx.x * kFrameRows;
    const int tid = threadIdx.x;
    if (p.row_live != nullptr) {          // a block of padding frames: nothing downstream reads its rows
        int any = 0;
        if (tid < kFrameRows && r0 + tid < nrows) {
            const int fr = r0 + tid, b = fr / L, l = fr - b * L;
            any = __ldg(p.row_live + frame_packed_row(p.d, pass, b, l));
        }
        if (!__syncthreads_or(any)) return;
    }
    for (int i = tid; i < kFrameRows * D; i += 256) {
        const int r = i / D, k = i - r * D;
        sF[r * Dp + k] = (r0 + r) < nrows ? load_as_float(p.frames[mod], p.frames_dt[mod], (int64_t)(r0 + r) * D + k) : 0.f;
    }
    __syncthreads();
    if (p.frames_bf16[mod] != nullptr) {      // bf16 copy (zero padded to ldf) for the tensor-core wgrad of the backward
        const int ldf = (D + 7) & ~7;
        for (int i = tid; i < kFrameRows * ldf; i += 256) {
            const int r = i / ldf, k = i - r * ldf;
            if (r0 + r < nrows)
                p.frames_bf16[mod][(int64_t)(r0 + r) * ldf + k] = __float2bfloat16_rn(k < D ? sF[r * Dp + k] : 0.f);
        }
    }
    float acc[kFrameRows][NCH];
#pragma unroll
    for (int r = 0; r < kFrameRows; ++r)
#pragma unroll
        for (int j = 0; j < NCH; ++j) acc[r][j] = 0.f;
    const float* __restrict__ wT = p.wT[mod];
    for (int k = 0; k < D; ++k) {
        float w[NCH];
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const int c = tid + 256 * j;
            w[j] = c < p.H ? __ldg(wT + (int64_t)k * p.H + c) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < kFrameRows; ++r) {
            const float f = sF[r * Dp + k];
#pragma unroll
            for (int j = 0; j < NCH; ++j) acc[r][j] = fmaf(f, w[j], acc[r][j]);
        }
    }
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int c = tid + 256 * j;
        if (c < p.H) {
            const float bias = __ldg(p.wb[mod] + c);
#pragma unroll
            for (int r = 0; r < kFrameRows; ++r) {
                const float v = fmaxf(acc[r][j] + bias, 0.f);
                sO[r * p.H + c] = p.exact_frames ? v : bf16_round(v);
            }
        }
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < kFrameRows; r += 8) {
        const int fr = r0 + r;
        if (fr >= nrows) break;
        const int b = fr / L, l = fr - b * L;
        const int row = frame_packed_row(p.d, pass, b, l);
        RowF<NCH> z;
        row_load_f32(z, sO + r * p.H, p.H, lane);
        row_store_bf16(z, p.pframe + (int64_t)frame_index(p.d, pass, b, l) * p.H, p.H, lane);
        float mean, rstd;
        row_stats(z, p.H, lane, p.eps2, mean, rstd);
        row_affine(z, p.H, lane, mean, rstd, p.ln2_g, p.ln2_b);
        row_dropout(z, p.H, lane, p.seed, kStreamEmb2, (uint64_t)row, p.thresh2, p.inv_keep2);
        if (lane == 0) {
            p.mean2[row] = mean;
            p.rstd2[row] = rstd;
        }
        row_store_bf16(z, p.x0 + (int64_t)row * p.H, p.H, lane);
        if (p.x0_f32 != nullptr) row_store_f32(z, p.x0_f32 + (int64_t)row * p.H, p.H, lane);
    }
}

// ------------------------------------------------------------------ embedding backward
template <int NCH>
__device__ __forceinline__ void row_dropout_bwd(RowF<NCH>& g, int H, int lane, uint64_t seed, uint32_t stream, uint64_t row,
                                                uint32_t thresh, float inv_keep) {
    row_dropout(g, H, lane, seed, stream, row, thresh, inv_keep);  // same mask, same scaling
}

template <int NCH>
__device__ __forceinline__ void row_atomic_add(const RowF<NCH>& g, float* __restrict__ dst, int H, int lane) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int e = (c * 32 + lane) * 8;
        if (e < H) {
            ptx::red_add_v4(dst + e, g.v[c][0], g.v[c][1], g.v[c][2], g.v[c][3]);
            ptx::red_add_v4(dst + e + 4, g.v[c][4], g.v[c][5], g.v[c][6], g.v[c][7]);
        }
    }
}

template <int NCH>
__global__ void __launch_bounds__(kEmbWarps * 32)
embed_text_bwd_kernel(const EmbedParams p) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntext = 3 * p.d.B * p.d.T;
    RowF<NCH> a_g1, a_b1, a_g2, a_b2, a_type0;
    row_zero(a_g1); row_zero(a_b1); row_zero(a_g2); row_zero(a_b2); row_zero(a_type0);
    for (int tr = blockIdx.x * kEmbWarps + warp; tr < ntext; tr += gridDim.x * kEmbWarps) {
        RowCoord c;
        c.pass = tr / (p.d.B * p.d.T);
        const int r = tr - c.pass * p.d.B * p.d.T;
        c.b = r / p.d.T;
        c.s = r - c.b * p.d.T;
        const int row = p.d.base(c.pass) + c.b * p.d.S(c.pass) + c.s;
        if (p.row_live != nullptr && __ldg(p.row_live + row) == 0) continue;      // padding row: its gradient is exactly zero
        RowF<NCH> e, g;
        long long id;
        int tt;
        text_embed_row(p, c, lane, e, id, tt);
        row_load_bf16(g, p.dx0 + (int64_t)row * p.H, p.H, lane);
        if (p.dx0b != nullptr) {
            RowF<NCH> g2;
            row_load_f32(g2, p.dx0b + (int64_t)row * p.H, p.H, lane);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
                for (int i = 0; i < 8; ++i) g.v[ch][i] += g2.v[ch][i];
        }
        const float mean1 = p.mean1[row], rstd1 = p.rstd1[row];
        if (c.pass != 0) {
            // recompute the joint LayerNorm's input: dropout1(LN1(e))
            RowF<NCH> z = e;
            row_affine(z, p.H, lane, mean1, rstd1, p.ln1_g, p.ln1_b);
            row_dropout(z, p.H, lane, p.seed, kStreamEmb1, (uint64_t)row, p.thresh1, p.inv_keep1);
            row_dropout_bwd(g, p.H, lane, p.seed, kStreamEmb2, (uint64_t)row, p.thresh2, p.inv_keep2);
            row_ln_bwd(z, g, p.H, lane, p.mean2[row], p.rstd2[row], p.ln2_g, a_g2, a_b2);
        }
        row_dropout_bwd(g, p.H, lane, p.seed, kStreamEmb1, (uint64_t)row, p.thresh1, p.inv_keep1);
        row_ln_bwd(e, g, p.H, lane, mean1, rstd1, p.ln1_g, a_g1, a_b1);
        // g = gradient of word[id] + type[tt] + pos[s]
        if (id != 0) row_atomic_add(g, p.g_word + (int64_t)id * p.H, p.H, lane);  // padding_idx = 0 gets no gradient
        row_atomic_add(g, p.g_pos + (int64_t)c.s * p.H, p.H, lane);
        if (tt == 0) {
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
                for (int i = 0; i < 8; ++i) a_type0.v[ch][i] += g.v[ch][i];
        } else {
            row_atomic_add(g, p.g_type + (int64_t)tt * p.H, p.H, lane);
        }
    }
    cta_flush_columns(a_g1, p.g_ln1_g, p.H, smem, warp, lane, kEmbWarps);
    cta_flush_columns(a_b1, p.g_ln1_b, p.H, smem, warp, lane, kEmbWarps);
    cta_flush_columns(a_g2, p.g_ln2_g, p.H, smem, warp, lane, kEmbWarps);
    cta_flush_columns(a_b2, p.g_ln2_b, p.H, smem, warp, lane, kEmbWarps);
    cta_flush_columns(a_type0, p.g_type, p.H, smem, warp, lane, kEmbWarps);
}

template <int NCH>
__global__ void __launch_bounds__(kEmbWarps * 32)
embed_frame_bwd_kernel(const EmbedParams p) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nfr = p.d.frame_rows();
    RowF<NCH> a_g2, a_b2;
    row_zero(a_g2); row_zero(a_b2);
    for (int fr = blockIdx.x * kEmbWarps + warp; fr < nfr; fr += gridDim.x * kEmbWarps) {
        int pass, b, l;
        if (fr < p.d.B * p.d.L1) { pass = 1; b = fr / p.d.L1; l = fr - b * p.d.L1; }
        else { const int r = fr - p.d.B * p.d.L1; pass = 2; b = r / p.d.L2; l = r - b * p.d.L2; }
        const int row = frame_packed_row(p.d, pass, b, l);
        RowF<NCH> z, g;
        if (p.row_live != nullptr && __ldg(p.row_live + row) == 0) {      // padding frame: dpre = 0, nothing read
            row_zero(g);
            row_store_bf16(g, p.dpre + (int64_t)fr * p.H, p.H, lane);
            continue;
        }
        row_load_bf16(z, p.pframe + (int64_t)fr * p.H, p.H, lane);
        row_load_bf16(g, p.dx0 + (int64_t)row * p.H, p.H, lane);
        if (p.dx0b != nullptr) {
            RowF<NCH> g2;
            row_load_f32(g2, p.dx0b + (int64_t)row * p.H, p.H, lane);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
                for (int i = 0; i < 8; ++i) g.v[ch][i] += g2.v[ch][i];
        }
        row_dropout_bwd(g, p.H, lane, p.seed, kStreamEmb2, (uint64_t)row, p.thresh2, p.inv_keep2);
        row_ln_bwd(z, g, p.H, lane, p.mean2[row], p.rstd2[row], p.ln2_g, a_g2, a_b2);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
            for (int i = 0; i < 8; ++i) g.v[ch][i] = z.v[ch][i] > 0.f ? g.v[ch][i] : 0.f;  // relu backward
        row_store_bf16(g, p.dpre + (int64_t)fr * p.H, p.H, lane);
    }
    cta_flush_columns(a_g2, p.g_ln2_g, p.H, smem, warp, lane, kEmbWarps);
    cta_flush_columns(a_b2, p.g_ln2_b, p.H, smem, warp, lane, kEmbWarps);
}

// dW[c][k] += sum_rows dpre[row][c] * frame[row][k];  db[c] += sum_rows dpre[row][c]
// grid (H / 64, row chunks); thread = (column c = tid % 64, k = tid / 64 + 4 i)
constexpr int kWgRows = 32;
template <int DI>
__global__ void __launch_bounds__(256)
frame_wgrad_kernel(const EmbedParams p, int mod, int rows_per_cta) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    extern __shared__ float wsm[];
    const int D = p.frame_dim[mod], L = mod == 0 ? p.d.L1 : p.d.L2;
    const int Dp = D + 1;
    float* sG = wsm;                   // [kWgRows][64]
    float* sF = wsm + kWgRows * 64;    // [kWgRows][Dp]
    const int nrows = p.d.B * L;
    const int fr_base = mod == 0 ? 0 : p.d.B * p.d.L1;
    const int c0 = blockIdx.x * 64;
    const int tid = threadIdx.x, c = tid & 63, kq = tid >> 6;
    const int rbeg = blockIdx.y * rows_per_cta, rend = min(nrows, rbeg + rows_per_cta);
    float acc[DI];
#pragma unroll
    for (int i = 0; i < DI; ++i) acc[i] = 0.f;
    float bacc = 0.f;
    for (int rc = rbeg; rc < rend; rc += kWgRows) {
        __syncthreads();
        for (int i = tid; i < kWgRows * 64; i += 256) {
            const int r = i >> 6, cc = i & 63;
            sG[i] = (rc + r) < rend && (c0 + cc) < p.H
                        ? __bfloat162float(p.dpre[(int64_t)(fr_base + rc + r) * p.H + c0 + cc]) : 0.f;
        }
        for (int i = tid; i < kWgRows * D; i += 256) {
            const int r = i / D, k = i - r * D;
            sF[r * Dp + k] = (rc + r) < rend ? load_as_float(p.frames[mod], p.frames_dt[mod], (int64_t)(rc + r) * D + k) : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < kWgRows; ++r) {
            const float gval = sG[r * 64 + c];
            if (kq == 0) bacc += gval;
#pragma unroll
            for (int i = 0; i < DI; ++i) {
                const int k = kq + 4 * i;
                if (k < D) acc[i] = fmaf(gval, sF[r * Dp + k], acc[i]);
            }
        }
    }
    if (c0 + c < p.H) {
#pragma unroll
        for (int i = 0; i < DI; ++i) {
            const int k = kq + 4 * i;
            if (k < D) atomicAdd(p.g_w[mod] + (int64_t)(c0 + c) * D + k, acc[i]);
        }
        if (kq == 0) atomicAdd(p.g_wb[mod] + c0 + c, bacc);
    }
}

// g_w[c][k] += pad[c][k]  (k < D; pad has row stride ldf)
__global__ void add_padded_kernel(float* __restrict__ gw, const float* __restrict__ pad, int H, int D, int ldf) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= H * D) return;
    const int c = i / D, k = i - c * D;
    gw[i] += pad[(size_t)c * ldf + k];
}

static int fill_embed(EmbedParams& p, const mmb_embed_args* a) {
    MMB_REQUIRE(a != nullptr, "embed: null args");
    MMB_REQUIRE(a->B > 0 && a->T > 0 && a->L[0] >= 0 && a->L[1] >= 0, "embed: bad dims");
    MMB_REQUIRE(a->H % 8 == 0 && a->H <= 1024 && a->H > 0, "embed: H=%d unsupported", a->H);
    MMB_REQUIRE(a->T <= a->max_pos, "embed: T=%d exceeds max_position_embeddings=%d", a->T, a->max_pos);
    p.d.B = a->B; p.d.T = a->T; p.d.L1 = a->L[0]; p.d.L2 = a->L[1];
    p.H = a->H; p.V = a->V; p.max_pos = a->max_pos;
    p.err_count = a->err_count;
    p.row_live = a->row_live;
    for (int i = 0; i < 3; ++i) p.ids[i] = (const long long*)a->ids[i];
    p.token_type = (const long long*)a->token_type;
    for (int i = 0; i < 2; ++i) {
        p.frame_dim[i] = a->frame_dim[i];
        p.frames[i] = a->frames[i];
        p.frames_dt[i] = a->frames_dtype[i];
        p.wT[i] = a->wT[i];
        p.wb[i] = a->wb[i];
        p.g_w[i] = a->g_w[i];
        p.g_wb[i] = a->g_wb[i];
    }
    p.word = a->word; p.pos = a->pos; p.type = a->type;
    p.ln1_g = a->ln1_g; p.ln1_b = a->ln1_b; p.eps1 = a->eps1;
    p.ln2_g = a->ln2_g; p.ln2_b = a->ln2_b; p.eps2 = a->eps2;
    p.thresh1 = dropout_threshold(a->p_drop1);
    p.thresh2 = dropout_threshold(a->p_drop2);
    p.inv_keep1 = dropout_inv_keep(a->p_drop1);
    p.inv_keep2 = dropout_inv_keep(a->p_drop2);
    p.seed = a->seed;
    p.x0 = (__nv_bfloat16*)a->x0;
    p.x0_f32 = a->x0_f32;
    p.mean1 = a->mean1; p.rstd1 = a->rstd1; p.mean2 = a->mean2; p.rstd2 = a->rstd2;
    p.pframe = (__nv_bfloat16*)a->pframe;
    p.exact_frames = a->exact_frames;
    for (int i = 0; i < 2; ++i) {
        p.frames_bf16[i] = (__nv_bfloat16*)a->frames_bf16[i];
        p.gw_pad[i] = a->gw_pad[i];
    }
    p.dx0 = (const __nv_bfloat16*)a->dx0;
    p.dx0b = a->dx0b;
    p.dpre = (__nv_bfloat16*)a->dpre;
    p.g_word = a->g_word; p.g_pos = a->g_pos; p.g_type = a->g_type;
    p.g_ln1_g = a->g_ln1_g; p.g_ln1_b = a->g_ln1_b; p.g_ln2_g = a->g_ln2_g; p.g_ln2_b = a->g_ln2_b;
    return MMB_OK;
}

}  // namespace mmb

using namespace mmb;

extern "C" int mmb_pack_prepare(const mmb_pack_args* a, void* stream) {
    MMB_REQUIRE(a && a->keybias && a->cu_seqlens && a->label_count, "pack_prepare: null pointer");
    MMB_REQUIRE(a->B > 0 && a->T > 0, "pack_prepare: bad dims");
    PrepParams p;
    p.d.B = a->B; p.d.T = a->T; p.d.L1 = a->L[0]; p.d.L2 = a->L[1];
    for (int i = 0; i < 3; ++i) {
        MMB_REQUIRE(a->mask_text[i] != nullptr, "pack_prepare: null text mask %d", i);
        p.mask_text[i] = a->mask_text[i];
        p.mask_text_dt[i] = a->mask_text_dtype[i];
        p.labels[i] = (const long long*)a->labels[i];
    }
    for (int i = 0; i < 2; ++i) {
        MMB_REQUIRE(a->L[i] == 0 || a->mask_frame[i] != nullptr, "pack_prepare: null frame mask %d", i);
        p.mask_frame[i] = a->mask_frame[i];
        p.mask_frame_dt[i] = a->mask_frame_dtype[i];
        MMB_REQUIRE(a->mask_frame_stride[i] >= 0, "pack_prepare: negative frame mask stride %d", i);
        p.mask_frame_stride[i] = a->mask_frame_stride[i] > 0 ? a->mask_frame_stride[i] : a->frame_dim[i];
    }
    p.keybias = a->keybias;
    p.cu_seqlens = a->cu_seqlens;
    p.label_count = a->label_count;
    p.kv_end = a->kv_end;
    p.row_label = a->row_label;
    p.vocab = a->vocab;
    MMB_CUDA(cudaMemsetAsync(a->label_count, 0, 4 * sizeof(int), (cudaStream_t)stream));
    if (a->kv_end) MMB_CUDA(cudaMemsetAsync(a->kv_end, 0, 3 * (size_t)a->B * sizeof(int), (cudaStream_t)stream));
    const int rows = p.d.rows();
    const int grid = min((rows + 255) / 256, num_sms() * 2);
    // first kernel of a step: a plain launch (its inputs arrive through event waits from a copy stream; see common.cuh)
    pack_prepare_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("pack_prepare_kernel");
}

extern "C" int mmb_embed_fwd(const mmb_embed_args* a, void* stream) {
    EmbedParams p;
    int rc = fill_embed(p, a);
    if (rc != MMB_OK) return rc;
    MMB_REQUIRE(p.ids[0] && p.ids[1] && p.ids[2] && p.token_type && p.word && p.pos && p.type && p.x0 && p.mean1 &&
                    p.rstd1 && p.mean2 && p.rstd2 && p.ln1_g && p.ln1_b && p.ln2_g && p.ln2_b,
                "embed_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int ntext = 3 * p.d.B * p.d.T;
    const int grid = min((ntext + kEmbWarps - 1) / kEmbWarps, num_sms() * 4);
    MMB_DISPATCH_NCH(p.H, (launch_pdl(embed_text_fwd_kernel<NCH>, dim3(grid), dim3(kEmbWarps * 32), (size_t)(0), st, p)));
    rc = check_launch("embed_text_fwd_kernel");
    if (rc != MMB_OK) return rc;
    for (int mod = 0; mod < 2; ++mod) {
        const int L = mod == 0 ? p.d.L1 : p.d.L2;
        if (L == 0) continue;
        MMB_REQUIRE(p.frames[mod] && p.wT[mod] && p.wb[mod] && p.pframe, "embed_fwd: null frame operand %d", mod);
        const int nrows = p.d.B * L;
        const size_t smem = (size_t)kFrameRows * (p.frame_dim[mod] + 1 + p.H) * sizeof(float);
        MMB_DISPATCH_NCH(p.H, {
            MMB_ENSURE_SMEM(160 * 1024, embed_frame_fwd_kernel<NCH>);
            launch_pdl(embed_frame_fwd_kernel<NCH>, dim3((nrows + kFrameRows - 1) / kFrameRows), dim3(256), (size_t)(smem), st, p, mod);
        });
        rc = check_launch("embed_frame_fwd_kernel");
        if (rc != MMB_OK) return rc;
    }
    return MMB_OK;
}

extern "C" int mmb_embed_bwd(const mmb_embed_args* a, void* stream) {
    EmbedParams p;
    int rc = fill_embed(p, a);
    if (rc != MMB_OK) return rc;
    MMB_REQUIRE(p.dx0 && p.g_word && p.g_pos && p.g_type && p.g_ln1_g && p.g_ln1_b && p.g_ln2_g && p.g_ln2_b,
                "embed_bwd: null pointer");
    p.err_count = nullptr;      // counted once, by the forward
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)kEmbWarps * p.H * sizeof(float);
    const int ntext = 3 * p.d.B * p.d.T;
    int grid = min((ntext + kEmbWarps - 1) / kEmbWarps, num_sms());
    MMB_DISPATCH_NCH(p.H, (launch_pdl(embed_text_bwd_kernel<NCH>, dim3(grid), dim3(kEmbWarps * 32), (size_t)(smem), st, p)));
    rc = check_launch("embed_text_bwd_kernel");
    if (rc != MMB_OK) return rc;
    const int nfr = p.d.frame_rows();
    if (nfr > 0) {
        MMB_REQUIRE(p.pframe && p.dpre, "embed_bwd: null frame buffers");
        grid = min((nfr + kEmbWarps - 1) / kEmbWarps, num_sms());
        MMB_DISPATCH_NCH(p.H, (launch_pdl(embed_frame_bwd_kernel<NCH>, dim3(grid), dim3(kEmbWarps * 32), (size_t)(smem), st, p)));
        rc = check_launch("embed_frame_bwd_kernel");
        if (rc != MMB_OK) return rc;
        for (int mod = 0; mod < 2; ++mod) {
            const int L = mod == 0 ? p.d.L1 : p.d.L2;
            if (L == 0) continue;
            MMB_REQUIRE(p.g_w[mod] && p.g_wb[mod] && p.frames[mod], "embed_bwd: null projection grads %d", mod);
            const int D = p.frame_dim[mod], nrows = p.d.B * L;
            MMB_REQUIRE(D <= 384, "embed_bwd: frame dim %d > 384 unsupported", D);
            if (p.frames_bf16[mod] != nullptr && p.gw_pad[mod] != nullptr) {
                // tensor-core path: g_w += dpre^T (H x rows) · frames (rows x D) as a split-K GEMM into the padded scratch,
                // g_wb += colsum(dpre)
                const int ldf = (D + 7) & ~7;
                const int fr_base = mod == 0 ? 0 : p.d.B * p.d.L1;
                MMB_CUDA(cudaMemsetAsync(p.gw_pad[mod], 0, (size_t)p.H * ldf * sizeof(float), st));
                mmb_gemm_args ga;
                memset(&ga, 0, sizeof(ga));
                ga.A = p.dpre + (int64_t)fr_base * p.H;
                ga.B = p.frames_bf16[mod];
                ga.C = p.gw_pad[mod];
                ga.lda = p.H;
                ga.ldb = ldf;
                ga.ldc = ldf;
                ga.M = p.H;
                ga.N = D;
                ga.K = nrows;
                ga.a_major = MMB_MAJOR_MN;
                ga.b_major = MMB_MAJOR_MN;
                ga.epilogue = MMB_EPI_ATOMIC_ADD_F32;
                ga.alpha = 1.0f;
                const int tile = D <= 128 ? 128 : 256, slots = D <= 128 ? num_sms() : num_sms() / 2;
                const int tiles = ((p.H + tile - 1) / tile) * ((D + tile - 1) / tile);
                int sk = slots / tiles, kb = (nrows + 63) / 64;
                if (sk > kb / 4) sk = kb / 4;
                ga.split_k = sk < 1 ? 1 : sk;
                rc = mmb_gemm(&ga, st);
                if (rc != MMB_OK) return rc;
                launch_pdl(add_padded_kernel, dim3((p.H * D + 255) / 256), dim3(256), (size_t)(0), st, p.g_w[mod], p.gw_pad[mod], p.H, D, ldf);
                rc = check_launch("add_padded_kernel");
                if (rc != MMB_OK) return rc;
                mmb_colsum_args ca;
                ca.X = p.dpre + (int64_t)fr_base * p.H;
                ca.out = p.g_wb[mod];
                ca.ld = p.H;
                ca.M = nrows;
                ca.N = p.H;
                ca.row_list = nullptr;
                rc = mmb_colsum_bf16(&ca, st);
                if (rc != MMB_OK) return rc;
                continue;
            }
            const int col_tiles = (p.H + 63) / 64;
            int chunks = (num_sms() * 2 + col_tiles - 1) / col_tiles;
            int rows_per_cta = (nrows + chunks - 1) / chunks;
            rows_per_cta = (rows_per_cta + kWgRows - 1) / kWgRows * kWgRows;
            chunks = (nrows + rows_per_cta - 1) / rows_per_cta;
            const size_t wsmem = (size_t)kWgRows * (64 + D + 1) * sizeof(float);
            dim3 g(col_tiles, chunks);
            if (D <= 96) {
                launch_pdl(frame_wgrad_kernel<24>, dim3(g), dim3(256), (size_t)(wsmem), st, p, mod, rows_per_cta);
            } else {
                MMB_ENSURE_SMEM(96 * 1024, frame_wgrad_kernel<96>);
                launch_pdl(frame_wgrad_kernel<96>, dim3(g), dim3(256), (size_t)(wsmem), st, p, mod, rows_per_cta);
            }
            rc = check_launch("frame_wgrad_kernel");
            if (rc != MMB_OK) return rc;
        }
    }
    return MMB_OK;
}

// ------------------------------------------------------------------ on-device MLM masking (model_utils.py:6-39)
namespace mmb {
struct MlmParams {
    long long* ids;
    long long* labels;
    long long* labels_dup;
    int special[8];
    int n_special, B, T, mask_id;
    uint32_t thr_sel, thr_rep;   // 16-bit thresholds: select iff u < thr_sel, replace iff u' < thr_rep
    uint64_t seed;
    uint32_t stream;
};
__global__ void __launch_bounds__(256)
mlm_mask_kernel(const MlmParams p) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= p.B * p.T) return;
    const long long id = p.ids[i];
    bool special = false;
#pragma unroll
    for (int s = 0; s < 8; ++s) special |= (s < p.n_special) && (id == (long long)p.special[s]);
    // one hash per element: two independent 16-bit uniforms (selection, replacement)
    const uint32_t bits = rng_pair(rng_row_key(p.seed, p.stream, (uint32_t)(i >> 16)), (uint32_t)(i & 0xFFFF));
    const bool selected = !special && (bits & 0xFFFFu) < p.thr_sel;
    const bool replaced = selected && (bits >> 16) < p.thr_rep;
    const long long lab = selected ? id : -100;
    p.labels[i] = lab;
    if (p.labels_dup != nullptr) {
        const int b = i / p.T, t = i - b * p.T;
        p.labels_dup[(size_t)b * 2 * p.T + t] = lab;
        p.labels_dup[(size_t)b * 2 * p.T + p.T + t] = lab;
    }
    if (replaced) p.ids[i] = p.mask_id;
}
}  // namespace mmb

extern "C" int mmb_mlm_mask(const mmb_mlm_mask_args* a, void* stream) {
    MMB_REQUIRE(a && a->ids && a->labels, "mlm_mask: null pointer");
    MMB_REQUIRE(a->B > 0 && a->T > 0 && a->n_special >= 0 && a->n_special <= 8, "mlm_mask: bad shape");
    MMB_REQUIRE(a->prob >= 0.f && a->prob <= 1.f && a->replace_prob >= 0.f && a->replace_prob <= 1.f, "mlm_mask: bad probability");
    MlmParams p;
    p.ids = (long long*)a->ids;
    p.labels = (long long*)a->labels;
    p.labels_dup = (long long*)a->labels_dup;
    for (int i = 0; i < 8; ++i) p.special[i] = a->special[i];
    p.n_special = a->n_special;
    p.B = a->B;
    p.T = a->T;
    p.mask_id = a->mask_id;
    p.thr_sel = (uint32_t)((double)a->prob * 65536.0 + 0.5);
    p.thr_rep = (uint32_t)((double)a->replace_prob * 65536.0 + 0.5);
    p.seed = a->seed;
    p.stream = a->rng_stream;
    const int n = a->B * a->T;
    mlm_mask_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("mlm_mask_kernel");
}
