// Masked-LM cross entropy over the vocabulary logits (forward statistics + backward dlogits).
//
// Replaces torch.nn.CrossEntropyLoss()(prediction_scores.view(-1, V), labels.view(-1)) of
// MMBertForPretraining.get_outputs (MMBertForPretraining.py:381-384): ignore_index = -100, mean over the
// labelled rows of ONE pass; the three passes are averaged at :427.  Logits are the bf16 output of the tied
// decoder GEMM, row stride ldl (V rounded up to a multiple of 8).
#include "common.cuh"

namespace mmb {

struct CeParams {
    const __nv_bfloat16* logits;  // [rows, ldl]
    __nv_bfloat16* dlogits;       // [rows, ldl] (backward)
    const long long* labels[3];   // per pass, flattened [B * S(pass)]
    int pass_base[4];             // packed row offsets of the passes; [3] = rows
    const int* label_count;       // [3]
    float* row_lse;               // [rows]   natural-log logsumexp of labelled rows
    float* loss_sum;              // [3]      sum over labelled rows of (lse - logit[label])
    const float* gscale;          // device scalar: upstream gradient of the joint loss (NULL = 1)
    float coef;                   // alpha / 3
    int V;
    int64_t ldl;
    int dense;                    // backward: also write the zero rows of unlabelled positions
};

__device__ __forceinline__ long long row_label(const CeParams& p, int row, int& pass) {
    pass = row < p.pass_base[1] ? 0 : (row < p.pass_base[2] ? 1 : 2);
    return p.labels[pass] ? p.labels[pass][row - p.pass_base[pass]] : -100;
}

__device__ __forceinline__ void online_merge(float& m, float& s, float m2, float s2) {
    const float mn = fmaxf(m, m2);
    s = (m == -INFINITY ? 0.f : s * __expf(m - mn)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mn));
    m = mn;
}

__global__ void __launch_bounds__(256)
ce_fwd_kernel(const CeParams p) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int row = blockIdx.x;
    int pass;
    const long long label = row_label(p, row, pass);
    if (label < 0 || label >= p.V) return;     // -100 = ignore; anything else out of range is counted by mmb_pack_prepare
    const __nv_bfloat16* x = p.logits + (int64_t)row * p.ldl;
    float m = -INFINITY, s = 0.f;
    const int nvec = (p.V + 7) / 8;
    for (int i = threadIdx.x; i < nvec; i += 256) {
        const uint4 q = *reinterpret_cast<const uint4*>(x + i * 8);  // ldl >= 8 * nvec: in-bounds of the row
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        float v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(w[j]);
            v[2 * j] = f.x;
            v[2 * j + 1] = f.y;
        }
        float mc = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (i * 8 + j >= p.V) v[j] = -INFINITY;
            mc = fmaxf(mc, v[j]);
        }
        float sc = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) sc += __expf(v[j] - mc);
        online_merge(m, s, mc, sc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        online_merge(m, s, m2, s2);
    }
    __shared__ float sm[8], ss[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        sm[warp] = m;
        ss[warp] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) online_merge(m, s, sm[w], ss[w]);
        const float lse = m + logf(s);
        p.row_lse[row] = lse;
        atomicAdd(p.loss_sum + pass, lse - __bfloat162float(x[label]));
    }
}

__global__ void __launch_bounds__(256)
ce_bwd_kernel(const CeParams p) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int row = blockIdx.x;
    int pass;
    const long long label = row_label(p, row, pass);
    __nv_bfloat16* dx = p.dlogits + (int64_t)row * p.ldl;
    const int nvec = (int)(p.ldl / 8);
    if (label < 0 || label >= p.V) {
        if (p.dense)
            for (int i = threadIdx.x; i < nvec; i += 256) reinterpret_cast<uint4*>(dx)[i] = make_uint4(0, 0, 0, 0);
        return;
    }
    const __nv_bfloat16* x = p.logits + (int64_t)row * p.ldl;
    const float lse = p.row_lse[row];
    const float scale = p.coef * (p.gscale ? *p.gscale : 1.f) / (float)p.label_count[pass];
    for (int i = threadIdx.x; i < nvec; i += 256) {
        const uint4 q = *reinterpret_cast<const uint4*>(x + i * 8);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        float v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(w[j]);
            v[2 * j] = f.x;
            v[2 * j + 1] = f.y;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = i * 8 + j;
            float g = col < p.V ? __expf(v[j] - lse) : 0.f;
            if (col == (int)label) g -= 1.f;
            v[j] = g * scale;
        }
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]);
        o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4], v[5]);
        o.w = pack_bf16x2(v[6], v[7]);
        reinterpret_cast<uint4*>(dx)[i] = o;
    }
}

// fp32 validation path: the same statistics on f32 logits (one CTA per labelled row)
__global__ void __launch_bounds__(256)
ce_fwd_f32_kernel(const CeParams p) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int row = blockIdx.x;
    int pass;
    const long long label = row_label(p, row, pass);
    if (label < 0 || label >= p.V) return;
    const float* x = reinterpret_cast<const float*>(p.logits) + (int64_t)row * p.ldl;
    float m = -INFINITY, s = 0.f;
    for (int i = threadIdx.x; i < p.V; i += 256) {
        const float v = x[i];
        const float mn = fmaxf(m, v);
        s = (m == -INFINITY ? 0.f : s * expf(m - mn)) + expf(v - mn);
        m = mn;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        const float mn = fmaxf(m, m2);
        s = (m == -INFINITY ? 0.f : s * expf(m - mn)) + (m2 == -INFINITY ? 0.f : s2 * expf(m2 - mn));
        m = mn;
    }
    __shared__ float sm[8], ss[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        sm[warp] = m;
        ss[warp] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            const float mn = fmaxf(m, sm[w]);
            s = (m == -INFINITY ? 0.f : s * expf(m - mn)) + (sm[w] == -INFINITY ? 0.f : ss[w] * expf(sm[w] - mn));
            m = mn;
        }
        const float lse = m + logf(s);
        p.row_lse[row] = lse;
        atomicAdd(p.loss_sum + pass, lse - x[label]);
    }
}

// ------------------------------------------------------------------ fused-CE path (no materialised logits)
// The decoder GEMM (MMB_EPI_CE_STATS, gemm_tcgen05.cu: ce_stats_warp) left, for every labelled row, one (max, sum exp)
// record per 128-column group, the label's logit and the row's bf16 logits in the dlogits buffer.
struct CeSparseParams {
    const int* row_label;   // [rows]
    float* stats;           // planes of `rows` floats: [2g] max, [2g+1] sum, [2G] label logit, [2G+1] row-written flag
    __nv_bfloat16* dlogits; // [rows, ldl]
    const int* label_count; // [3]
    float* row_lse;
    float* loss_sum;        // [3]
    const float* gscale;
    float* dbias;           // [V]
    int pass_base[4];
    float coef;
    int V, G, rows;
    int64_t ldl;
};

// one warp per row: merges the G group records of a labelled row
__global__ void __launch_bounds__(256)
ce_sparse_fwd_kernel(const CeSparseParams p) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= p.rows) return;
    if (p.row_label[row] == -100) return;
    float m = -INFINITY, s = 0.f;
    for (int g = lane; g < p.G; g += 32)
        online_merge(m, s, p.stats[(size_t)(2 * g) * p.rows + row], p.stats[(size_t)(2 * g + 1) * p.rows + row]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        online_merge(m, s, m2, s2);
    }
    if (lane == 0) {
        const float lse = m + logf(s);
        p.row_lse[row] = lse;
        const int pass = row < p.pass_base[1] ? 0 : (row < p.pass_base[2] ? 1 : 2);
        atomicAdd(p.loss_sum + pass, lse - p.stats[(size_t)(2 * p.G) * p.rows + row]);
    }
}

// one CTA per row.  Labelled: logits (bf16, written by the GEMM epilogue) -> coef * gscale / count * (softmax - onehot),
// in place.  Unlabelled but written in an earlier step: back to zero.  Everything else is already zero.
__global__ void __launch_bounds__(256)
ce_sparse_bwd_kernel(const CeSparseParams p) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int row = blockIdx.x;
    const int label = p.row_label[row];
    float* flag = p.stats + (size_t)(2 * p.G + 1) * p.rows + row;
    __nv_bfloat16* dx = p.dlogits + (int64_t)row * p.ldl;
    const int nvec = (int)(p.ldl / 8);
    if (label == -100) {
        if (*flag != 0.f) {
            for (int i = threadIdx.x; i < nvec; i += 256) reinterpret_cast<uint4*>(dx)[i] = make_uint4(0, 0, 0, 0);
            __syncthreads();
            if (threadIdx.x == 0) *flag = 0.f;
        }
        return;
    }
    const int pass = row < p.pass_base[1] ? 0 : (row < p.pass_base[2] ? 1 : 2);
    const float lse = p.row_lse[row];
    const float scale = p.coef * (p.gscale ? *p.gscale : 1.f) / (float)p.label_count[pass];
    for (int i = threadIdx.x; i < nvec; i += 256) {
        const uint4 q = reinterpret_cast<const uint4*>(dx)[i];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        float v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(w[j]);
            v[2 * j] = f.x;
            v[2 * j + 1] = f.y;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = i * 8 + j;
            float g = col < p.V ? __expf(v[j] - lse) : 0.f;
            if (col == label) g -= 1.f;
            v[j] = g * scale;
        }
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]);
        o.y = pack_bf16x2(v[2], v[3]);
        o.z = pack_bf16x2(v[4], v[5]);
        o.w = pack_bf16x2(v[6], v[7]);
        reinterpret_cast<uint4*>(dx)[i] = o;
    }
}

// decoder-bias gradient: column sums of dlogits over the LABELLED rows (all other rows are zero).
// grid (column strips of 512, row slices); each thread owns a column pair.
constexpr int kRowSlice = 1024;
__global__ void __launch_bounds__(256)
colsum_rows_kernel(const CeSparseParams p) {
    pdl_trigger();     // programmatic dependent launch (common.cuh): no global access before the wait
    pdl_wait();
    const int col = (blockIdx.x * 256 + threadIdx.x) * 2;
    const int r0 = blockIdx.y * kRowSlice, r1 = min(p.rows, r0 + kRowSlice);
    __shared__ int s_rows[kRowSlice];
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int r = r0 + threadIdx.x; r < r1; r += 256)
        if (p.row_label[r] != -100) s_rows[atomicAdd(&s_n, 1)] = r;
    __syncthreads();
    const int n = s_n;
    if (n == 0 || col >= p.V) return;
    float a0 = 0.f, a1 = 0.f;
    for (int k = 0; k < n; ++k) {
        const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p.dlogits + (int64_t)s_rows[k] * p.ldl + col));
        a0 += f.x;
        a1 += f.y;
    }
    atomicAdd(p.dbias + col, a0);
    if (col + 1 < p.V) atomicAdd(p.dbias + col + 1, a1);
}

static int fill_sparse(CeSparseParams& p, const mmb_ce_sparse_args* a) {
    MMB_REQUIRE(a && a->row_label && a->stats && a->label_count && a->row_lse, "ce_sparse: null pointer");
    MMB_REQUIRE(a->V > 128 && a->ldl >= (a->V + 7) / 8 * 8 && a->ldl % 8 == 0, "ce_sparse: V=%d ldl=%lld", a->V, (long long)a->ldl);
    p.row_label = a->row_label;
    p.stats = a->stats;
    p.dlogits = (__nv_bfloat16*)a->dlogits;
    p.label_count = a->label_count;
    p.row_lse = a->row_lse;
    p.loss_sum = a->loss_sum;
    p.gscale = a->gscale;
    p.dbias = a->dbias;
    const int B = a->B, T = a->T;
    p.pass_base[0] = 0;
    p.pass_base[1] = B * T;
    p.pass_base[2] = p.pass_base[1] + B * (T + a->L[0]);
    p.pass_base[3] = p.pass_base[2] + B * (T + a->L[1]);
    p.rows = p.pass_base[3];
    p.coef = a->coef;
    p.V = a->V;
    p.G = (a->V + 127) / 128;
    p.ldl = a->ldl;
    return MMB_OK;
}

static int fill(CeParams& p, const mmb_ce_args* a) {
    MMB_REQUIRE(a && a->logits && a->label_count && a->row_lse, "ce: null pointer");
    MMB_REQUIRE(a->V > 0 && a->ldl >= (a->logits_f32 ? a->V : (a->V + 7) / 8 * 8) && (a->logits_f32 || a->ldl % 8 == 0),
                "ce: ldl=%lld must be a multiple of 8 >= V=%d", (long long)a->ldl, a->V);
    p.logits = (const __nv_bfloat16*)a->logits;
    p.dlogits = (__nv_bfloat16*)a->dlogits;
    for (int i = 0; i < 3; ++i) p.labels[i] = (const long long*)a->labels[i];
    const int B = a->B, T = a->T;
    p.pass_base[0] = 0;
    p.pass_base[1] = B * T;
    p.pass_base[2] = p.pass_base[1] + B * (T + a->L[0]);
    p.pass_base[3] = p.pass_base[2] + B * (T + a->L[1]);
    p.label_count = a->label_count;
    p.row_lse = a->row_lse;
    p.loss_sum = a->loss_sum;
    p.gscale = a->gscale;
    p.coef = a->coef;
    p.V = a->V;
    p.ldl = a->ldl;
    p.dense = a->dense;
    return MMB_OK;
}

}  // namespace mmb

using namespace mmb;

extern "C" int mmb_ce_fwd(const mmb_ce_args* a, void* stream) {
    CeParams p;
    int rc = fill(p, a);
    if (rc != MMB_OK) return rc;
    MMB_REQUIRE(a->loss_sum != nullptr, "ce_fwd: null loss_sum");
    MMB_CUDA(cudaMemsetAsync(a->loss_sum, 0, 3 * sizeof(float), (cudaStream_t)stream));
    if (a->logits_f32) {
        launch_pdl(ce_fwd_f32_kernel, dim3(p.pass_base[3]), dim3(256), (size_t)(0), (cudaStream_t)stream, p);
        return check_launch("ce_fwd_f32_kernel");
    }
    launch_pdl(ce_fwd_kernel, dim3(p.pass_base[3]), dim3(256), (size_t)(0), (cudaStream_t)stream, p);
    return check_launch("ce_fwd_kernel");
}

extern "C" int mmb_ce_bwd(const mmb_ce_args* a, void* stream) {
    CeParams p;
    int rc = fill(p, a);
    if (rc != MMB_OK) return rc;
    MMB_REQUIRE(a->dlogits != nullptr, "ce_bwd: null dlogits");
    MMB_REQUIRE(!a->logits_f32, "ce_bwd: the fp32 validation path is forward-only");
    launch_pdl(ce_bwd_kernel, dim3(p.pass_base[3]), dim3(256), (size_t)(0), (cudaStream_t)stream, p);
    return check_launch("ce_bwd_kernel");
}

extern "C" int mmb_ce_sparse_fwd(const mmb_ce_sparse_args* a, void* stream) {
    CeSparseParams p;
    int rc = fill_sparse(p, a);
    if (rc != MMB_OK) return rc;
    MMB_REQUIRE(a->loss_sum != nullptr, "ce_sparse_fwd: null loss_sum");
    MMB_CUDA(cudaMemsetAsync(a->loss_sum, 0, 3 * sizeof(float), (cudaStream_t)stream));
    launch_pdl(ce_sparse_fwd_kernel, dim3((p.rows + 7) / 8), dim3(256), (size_t)(0), (cudaStream_t)stream, p);
    return check_launch("ce_sparse_fwd_kernel");
}

extern "C" int mmb_ce_sparse_bwd(const mmb_ce_sparse_args* a, void* stream) {
    CeSparseParams p;
    int rc = fill_sparse(p, a);
    if (rc != MMB_OK) return rc;
    MMB_REQUIRE(a->dlogits != nullptr, "ce_sparse_bwd: null dlogits");
    launch_pdl(ce_sparse_bwd_kernel, dim3(p.rows), dim3(256), (size_t)(0), (cudaStream_t)stream, p);
    return check_launch("ce_sparse_bwd_kernel");
}

extern "C" int mmb_colsum_rows_bf16(const mmb_ce_sparse_args* a, void* stream) {
    CeSparseParams p;
    int rc = fill_sparse(p, a);
    if (rc != MMB_OK) return rc;
    MMB_REQUIRE(a->dlogits != nullptr && a->dbias != nullptr, "colsum_rows: null pointer");
    dim3 grid((p.V + 511) / 512, (p.rows + kRowSlice - 1) / kRowSlice);
    launch_pdl(colsum_rows_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)stream, p);
    return check_launch("colsum_rows_kernel");
}
