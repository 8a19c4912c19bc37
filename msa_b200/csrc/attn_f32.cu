// fp32 validation path (forward only, see include/mmbert_sm100.h): masked self-attention with fp32 storage and fp32
// CUDA-core arithmetic, one warp per (query row, head).  Same arithmetic as eager_attention_forward
// (modeling_bert.py:115-140): scores = Q K^T / sqrt(64) + additive mask, softmax over the keys, context = P V; no
// dropout (the fp32 parity runs are dropout-free).  Correctness-first: not on the benchmarked path.
#include "common.cuh"

namespace mmb {

constexpr int kF32Warps = 4;

__global__ void __launch_bounds__(kF32Warps * 32)
attn_f32_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ ctx, const float* __restrict__ keybias,
                    const int* __restrict__ cu_seqlens, int H, int nheads, int nseq, int max_seqlen, int total_rows) {
    extern __shared__ float sc[];                       // [kF32Warps][max_seqlen] scores / probabilities
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * kF32Warps + warp;     // (row, head)
    if (item >= total_rows * nheads) return;
    const int row = item / nheads, head = item - row * nheads;
    // the sequence that owns this packed row
    int lo = 0, hi = nseq;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (cu_seqlens[mid] <= row) lo = mid; else hi = mid;
    }
    const int row0 = cu_seqlens[lo], S = cu_seqlens[lo + 1] - row0;
    float* s = sc + (size_t)warp * max_seqlen;
    const int64_t ld = 3 * (int64_t)H;
    const float* q = qkv + (int64_t)row * ld + head * 64;
    float qr[64];
#pragma unroll
    for (int d = 0; d < 64; ++d) qr[d] = q[d];
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) {
        const float* k = qkv + (int64_t)(row0 + j) * ld + H + head * 64;
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < 64; ++d) acc = fmaf(qr[d], k[d], acc);
        acc = acc * 0.125f + keybias[row0 + j];
        s[j] = acc;
        mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) {
        const float e = expf(s[j] - mx);
        s[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    float o0 = 0.f, o1 = 0.f;                            // this lane's two context columns
    for (int j = 0; j < S; ++j) {
        const float p = s[j];
        const float* v = qkv + (int64_t)(row0 + j) * ld + 2 * H + head * 64;
        o0 = fmaf(p, v[lane], o0);
        o1 = fmaf(p, v[lane + 32], o1);
    }
    float* out = ctx + (int64_t)row * H + head * 64;
    out[lane] = o0 * inv;
    out[lane + 32] = o1 * inv;
}

}  // namespace mmb

using namespace mmb;

extern "C" int mmb_attn_f32_fwd(const mmb_attn_f32_args* a, void* stream) {
    MMB_REQUIRE(a && a->qkv && a->ctx && a->keybias && a->cu_seqlens, "attn_f32: null pointer");
    MMB_REQUIRE(a->H > 0 && a->nheads > 0 && a->H == a->nheads * 64, "attn_f32: head dim must be 64");
    MMB_REQUIRE(a->nseq > 0 && a->max_seqlen > 0 && a->total_rows > 0, "attn_f32: empty batch");
    const size_t smem = (size_t)kF32Warps * a->max_seqlen * sizeof(float);
    MMB_REQUIRE(smem <= 200 * 1024, "attn_f32: max_seqlen %d too long for the validation kernel", a->max_seqlen);
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        MMB_CUDA(cudaFuncSetAttribute(attn_f32_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    const long long items = (long long)a->total_rows * a->nheads;
    attn_f32_fwd_kernel<<<(unsigned)((items + kF32Warps - 1) / kF32Warps), kF32Warps * 32, smem, (cudaStream_t)stream>>>(
        a->qkv, a->ctx, a->keybias, a->cu_seqlens, a->H, a->nheads, a->nseq, a->max_seqlen, a->total_rows);
    return check_launch("attn_f32_fwd_kernel");
}
