// Masked self-attention over packed variable-length sequences (text | frames), head dim 64: C-ABI entry points.
//
// Replaces BertSelfAttention's score/softmax/dropout/context chain:
//   modeling_bert.py:115-140 (eager_attention_forward) / sdpa_attention.py:40-102, called from :168-207,
//   with the additive mask of MMBertForPretraining.py:57-154,246-250 ((1-m) * -10000 per key, text ⊕ frames),
// and its autograd backward.  Q, K, V are read in place from the fused QKV projection output [rows, 3H]; the context is
// written [rows, H].  The kernels are in attn_tc.cu (forward) and attn_bwd_tc.cu (backward), both on
// tcgen05 / TMEM / TMA; this file holds the backward's preparation pass and the argument checks.
//
// Backward preparation (one HBM-bound pass over dO and O, 8 lanes per (row, head)): everything the backward kernels
// need per (head, packed row) is laid out as 16-byte records that their TMA producer copies next to the operand
// tiles — so the producer warp issues copies only and no kernel thread hashes or gathers per-row data in its loop:
//     Rq[h][row] = { -LSE (log2 domain, from the forward), -D with D = rowsum(dO ∘ O), dropout key of the probability
//                    ROW (query `row` of head h), 0 }
//     Rk[h][row] = { key bias * log2(e), dropout key of the probability COLUMN (key `row` of head h), 0, 0 }
// ws = [Rq planes: nheads][Rk planes: nheads], each plane total_rows records.  (16-byte records because TMA needs
// 16-byte aligned source addresses and the packed row offsets of the sequences are arbitrary.)
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace mmb {

#ifndef MMB_ATTN_FWD_DEFAULT_WS
#define MMB_ATTN_FWD_DEFAULT_WS 1
#endif
constexpr int kD = 64;       // head dim
constexpr float kLog2e = 1.4426950408889634f;

int launch_attn_fwd_tc(const mmb_attn_args* a, cudaStream_t stream);                        // attn_tc.cu
int launch_attn_fwd_ws(const mmb_attn_args* a, cudaStream_t stream);                        // attn_fwd_ws.cu
int launch_attn_bwd_tc(const mmb_attn_args* a, cudaStream_t stream);          // attn_bwd_tc.cu

__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ dctx, const __nv_bfloat16* __restrict__ ctx,
                     const float* __restrict__ lse, const float* __restrict__ keybias, uint4* __restrict__ ws,
                     int total_rows, int H, int nheads, int drop, uint64_t seed, uint32_t rng_stream) {
    // one 8-lane group per (row, head): 8 lanes x 8 bf16 = 64
    const int gidx = (blockIdx.x * 256 + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    const bool live = gidx < total_rows * nheads;
    const int row = live ? gidx / nheads : 0, head = live ? gidx - row * nheads : 0;
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
    if (!live) return;
    const uint32_t prob = (uint32_t)head * (uint32_t)total_rows + (uint32_t)row;
    const int64_t o = (int64_t)head * total_rows + row;
    if (sub == 0) {
        ws[o] = make_uint4(__float_as_uint(-lse[o]), __float_as_uint(-s), drop ? attn_drop_qkey(seed, rng_stream, prob) : 1u, 0u);
    } else if (sub == 1) {
        ws[(int64_t)nheads * total_rows + o] =
            make_uint4(__float_as_uint(keybias[row] * kLog2e), drop ? attn_drop_kkey(seed, rng_stream, prob) : 1u, 0u, 0u);
    }
}

static int check_args(const mmb_attn_args* a) {
    MMB_REQUIRE(a && a->qkv && a->keybias && a->cu_seqlens, "attn: null pointer");
    MMB_REQUIRE(a->H > 0 && a->nheads > 0 && a->H == a->nheads * kD, "attn: head dim must be 64 (H=%d heads=%d)", a->H,
                a->nheads);
    MMB_REQUIRE(a->nseq > 0 && a->max_seqlen > 0 && a->total_rows > 0, "attn: empty batch");
    MMB_REQUIRE(a->p_drop >= 0.f && a->p_drop < 1.f, "attn: p_drop=%f out of range", (double)a->p_drop);
    return MMB_OK;
}

}  // namespace mmb

using namespace mmb;

extern "C" size_t mmb_attn_bwd_workspace_bytes(int total_rows, int nheads) {
    if (total_rows <= 0 || nheads <= 0) return 0;
    return 2 * (size_t)nheads * (size_t)total_rows * 16;
}

extern "C" int mmb_attn_fwd(const mmb_attn_args* a, void* stream) {
    int rc = check_args(a);
    if (rc != MMB_OK) return rc;
    MMB_REQUIRE(a->ctx != nullptr, "attn_fwd: null ctx");
    // two forward kernels with the same contract: the persistent warp-specialised one (attn_fwd_ws.cu) and the
    // 2-CTA-per-SM one (attn_tc.cu); flags bit 1 / bit 2 force one of them, MMB_ATTN_FWD=ws|tc sets the default (A/B runs)
    static const int dflt = [] {
        const char* e = getenv("MMB_ATTN_FWD");
        return (e != nullptr && e[0] == 't') ? 0 : ((e != nullptr && e[0] == 'w') ? 1 : MMB_ATTN_FWD_DEFAULT_WS);
    }();
    const bool ws = (a->flags & 2) ? true : ((a->flags & 4) ? false : dflt != 0);
    return ws ? launch_attn_fwd_ws(a, (cudaStream_t)stream) : launch_attn_fwd_tc(a, (cudaStream_t)stream);
}

extern "C" int mmb_attn_bwd(const mmb_attn_args* a, void* stream) {
    int rc = check_args(a);
    if (rc != MMB_OK) return rc;
    MMB_REQUIRE(a->ctx && a->lse && a->dctx && a->dqkv && a->bwd_ws, "attn_bwd: null pointer");
    MMB_REQUIRE(((uintptr_t)a->bwd_ws % 16) == 0, "attn_bwd: workspace must be 16-byte aligned");
    const long long groups = (long long)a->total_rows * a->nheads;
    attn_bwd_prep_kernel<<<(unsigned)((groups * 8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)a->dctx, (const __nv_bfloat16*)a->ctx, a->lse, a->keybias, (uint4*)a->bwd_ws, a->total_rows,
        a->H, a->nheads, a->p_drop > 0.f ? 1 : 0, a->seed, a->rng_stream);
    rc = check_launch("attn_bwd_prep_kernel");
    if (rc != MMB_OK) return rc;
    return launch_attn_bwd_tc(a, (cudaStream_t)stream);
}
