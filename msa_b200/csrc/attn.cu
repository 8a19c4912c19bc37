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
                     int total_rows, int H, int nheads, int drop, uint64_t seed, uint32_t rng_stream,
                     const int* __restrict__ row_list) {
    pdl_trigger();
    pdl_wait();
    // one 8-lane group per (row, head): 8 lanes x 8 bf16 = 64; with a row list only the rows the backward kernels read
    // (live rows and the rest of their 128-row tiles: the first n_live + n_tile entries).  Persistent warps walk the groups
    // (row-major: a warp reads 512 contiguous bytes of dctx / ctx) with a grid stride: ~27 600 one-shot CTAs cost more in
    // scheduling than the kernel's ~35 us of HBM time.
    const int sub = threadIdx.x & 7, lane = threadIdx.x & 31;
    const int nrows = row_list != nullptr ? __ldg(row_list) + __ldg(row_list + 1) : total_rows;
    const int ngroups = nrows * nheads;
    const int nwarps = gridDim.x * 8;
    for (int base = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 4; base < ngroups; base += nwarps * 4) {
        const int gidx = base + (lane >> 3);
        const bool live = gidx < ngroups;
        const int li = live ? gidx / nheads : 0, head = live ? gidx - li * nheads : 0;
        const int row = row_list != nullptr ? __ldg(row_list + 4 + li) : li;
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
        if (!live) continue;
        const uint32_t prob = (uint32_t)head * (uint32_t)total_rows + (uint32_t)row;
        const int64_t o = (int64_t)head * total_rows + row;
        // lane 0 writes the row record, lane 1 the column record; both hash in ONE instruction stream (two divergent branches
        // would run the ~25-instruction hash twice per warp — the kernel was issue-bound at 66 %, DRAM 49 %)
        uint32_t key = 1u;
        if (drop && sub < 2)
            key = rng_row_key(sub == 0 ? seed : (seed ^ 0x9E3779B97F4A7C15ull), sub == 0 ? rng_stream : (rng_stream ^ 0x5bd1e995u), prob) | 1u;
        if (sub == 0) {
            ws[o] = make_uint4(__float_as_uint(-lse[o]), __float_as_uint(-s), key, 0u);
        } else if (sub == 1) {
            ws[(int64_t)nheads * total_rows + o] = make_uint4(__float_as_uint(keybias[row] * kLog2e), key, 0u, 0u);
        }
    }
}

// Work lists of the persistent kernels (mmb_attn_schedule; record layout in the header).  One CTA: the sequences are
// rank-sorted in shared memory (O(nseq^2 / threads) comparisons: ~40 at 3 x 64 sequences), one thread runs the prefix
// sum over the sorted order, then every sequence writes its own records.
constexpr int kSchedThreads = 1024;
constexpr int kSchedMaxSeqs = 8192;
__global__ void __launch_bounds__(kSchedThreads)
attn_schedule_kernel(const int* __restrict__ cu, const int* __restrict__ kv_end, const int* __restrict__ row_label,
                     int4* __restrict__ work, int nseq, int nheads, int cap, int* __restrict__ row_list) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ int sm[];
    int* len = sm;                 // sequence length
    int* eff = sm + nseq;          // keys before the all-masked tail
    int* ord_q = sm + 2 * nseq;    // rank -> sequence, by eff (forward / dQ items)
    int* ord_kv = sm + 3 * nseq;   // rank -> sequence, by len (dK/dV items)
    int* off_q = sm + 4 * nseq;    // sequence -> first record
    int* off_kv = sm + 5 * nseq;
    const int tid = threadIdx.x;
    for (int i = tid; i < nseq; i += kSchedThreads) {
        const int S = cu[i + 1] - cu[i];
        const int e = kv_end != nullptr ? kv_end[i] : 0;
        len[i] = S;
        eff[i] = (e > 0 && e < S) ? e : S;
    }
    __syncthreads();
    // zero-gradient query tail (header): holds iff no row at or behind its sequence's kv_end carries a label
    int labelled_tail = row_label == nullptr ? 1 : 0;
    if (row_label != nullptr) {
        // one warp per sequence (the tails are short and the loads independent: 32 sequences in flight)
        const int warp = tid >> 5, lane = tid & 31;
        for (int i = warp; i < nseq; i += kSchedThreads / 32) {
            const int r1 = cu[i + 1];
            for (int r = cu[i] + eff[i] + lane; r < r1; r += 32) labelled_tail |= __ldg(row_label + r) != -100 ? 1 : 0;
        }
    }
    const bool qskip = __syncthreads_or(labelled_tail) == 0;
    for (int i = tid; i < nseq; i += kSchedThreads) {
        const int ei = eff[i], li = len[i];
        int rq = 0, rkv = 0;
        for (int j = 0; j < nseq; ++j) {
            const int ej = eff[j], lj = len[j];
            rq += (ej > ei || (ej == ei && j < i)) ? 1 : 0;
            rkv += (lj > li || (lj == li && j < i)) ? 1 : 0;
        }
        ord_q[rq] = i;
        ord_kv[qskip ? rq : rkv] = i;      // query tail skipped: a dK/dV item costs ceil(eff / 64) steps, not ceil(len / 64)
    }
    __syncthreads();
    int* off_z = ord_q;            // sequence -> first record of its fully masked key tiles (reuses ord_q once consumed)
    if (tid == 0) {
        int nq = 0, nkv = 0;
        for (int r = 0; r < nseq; ++r) {
            const int iq = ord_q[r], ikv = ord_kv[r];
            off_q[iq] = nq;
            nq += nheads * ((len[iq] + 127) / 128);
            off_kv[ikv] = nkv;
            nkv += nheads * ((eff[ikv] + 127) / 128);
        }
        int nz = nkv;              // key tiles behind kv_end: dK = dV = 0, the kernel only stores zeros — they go last
        for (int i = 0; i < nseq; ++i) {
            off_z[i] = nz;
            nz += nheads * ((len[i] + 127) / 128 - (eff[i] + 127) / 128);
        }
        work[0] = make_int4(nq, nz, cap | (qskip ? (1 << 30) : 0), nkv);
    }
    __syncthreads();
    // records: (head, tile) in index order within a sequence, so that the CTAs working side by side share K / V in L2
    for (int ih = tid; ih < nseq * nheads; ih += kSchedThreads) {      // one thread per (sequence, head)
        const int i = ih / nheads, h = ih - i * nheads;
        const int row0 = cu[i], S = len[i], e = eff[i];
        const int tq = (S + 127) / 128, tkv = (e + 127) / 128;
        int4* wq = work + 1 + off_q[i] + h * tq;
        int4* wkv = work + 1 + cap + off_kv[i] + h * tkv;
        int4* wz = work + 1 + cap + off_z[i] + h * (tq - tkv);
        for (int t = 0; t < tq; ++t) wq[t] = make_int4(row0, S, e, (h << 16) | t);
        for (int t = 0; t < tkv; ++t) wkv[t] = make_int4(row0, S, e, (h << 16) | t);
        for (int t = tkv; t < tq; ++t) wz[t - tkv] = make_int4(row0, S, e, (h << 16) | t);
    }
    if (row_list == nullptr) return;
    // Row list (header: mmb_attn_schedule_args.row_list): live rows | rest of their 128-row tile | dead rows.  Without the
    // verified premise every row is live.  The offset arrays of the work lists are dead by now and are reused.
    __syncthreads();
    int* off_live = ord_q;
    int* off_tile = ord_kv;
    int* off_dead = off_q;
    __shared__ int tot[2];
    if (tid == 0) {
        int nl = 0, nt = 0, nd = 0;
        for (int i = 0; i < nseq; ++i) {
            const int S = len[i], e = qskip ? eff[i] : S;
            const int te = min(S, (e + 127) / 128 * 128);
            off_live[i] = nl;
            off_tile[i] = nt;
            off_dead[i] = nd;
            nl += e;
            nt += te - e;
            nd += S - te;
        }
        tot[0] = nl;
        tot[1] = nt;
        row_list[0] = nl;
        row_list[1] = nt;
        row_list[2] = nl + nt + nd;
        row_list[3] = qskip ? 1 : 0;
    }
    __syncthreads();
    {
        const int warp = tid >> 5, lane = tid & 31;
        int* live = row_list + 4;
        int* tile = live + tot[0];
        int* dead = tile + tot[1];
        int* flag = live + cu[nseq];           // per-row flags behind the list (mmb_gemm_args.row_live)
        for (int i = warp; i < nseq; i += kSchedThreads / 32) {
            const int row0 = cu[i], S = len[i], e = qskip ? eff[i] : S;
            const int te = min(S, (e + 127) / 128 * 128);
            for (int j = lane; j < S; j += 32) {
                flag[row0 + j] = j < e ? 1 : 0;
                if (j < e) live[off_live[i] + j] = row0 + j;
                else if (j < te) tile[off_tile[i] + (j - e)] = row0 + j;
                else dead[off_dead[i] + (j - te)] = row0 + j;
            }
        }
    }
}

static int check_args(const mmb_attn_args* a) {
    MMB_REQUIRE(a && a->qkv && a->keybias && a->cu_seqlens, "attn: null pointer");
    MMB_REQUIRE(a->H > 0 && a->nheads > 0 && a->H == a->nheads * kD, "attn: head dim must be 64 (H=%d heads=%d)", a->H,
                a->nheads);
    MMB_REQUIRE(a->nseq > 0 && a->max_seqlen > 0 && a->total_rows > 0, "attn: empty batch");
    MMB_REQUIRE(a->p_drop >= 0.f && a->p_drop < 1.f, "attn: p_drop=%f out of range", (double)a->p_drop);
    MMB_REQUIRE(((uintptr_t)a->work % 16) == 0, "attn: work lists must be 16-byte aligned");
    return MMB_OK;
}

}  // namespace mmb

using namespace mmb;

extern "C" size_t mmb_attn_bwd_workspace_bytes(int total_rows, int nheads) {
    if (total_rows <= 0 || nheads <= 0) return 0;
    return 2 * (size_t)nheads * (size_t)total_rows * 16;
}

extern "C" size_t mmb_attn_schedule_bytes(int nseq, int nheads, int max_seqlen) {
    if (nseq <= 0 || nheads <= 0 || max_seqlen <= 0) return 0;
    const size_t cap = (size_t)nseq * (size_t)nheads * (size_t)((max_seqlen + 127) / 128);
    return (1 + 2 * cap) * 16;
}

extern "C" size_t mmb_row_list_ints(int rows) { return rows > 0 ? 4 + 2 * (size_t)rows : 0; }

extern "C" int mmb_attn_schedule(const mmb_attn_schedule_args* a, void* stream) {
    MMB_REQUIRE(a && a->cu_seqlens && a->work, "attn_schedule: null pointer");
    MMB_REQUIRE(a->nseq > 0 && a->nseq <= kSchedMaxSeqs, "attn_schedule: nseq=%d not in [1, %d]", a->nseq, kSchedMaxSeqs);
    MMB_REQUIRE(a->nheads > 0 && a->nheads < 32768 && a->max_seqlen > 0, "attn_schedule: bad nheads / max_seqlen");
    MMB_REQUIRE(((uintptr_t)a->work % 16) == 0, "attn_schedule: work must be 16-byte aligned");
    const size_t cap = (size_t)a->nseq * (size_t)a->nheads * (size_t)((a->max_seqlen + 127) / 128);
    MMB_REQUIRE(cap < (1u << 30), "attn_schedule: too many work items");
    const int smem = 6 * a->nseq * (int)sizeof(int);
    if (smem > 48 * 1024)
        MMB_CUDA(cudaFuncSetAttribute(attn_schedule_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    launch_pdl(attn_schedule_kernel, dim3(1), dim3(kSchedThreads), (size_t)smem, (cudaStream_t)stream, a->cu_seqlens, a->kv_end,
               a->row_label, (int4*)a->work, a->nseq, a->nheads, (int)cap, a->row_list);
    return check_launch("attn_schedule_kernel");
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
    const long long ctas = (groups * 8 + 255) / 256;
    launch_pdl(attn_bwd_prep_kernel, dim3((unsigned)(ctas < (long long)num_sms() * 8 ? ctas : (long long)num_sms() * 8)), dim3(256), 0,
               (cudaStream_t)stream,
               (const __nv_bfloat16*)a->dctx, (const __nv_bfloat16*)a->ctx, a->lse, a->keybias, (uint4*)a->bwd_ws, a->total_rows,
               a->H, a->nheads, a->p_drop > 0.f ? 1 : 0, a->seed, a->rng_stream, a->row_list);
    rc = check_launch("attn_bwd_prep_kernel");
    if (rc != MMB_OK) return rc;
    return launch_attn_bwd_tc(a, (cudaStream_t)stream);
}
