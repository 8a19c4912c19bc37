/*
 * mmbert_sm100.h — C ABI of libmmbert_sm100.so, the B200 (sm_100a) implementation of the MMBert
 * forward/backward hot path of kimkyeonghun/MSA.
 *
 * The reference has no FFI of its own: its hot path is Python calling torch/transformers modules
 * (MMBertForPretraining.py:392-449, MMBertEmbedding.py:57-72, transformers/models/bert/modeling_bert.py).
 * Each entry point below replaces the library dispatches behind one of those call sites; the site is
 * cited on every declaration.  The Python binding (msa_b200/capi.py, ctypes) is the only caller.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - every op: int mmb_<op>(const mmb_<op>_args*, void* cuda_stream) -> MMB_OK (0) or a negative status;
 *     mmb_last_error() returns a thread-local message.
 *   - the caller owns ALL device memory (inputs, outputs, saved-for-backward buffers, workspaces);
 *     the library never allocates device memory, never synchronises the device and only ever
 *     launches on the stream it is given.
 *   - pointers are device pointers, 16-byte aligned, row-major; leading dimensions are in elements.
 *   - bf16 = __nv_bfloat16 storage (uint16_t here), f32 = float.
 */
#ifndef MMBERT_SM100_H
#define MMBERT_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMB_VERSION 204 /* round 2; 200: mmb_pack_args.mask_frame_stride; 201: MMB_EPI_CE_STATS, mmb_gemm_args.aux2,
                           mmb_pack_args.row_label / vocab, mmb_ce_sparse_*, mmb_embed_args.err_count;
                           202: mmb_attn_schedule_args.row_label (zero-gradient query tail skipped by the backward);
                           203: mmb_gemm_args.colsum;
                           204: mmb_attn_args.flags bit 3, row lists (mmb_attn_schedule_args.row_list and the row_list
                           members of the row kernels), mmb_gemm_args.row_live / dead_rows_zeroed, mmb_embed_args.row_live */

enum mmb_status {
    MMB_OK = 0,
    MMB_EINVAL = -1, /* bad shape / alignment / null pointer */
    MMB_EARCH = -2,  /* device is not sm_100 */
    MMB_ECUDA = -3   /* CUDA runtime / driver failure, see mmb_last_error() */
};

int mmb_version(void);
const char* mmb_last_error(void);
/* MMB_OK when the current device is compute capability 10.x, MMB_EARCH otherwise. */
int mmb_check_device(void);
/* Number of CUDA kernels this library has launched in the calling process (monotonic; bench.py reports the
 * per-step difference as gpu_launches). */
long long mmb_launch_count(void);
/* Data-parallel runs (msa_b200/ddp.py; the reference has no distributed code): the persistent kernels (GEMM, attention:
 * one CTA or CTA pair per SM) leave n SMs idle so that the concurrent NCCL all-reduce kernels start at once on free
 * SMs instead of waiting for, and then delaying, a persistent wave.  0 (default; also the MMB_RESERVE_SMS environment
 * variable) = use every SM.  Process-wide. */
int mmb_set_reserved_sms(int n);

/* ------------------------------------------------------------------------------------------------
 * Dense bf16 GEMM on tcgen05 tensor cores (TMA -> smem ring -> tcgen05.mma -> TMEM -> epilogue).
 *   C[M,N] = epilogue( alpha * sum_k A[m,k] * B[n,k] )
 * Replaces every nn.Linear on the path and its autograd backward:
 *   modeling_bert.py:179-181 (Q,K,V), :295 (attention output), :340 (intermediate), :353 (output),
 *   :482 (LM transform), :500 (tied decoder); MMBertEmbedding.py:62,64 is NOT routed here (fused embed).
 * Operand storage:
 *   a_major == MMB_MAJOR_K : A is [M,K] row-major (lda = row stride)     — activations / dY for dgrad
 *   a_major == MMB_MAJOR_MN: A is [K,M] row-major (lda = row stride)     — dY^T for wgrad, read in place
 *   b_major likewise for B ([N,K] or [K,N]).
 * Requirements: lda, ldb, ldc multiples of 8; K > 0; pointers 16-byte aligned.  M, N, K arbitrary
 * otherwise (TMA zero-fills out-of-bounds, stores are bounds-checked).
 */
enum { MMB_MAJOR_K = 0, MMB_MAJOR_MN = 1 };
enum {
    MMB_EPI_STORE_BF16 = 0,     /* C(bf16) = acc + bias                                             */
    MMB_EPI_GELU_BF16 = 1,      /* aux(bf16) = acc + bias (if aux != NULL); C(bf16) = gelu_erf(...)  */
    MMB_EPI_RELU_BF16 = 2,      /* C(bf16) = relu(acc + bias)                                        */
    MMB_EPI_STORE_F32 = 3,      /* C(f32)  = acc + bias                                              */
    MMB_EPI_ATOMIC_ADD_F32 = 4, /* C(f32) += acc   (red.global.add; split-K capable; bias ignored)   */
    MMB_EPI_DGELU_BF16 = 5,     /* C(bf16) = acc * gelu_erf'(aux[m,n]); aux is an INPUT (pre-act)    */
    MMB_EPI_GELU_GRAD_BF16 = 6, /* C(bf16) = gelu_erf(acc + bias); aux(bf16, out) = gelu_erf'(acc + bias): the
                                   forward of BertIntermediate saves the activation's derivative instead of the
                                   pre-activation, so that its backward is a plain multiply (next mode)   */
    MMB_EPI_MUL_AUX_BF16 = 7,   /* C(bf16) = acc * aux[m,n]; aux is an INPUT                           */
    MMB_EPI_CE_STATS = 8        /* tied-decoder GEMM fused with the masked-LM cross entropy: the [M,N] logits are NOT
                                   written.  aux = int32 row_label[M] (vocabulary id or -100, from mmb_pack_prepare);
                                   for every labelled row the epilogue keeps an online (max, sum exp) of acc + bias per
                                   128-column group, the label's logit, and — if C != NULL — stores that row's logits
                                   (bf16, row stride ldc) into C, which is the buffer mmb_ce_sparse_bwd turns into
                                   dlogits in place.  aux2 = float workspace of mmb_ce_stats_floats(M, N) floats (planes
                                   of M: max / sum per group, label logit, row-written flag; zero it once, the flag
                                   plane must persist across steps).  Unlabelled rows cost no epilogue work.  N > 128. */
};

typedef struct mmb_gemm_args {
    const void* A;
    const void* B;
    void* C;
    void* aux;         /* see epilogue; row stride = ldaux */
    void* aux2;        /* MMB_EPI_CE_STATS workspace, else NULL */
    const float* bias; /* [N] f32 or NULL */
    int64_t lda, ldb, ldc, ldaux;
    int32_t M, N, K;
    int32_t a_major, b_major;
    int32_t epilogue;
    int32_t split_k; /* >= 1; > 1 only with MMB_EPI_ATOMIC_ADD_F32 */
    float alpha;
    /* debug overrides for bring-up (0 = use built-in values) */
    int32_t dbg_flags;
    float* colsum; /* NULL, or [N] f32: colsum[n] += sum over rows of the bf16-ROUNDED C[:, n] (the bias gradient of the
                      Linear whose output gradient C is: exactly what its wgrad GEMM reads).  bf16-output epilogues
                      only; taken from the epilogue's staging tiles of the CTA-pair kernel, else by a mmb_colsum_bf16
                      launch behind the GEMM */
    const int32_t* row_live; /* NULL, or int32 [M]: 0 marks an output row nobody reads (a padding row, see
                      mmb_attn_schedule_args.row_list: its per-row flags).  A hint the epilogue MAY use, in units of its
                      32-row slices: MMB_EPI_STORE_BF16 (without colsum) / MMB_EPI_GELU_BF16 / MMB_EPI_GELU_GRAD_BF16 leave C / aux of
                      an all-dead slice unwritten (no bias / activation math, no stores), MMB_EPI_MUL_AUX_BF16 writes zeros there without reading aux (its wgrad consumer
                      reads every row).  The matrix product itself is computed for every tile.  Other epilogues ignore it. */
    int32_t dead_rows_zeroed; /* MMB_EPI_MUL_AUX_BF16 with row_live: non-zero = the caller guarantees that the all-dead
                      slices of C already hold zeros (an earlier launch of this step into the same buffer under the same
                      flags): they are skipped instead of zero-filled */
} mmb_gemm_args;

int mmb_gemm(const mmb_gemm_args* a, void* stream);
/* floats of the MMB_EPI_CE_STATS workspace for an [M, N] decoder GEMM */
size_t mmb_ce_stats_floats(int M, int N);

/* ------------------------------------------------------------------------------------------------
 * dropout + residual + LayerNorm (warp-per-row, HBM-bound).
 *   out = LayerNorm_eps( dropout_p(y) + res ) * gamma + beta ;  mean/rstd saved for backward.
 * Replaces BertSelfOutput.forward (modeling_bert.py:295-297), BertOutput.forward (:353-355) and, with
 * res == NULL and p_drop == 0, the LayerNorm of BertPredictionHeadTransform (:484).  The dense bias is
 * added by the producing GEMM's epilogue.  Dropout masks come from a counter-based generator keyed by
 * (seed, rng_stream, row * H + col): backward regenerates them, nothing is stored.
 */
typedef struct mmb_drln_fwd_args {
    const void* y;    /* [M,H] bf16 dense output (bias included) */
    const float* res; /* [M,H] f32 residual stream or NULL */
    const float* gamma;
    const float* beta;
    void* out;      /* [M,H] bf16 (GEMM operand copy) */
    float* out_f32; /* [M,H] f32 residual-stream copy, or NULL */
    float* mean; /* [M] */
    float* rstd; /* [M] */
    int32_t M, H;
    float eps;
    float p_drop;
    uint64_t seed;
    uint32_t rng_stream;
    int32_t y_f32; /* fp32 validation path: y is [M,H] f32 and out may be NULL (only out_f32 is written) */
    const int32_t* row_list; /* NULL, or mmb_attn_schedule's row list: only its live rows are computed, every other row of
                                out / out_f32 / mean / rstd keeps its previous contents (see mmb_attn_schedule_args) */
} mmb_drln_fwd_args;
int mmb_dropout_residual_ln_fwd(const mmb_drln_fwd_args* a, void* stream);

/* Backward of the above (autograd of modeling_bert.py:295-297 / :353-355).
 *   dz = LayerNormBackward(g1 + g2);  d_res = dz;  d_y = dz * mask / (1-p)
 *   dgamma += sum_rows (g1+g2) * xhat;  dbeta += sum_rows (g1+g2);  dbias += sum_rows d_y   (fp32 atomics)
 * With gelu_aux (LM-head transform: LayerNorm(gelu(dense(x))), modeling_bert.py:482-484) d_y is additionally
 * multiplied by gelu'(aux) so that it is the gradient of the dense output.
 */
typedef struct mmb_drln_bwd_args {
    const void* g1;  /* [M,H] bf16 gradient of out (from the consuming GEMM's dgrad) */
    const float* g2; /* [M,H] f32 gradient of out through the residual stream, or NULL */
    const void* y;
    const float* res; /* f32 or NULL */
    const float* mean;
    const float* rstd;
    const float* gamma;
    void* d_y;     /* [M,H] bf16 */
    float* d_res;  /* [M,H] f32 or NULL */
    float* dgamma; /* [H] accumulated */
    float* dbeta;  /* [H] accumulated */
    float* dbias;  /* [H] accumulated, or NULL */
    const void* gelu_aux; /* [M,H] bf16 or NULL: y = gelu(aux) -> d_y is chained through gelu'(aux) */
    int32_t M, H;
    float p_drop;
    uint64_t seed;
    uint32_t rng_stream;
    const int32_t* row_list; /* NULL, or mmb_attn_schedule's row list: the live rows are computed, d_y / d_res of every
                                other row are set to zero without reading anything (their gradient is exactly zero) */
    int32_t dead_rows_zeroed; /* with row_list: non-zero = the caller guarantees that the non-live rows of d_y and d_res
                                already hold zeros (an earlier launch of this step wrote them into the same buffers under
                                the same list and nothing else writes there): they are then not written again */
} mmb_drln_bwd_args;
int mmb_dropout_residual_ln_bwd(const mmb_drln_bwd_args* a, void* stream);

/* out[n] += sum_m X[m,n] — bias gradients of the dense layers whose dY is produced by a GEMM / attention
 * kernel (autograd of nn.Linear bias, modeling_bert.py:179-181, :340). */
typedef struct mmb_colsum_args {
    const void* X; /* [M,N] bf16 */
    float* out;    /* [N] f32, accumulated */
    int64_t ld;
    int32_t M, N;
    const int32_t* row_list; /* NULL, or mmb_attn_schedule's row list: only the live rows are summed (the others are zero) */
} mmb_colsum_args;
int mmb_colsum_bf16(const mmb_colsum_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Masked multi-head self-attention over packed variable-length sequences, head dim 64.
 *   ctx = dropout_p(softmax(Q K^T / 8 + keybias)) V      per (sequence, head)
 * Replaces BertSelfAttention.forward's attention_interface call (modeling_bert.py:194-206 ->
 * eager_attention_forward :115-140) with the extended mask built at MMBertForPretraining.py:246-250.
 * Q|K|V are read in place from the fused projection output qkv [rows, 3H] (Q at column h*64, K at
 * H + h*64, V at 2H + h*64); keybias[row] = (1 - mask) * -10000 for the key stored at that packed row.
 * lse ([heads, rows] f32, log2 domain) is saved for backward.  mmb_attn_bwd writes dQ|dK|dV into dqkv with
 * the same layout (autograd of the same lines); bwd_ws is caller-provided scratch of
 * mmb_attn_bwd_workspace_bytes(total_rows, nheads) bytes, 16-byte aligned (per-(head, row) records: -LSE,
 * -rowsum(dO * O), dropout keys and the scaled key bias, laid out for the kernels' TMA producer).
 * Dropout masks are a function of (seed, rng_stream, head, packed row of the query, packed row of the key): forward
 * and backward regenerate identical masks, nothing is stored.
 */
typedef struct mmb_attn_args {
    const void* qkv;        /* [rows, 3H] bf16 */
    void* ctx;              /* [rows, H] bf16 (output of fwd, input of bwd) */
    float* lse;             /* [heads, rows] */
    const float* keybias;   /* [rows] */
    const int32_t* cu_seqlens; /* [nseq + 1] row offsets */
    const void* dctx;       /* bwd: [rows, H] bf16 */
    void* dqkv;             /* bwd: [rows, 3H] bf16 */
    void* bwd_ws;           /* bwd scratch: mmb_attn_bwd_workspace_bytes(total_rows, nheads) bytes */
    const int32_t* kv_end;  /* [nseq] or NULL: keys at index >= kv_end[seq] are all masked (bias -10000) and are
                               skipped in whole tiles — exact in fp32: exp(-10000 + x - max) == 0 whenever an
                               unmasked key exists; a sequence with kv_end == 0 is processed in full */
    int32_t H, nheads, nseq, max_seqlen, total_rows;
    float p_drop;
    uint64_t seed;
    uint32_t rng_stream;
    uint32_t flags; /* 0 = default; bring-up: bit 1 forces the persistent warp-specialised forward kernel, bit 2 the other one.
                       bit 3 (value 8), forward with work lists only: if mmb_attn_schedule set the zero-gradient-tail bit
                       (no row at or behind kv_end carries a label), the 128-query tiles that lie ENTIRELY behind kv_end
                       are not computed and their ctx / lse rows keep their previous contents.  Those rows are padding:
                       masked keys for every query of every layer and never read by a loss, so every value the model
                       returns (and every gradient) is unchanged bit for bit — callers that read the hidden states or the
                       vocabulary logits of padded positions (materialised pred_t / pred_v / pred_s) must leave it clear.
                       ctx must hold finite values (e.g. zero-initialised once): the next GEMM still reads those rows.
                       bit 4 (value 16), backward with work lists only: the caller guarantees that the dqkv rows of the 128-row
                       tiles that lie entirely behind kv_end already hold zeros (an earlier mmb_attn_bwd of this step wrote
                       them into the same buffer for the same batch); with the zero-gradient-tail bit set those tiles are
                       then not written again. */
    const void* work;       /* NULL, or the work lists written by mmb_attn_schedule for the same cu_seqlens / kv_end /
                               nheads / max_seqlen: the persistent kernels then take their (sequence, head, tile) items
                               longest first instead of in index order (same results bit for bit; evens out the CTAs) */
    const int32_t* row_list; /* NULL, or mmb_attn_schedule's row list: mmb_attn_bwd prepares its per-row records only for
                                the rows the backward kernels read (live rows and the rest of their 128-row tile) */
} mmb_attn_args;
size_t mmb_attn_bwd_workspace_bytes(int total_rows, int nheads);
int mmb_attn_fwd(const mmb_attn_args* a, void* stream);
int mmb_attn_bwd(const mmb_attn_args* a, void* stream);

/* Work lists for the persistent attention kernels, built on the device once per batch (the sequence lengths and
 * kv_end are device data: no host round trip) and shared by every layer's forward and backward launch.
 * work = 16-byte records: [0] = {count_q, count_kv, cap, count_kv_unmasked}, cap = nseq * nheads * ceil(max_seqlen / 128);
 * [1, 1 + cap): one record {first row, length, effective key count, head << 16 | tile} per 128-query tile of every
 * (sequence, head), sequences ordered by effective key count (length, or kv_end where that is smaller), longest first
 * — an item of the forward and of the dQ pass costs ceil(effective keys / 64) steps; [1 + cap, 1 + 2 cap): the 128-key
 * tiles of the dK/dV pass, first those that hold an unmasked key (sequences ordered by length: ceil(length / 64) steps
 * each), then the fully masked ones (zero fill only).  nseq <= 8192.
 *
 * Zero-gradient query tail (row_label != NULL).  In this model the loss reaches the encoder output only at rows that
 * carry a masked-LM label and at row 0 of every sequence (pooler / alignment heads, MMBertForPretraining.py:295-302,
 * 406-425).  A row at or behind kv_end is a masked key for every query, so no gradient reaches it through K or V
 * either: if no such row is labelled, dctx is EXACTLY zero on the rows >= kv_end of every layer, and those query rows
 * contribute exactly nothing to dQ (their own rows: zero), dK and dV.  mmb_attn_schedule verifies the premise on the
 * device (every row_label at or behind kv_end is -100) and, if it holds, sets bit 30 of the header's cap field and
 * orders the dK/dV list by effective key count; mmb_attn_bwd then runs the dK/dV pass over ceil(kv_end / 64) query steps
 * instead of ceil(length / 64), and the dQ pass over the dK/dV list (query tiles behind kv_end only store zeros).
 * Same results bit for bit as the full sweep; the forward is unaffected (those rows' outputs are defined) unless the
 * caller sets mmb_attn_args.flags bit 3. */
typedef struct mmb_attn_schedule_args {
    const int32_t* cu_seqlens; /* [nseq + 1] */
    const int32_t* kv_end;     /* [nseq] or NULL */
    void* work;                /* out: mmb_attn_schedule_bytes(...) bytes, 16-byte aligned */
    int32_t nseq, nheads, max_seqlen;
    const int32_t* row_label;  /* [rows] from mmb_pack_prepare, or NULL (no query-tail skipping) */
    int32_t* row_list;         /* NULL, or out: int32 [4 + 2 rows] — the packed rows in three groups,
                                  [0] = n_live  rows before their sequence's kv_end,
                                  [1] = n_tile  rows at or behind kv_end that share a 128-row attention tile with a live row,
                                  [2] = rows, [3] = 1 if the zero-gradient-tail premise above holds (else every row is live),
                                  [4, 4 + rows) = the live rows (ascending), then the tile rows, then the rest;
                                  [4 + rows, 4 + 2 rows) = per-row flags, 1 = live (mmb_gemm_args.row_live).
                                  The buffer holds 4 + 2 rows ints.
                                  A row that is not live is padding: a masked key for every query of every layer whose
                                  value reaches no loss and whose gradient is exactly zero.  The row kernels (LayerNorm
                                  forward / backward, column sums, the attention backward's preparation) take the list to
                                  leave those rows alone; nothing the model returns changes.  Buffers whose dead rows are
                                  read by a GEMM must hold finite values (zero-initialised once). */
} mmb_attn_schedule_args;
size_t mmb_attn_schedule_bytes(int nseq, int nheads, int max_seqlen);
size_t mmb_row_list_ints(int rows); /* int32 elements of mmb_attn_schedule_args.row_list for a batch of `rows` packed rows */
int mmb_attn_schedule(const mmb_attn_schedule_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Packed batch: the reference's three encoder passes (MMBertForPretraining.py:402-404) are packed into
 * one variable-length batch of 3B sequences, rows ordered
 *   [pass 0: B x T text] [pass 1: B x (T + Lv) text|visual] [pass 2: B x (T + La) text|speech].
 * Input tensors are consumed in the dtypes the reference's collate produces (model_utils.py:117-142):
 * ids/labels int64, frames float64, masks float64 or int64 — selected by an MMB_DT_* code.
 */
enum { MMB_DT_F32 = 0, MMB_DT_F64 = 1, MMB_DT_I64 = 2, MMB_DT_I32 = 3, MMB_DT_U8 = 4 };

/* keybias[row] = (1 - mask) * -10000 (MMBertForPretraining.py:152-153; frame masks: feature 0 only, :75-77),
 * cu_seqlens[3B+1] row offsets, label_count[3] = number of labels != -100 per pass (CrossEntropyLoss
 * mean denominator, :381-383). */
typedef struct mmb_pack_args {
    const void* mask_text[3]; /* [B,T] text-half mask of each pass */
    int32_t mask_text_dtype[3];
    const void* mask_frame[2]; /* [B,L,D] visual / speech masks (only feature 0 is read, MMBertForPretraining.py:74-77) */
    int32_t mask_frame_dtype[2];
    int32_t frame_dim[2];
    int32_t mask_frame_stride[2]; /* elements between consecutive frames of mask_frame: 0 = frame_dim (a [B,L,D] mask),
                                     1 = a [B,L] mask (the reference takes those too, :74) */
    const void* labels[3]; /* int64 [B,T], [B,T+Lv], [B,T+La]; may be NULL */
    int32_t* row_label;    /* [rows] out or NULL: the label of every packed row as int32 (-100 = none), for the fused CE.
                              With vocab > 0 a label outside [0, vocab) counts as an ERROR: it is treated as -100 and
                              label_count[3] is incremented (mmb_heads_fwd then returns a NaN joint loss — torch's
                              CrossEntropyLoss device-asserts there) */
    int32_t vocab;
    float* keybias;        /* [rows] */
    int32_t* cu_seqlens;   /* [3B+1] */
    int32_t* label_count;  /* [4]: labelled rows per pass, [3] = count of out-of-range labels / ids */
    int32_t* kv_end;       /* [3B] or NULL: 1 + index of the last key whose mask is set, per sequence */
    int32_t B, T;
    int32_t L[2];
} mmb_pack_args;
int mmb_pack_prepare(const mmb_pack_args* a, void* stream);

/* Fused embeddings, forward and backward.
 *   text rows : LN_eps1(word[id] + type[tt] + pos[s]) -> dropout(p1) [-> LN_eps2 -> dropout(p2) in joint passes]
 *   frame rows: relu(W f32(frame) + b) (rounded to bf16, saved in pframe) -> LN_eps2 -> dropout(p2)
 * Replaces BertEmbeddings.forward (modeling_bert.py:103-112) and JointEmbeddings.forward
 * (MMBertEmbedding.py:61-70).  wT are the projection weights transposed to [D,H] (refreshed with the
 * bf16 weight copies).  mmb_embed_bwd consumes dx0 and ACCUMULATES (fp32 atomics) into the g_* buffers:
 * word-embedding rows (id 0 = padding_idx gets none), position, token-type, both LayerNorms, Wv/Ws and biases.
 */
typedef struct mmb_embed_args {
    const void* ids[3];     /* int64 [B,T]: text, text-with-visual, text-with-speech ids */
    const void* token_type; /* int64 [B,T] (pass 0; joint passes use type 0, MMBertForPretraining.py:223) */
    const void* frames[2];  /* [B,Lv,Dv], [B,La,Da] */
    int32_t frames_dtype[2];
    int32_t frame_dim[2];
    const float* word; /* [V,H] */
    const float* pos;  /* [max_pos,H] */
    const float* type; /* [2,H] */
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    const float* wT[2]; /* [D,H] */
    const float* wb[2]; /* [H] */
    float eps1, eps2, p_drop1, p_drop2;
    uint64_t seed;
    void* x0;      /* [rows,H] bf16 */
    float* x0_f32; /* [rows,H] f32 residual-stream copy, or NULL */
    float *mean1, *rstd1, *mean2, *rstd2; /* [rows] */
    void* pframe;                         /* [B*(Lv+La), H] bf16 */
    /* backward */
    const void* dx0;  /* [rows,H] bf16 */
    const float* dx0b; /* [rows,H] f32 residual-stream gradient to add, or NULL */
    void* dpre;      /* [B*(Lv+La), H] bf16 scratch */
    float *g_word, *g_pos, *g_type, *g_ln1_g, *g_ln1_b, *g_ln2_g, *g_ln2_b;
    float* g_w[2];  /* [H,D] */
    float* g_wb[2]; /* [H] */
    int32_t B, T;
    int32_t L[2];
    int32_t H, V, max_pos;
    int32_t exact_frames; /* fp32 validation path: keep relu(W f + b) in fp32 (no bf16 rounding) on the way into x0_f32 */
    /* optional (both or neither per modality): lets mmb_embed_bwd compute the projection weight gradient
     * g_w = dpre^T · frames on the tensor cores (mmb_gemm, split-K) instead of a CUDA-core kernel.
     *   frames_bf16[m]: [B*L[m], ldf] bf16 copy of the frames, ldf = frame_dim[m] rounded up to 8, written by mmb_embed_fwd
     *   gw_pad[m]:      [H, ldf] f32 scratch (zeroed and consumed by mmb_embed_bwd) */
    void* frames_bf16[2];
    float* gw_pad[2];
    int32_t* err_count; /* or NULL: incremented for every token id outside [0, V) (such an id reads the padding row 0
                           instead of out-of-bounds memory; torch's nn.Embedding device-asserts there) */
    const int32_t* row_live; /* NULL, or the per-row flags of mmb_attn_schedule's row list (mmb_gemm_args.row_live): the
                           forward leaves a block of 16 frames alone if none of its packed rows is live (x0, pframe and
                           frames_bf16 keep their previous contents there), the backward skips dead text rows and writes
                           dpre = 0 for dead frames without reading anything.  Text rows are always embedded (the id
                           range check covers every position). */
} mmb_embed_args;
int mmb_embed_fwd(const mmb_embed_args* a, void* stream);
int mmb_embed_bwd(const mmb_embed_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Masked-LM cross entropy over the tied-decoder logits (bf16 [rows, ldl], ldl = V rounded up to 8).
 * Replaces CrossEntropyLoss()(prediction_scores.view(-1,V), labels.view(-1)) at
 * MMBertForPretraining.py:381-384 (ignore_index -100, mean over the labelled rows of each pass).
 * fwd: row_lse[row] and loss_sum[pass] for labelled rows.  bwd: dlogits = coef * gscale / count[pass] *
 * (softmax - onehot) for labelled rows and (dense != 0) zeros elsewhere — the reference's autograd runs the
 * decoder dgrad/wgrad GEMMs over all rows; dense keeps that work faithful.
 */
typedef struct mmb_ce_args {
    const void* logits;    /* [rows, ldl] bf16 (unused by the mmb_ce_sparse_* entry points) */
    void* dlogits;         /* [rows, ldl] bf16 (bwd) */
    const void* labels[3]; /* int64 per pass */
    const int32_t* label_count; /* [3] from mmb_pack_prepare */
    float* row_lse;        /* [rows] */
    float* loss_sum;       /* [3] */
    const float* gscale;   /* device scalar: d(joint_loss) upstream gradient, NULL = 1 */
    float coef;            /* alpha / 3 */
    int32_t V;
    int64_t ldl;
    int32_t B, T;
    int32_t L[2];
    int32_t dense;
    int32_t logits_f32; /* fp32 validation path (mmb_ce_fwd only): logits are f32, ldl in f32 elements */
} mmb_ce_args;
int mmb_ce_fwd(const mmb_ce_args* a, void* stream);
int mmb_ce_bwd(const mmb_ce_args* a, void* stream);

/* The same cross entropy WITHOUT a materialised [rows, V] logits matrix (training default; SURVEY.md §2.1 "fused CE").
 * Forward: the decoder GEMM runs with MMB_EPI_CE_STATS; mmb_ce_sparse_fwd merges the per-group (max, sum exp) records of
 * the labelled rows into row_lse and loss_sum[pass] += lse - logit[label].
 * Backward: mmb_ce_sparse_bwd rewrites, in place in dlogits, the labelled rows (they hold their bf16 logits) into
 * coef * gscale / count[pass] * (softmax - onehot) and zeroes rows that were labelled in an earlier step but are not now
 * (row-written flags in the stats workspace) — every other row of dlogits is still zero from the plan's one-time
 * initialisation, so the dense decoder dgrad / wgrad GEMMs read exactly the matrix the reference's autograd forms.
 * mmb_colsum_rows_bf16 adds the column sums of the labelled rows to the decoder bias gradient (the unlabelled rows are
 * zero).  HBM traffic per step: the labelled rows only (~1 % of the rows) instead of four full passes over [rows, V]. */
typedef struct mmb_ce_sparse_args {
    const int32_t* row_label;   /* [rows] from mmb_pack_prepare */
    float* stats;               /* mmb_ce_stats_floats(rows, V) floats (the GEMM's aux2) */
    void* dlogits;              /* [rows, ldl] bf16 (bwd) */
    const int32_t* label_count; /* [3] */
    float* row_lse;             /* [rows] */
    float* loss_sum;            /* [3]: zeroed and filled by fwd */
    const float* gscale;        /* device scalar upstream gradient, NULL = 1 */
    float* dbias;               /* [V] f32 decoder-bias gradient, accumulated by mmb_colsum_rows_bf16 */
    float coef;                 /* alpha / 3 */
    int32_t V;
    int64_t ldl;
    int32_t B, T;
    int32_t L[2];
} mmb_ce_sparse_args;
int mmb_ce_sparse_fwd(const mmb_ce_sparse_args* a, void* stream);
int mmb_ce_sparse_bwd(const mmb_ce_sparse_args* a, void* stream);
int mmb_colsum_rows_bf16(const mmb_ce_sparse_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Pooler + alignment/NSP heads + score-attention fusion + sentiment head + CPC x3 + loss combination
 * on the 3B [CLS] rows.  Replaces BertPooler.forward (modeling_bert.py:462-468),
 * MMBertPreTrainingHeads.forward (MMBertForPretraining.py:295-302), the fusion/classifier block (:406-415),
 * CPC.forward (MMBertEmbedding.py:21-32) and the losses (:386-388, :427-443), plus autograd of all of it.
 * All weights fp32 in nn.Linear layout [out,in].  losses[8] = joint, mlm, ap, label, nce, mlm_t, mlm_v, mlm_s.
 * mmb_heads_bwd ACCUMULATES parameter gradients into g_* and ADDS the [CLS]-row gradients into dseq_out.
 * cls.seq_relationship is evaluated for its output only (never in a loss): it has no gradient entry.
 */
typedef struct mmb_heads_args {
    const void* seq_out;       /* [rows,H] bf16 encoder output */
    void* dseq_out;            /* [rows,H] bf16 gradient buffer (bwd: [CLS] rows are incremented) */
    const int32_t* cu_seqlens; /* [3B+1] */
    void* workspace;           /* mmb_heads_workspace_bytes(B,H) bytes, must persist fwd -> bwd */
    const float *w_pooler, *b_pooler, *w_seqrel, *b_seqrel, *w_align, *b_align, *w_attn, *b_attn;
    const float *w_c11, *b_c11, *w_c12, *b_c12;
    const float* w_v[3]; /* vt, vv, vs */
    const float* b_v[3];
    const float* w_cpc[3]; /* cpc_zt, cpc_zv, cpc_za .net */
    const float* b_cpc[3];
    float *g_w_pooler, *g_b_pooler, *g_w_align, *g_b_align, *g_w_attn, *g_b_attn, *g_w_c11, *g_b_c11, *g_w_c12, *g_b_c12;
    float* g_w_v[3];
    float* g_b_v[3];
    float* g_w_cpc[3];
    float* g_b_cpc[3];
    const void* ap_label[2]; /* int64 [B] visual, speech */
    const float* sentiment;  /* [B] f32 (regression target; class index as a float in the classification branch) */
    const float* ce_loss_sum;   /* [3] from mmb_ce_fwd */
    const int32_t* label_count; /* [3] */
    float* losses;     /* [8] */
    float* logits_out; /* [B]  tanh applied iff num_labels == 1; num_labels not in {1,7} = the reference's CrossEntropyLoss
                          branch (MMBertForPretraining.py:437-442) on the [B,1] classifier output: label loss 0 (NaN if a
                          target is not class 0, where torch raises), logits_out = argmax = 0, no label gradient */
    float* rel_out;    /* [B,2] seq_relationship(pooled_text) or NULL */
    float* align_out;  /* [2B,2] align scores (visual rows, then speech rows) or NULL */
    const float* gscale; /* device scalar upstream gradient, NULL = 1 */
    float alpha, beta;
    int32_t B, H, num_labels;
    int32_t seq_out_f32; /* fp32 validation path (mmb_heads_fwd only): seq_out is f32 */
} mmb_heads_args;
size_t mmb_heads_workspace_bytes(int B, int H);
int mmb_heads_fwd(const mmb_heads_args* a, void* stream);
int mmb_heads_bwd(const mmb_heads_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * On-device MLM masking.  Replaces model_utils.mask_tokens (model_utils.py:6-39), which the
 * reference runs three times per step (trainer.py:45-47) with host round trips: 15 % (prob) of the non-special tokens are
 * selected; labels = original id where selected, -100 elsewhere; 80 % (replace_prob) of the selected positions get
 * mask_id, the rest keep their token (the reference's random-word branch is commented out, :34-37).  ids are updated IN
 * PLACE like the reference's `inputs`.  Special tokens = the ids listed in special[] (BERT: [PAD] 0, [UNK] 100, [CLS] 101,
 * [SEP] 102, [MASK] 103 — tokenizer.all_special_ids).  Decisions come from the counter-based generator of common.cuh
 * keyed by (seed, rng_stream, element index): same seed -> same mask.  If labels_dup is not NULL it receives the
 * reference's cat((labels, labels), -1) of trainer.py:50,53 ([B, 2T]) in the same pass.
 */
typedef struct mmb_mlm_mask_args {
    void* ids;        /* int64 [B,T] in/out */
    void* labels;     /* int64 [B,T] out */
    void* labels_dup; /* int64 [B,2T] out or NULL */
    int32_t special[8];
    int32_t n_special;
    int32_t B, T;
    int32_t mask_id;
    float prob, replace_prob;
    uint64_t seed;
    uint32_t rng_stream;
} mmb_mlm_mask_args;
int mmb_mlm_mask(const mmb_mlm_mask_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * fp32 validation path (forward only).  BASELINE.json asks for an fp32 path within 1e-4 relative of the reference on
 * logits and loss; the tensor-core path above is bf16 by construction, so the same forward can also be run with fp32
 * storage and plain fp32 CUDA-core arithmetic: mmb_embed_fwd (exact_frames), mmb_linear_f32, mmb_attn_f32_fwd,
 * mmb_dropout_residual_ln_fwd (y_f32), mmb_ce_fwd (logits_f32), mmb_heads_fwd (seq_out_f32).  Correctness-first
 * kernels for parity checks at small sizes; they are not on the benchmarked path and have no backward.
 *
 * mmb_linear_f32: Y[M,N] = act(X[M,K] W[N,K]^T + bias[N])   (nn.Linear layout; modeling_bert.py:179-181, :295, :340, :353,
 * :482, :500) with act = MMB_ACT_*.   mmb_attn_f32_fwd: softmax(Q K^T / 8 + keybias) V per (sequence, head), no dropout
 * (modeling_bert.py:115-140), Q|K|V read from qkv [rows, 3H] f32 as in mmb_attn_fwd.
 */
enum { MMB_ACT_NONE = 0, MMB_ACT_TANH = 1, MMB_ACT_RELU = 2, MMB_ACT_GELU = 3 };
typedef struct mmb_linear_f32_args {
    const float* X;
    const float* W;
    const float* bias; /* [N] or NULL */
    float* Y;
    int64_t ldx, ldw, ldy;
    int32_t M, N, K;
    int32_t act;
} mmb_linear_f32_args;
int mmb_linear_f32(const mmb_linear_f32_args* a, void* stream);

typedef struct mmb_attn_f32_args {
    const float* qkv;          /* [rows, 3H] */
    float* ctx;                /* [rows, H] */
    const float* keybias;      /* [rows] */
    const int32_t* cu_seqlens; /* [nseq + 1] */
    int32_t H, nheads, nseq, max_seqlen, total_rows;
} mmb_attn_f32_args;
int mmb_attn_f32_fwd(const mmb_attn_f32_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Weight maintenance.  mmb_cast_bf16 refreshes the bf16 GEMM operand copies from the fp32 master
 * parameters; mmb_transpose_f32 builds the [D,H] frame-projection weights the embedding kernel streams.
 * mmb_adamw is one fused pass over a flat parameter range with the semantics of the AdamW the reference
 * constructs (train.py:76-92 -> transformers.optimization.AdamW, <= 4.x): m/v update, step size
 * lr * sqrt(1-beta2^t) / (1-beta1^t) when correct_bias, eps added to sqrt(v), then p -= lr * wd * p;
 * optionally writes the refreshed bf16 copy in the same pass.  grad_scale multiplies the gradient first
 * (1/world_size after a SUM all-reduce).
 */
int mmb_cast_bf16(const float* src, void* dst, size_t n, void* stream);
int mmb_transpose_f32(const float* src, float* dst, int R, int C, void* stream);
typedef struct mmb_adamw_args {
    float* p;
    const float* g;
    float* m;
    float* v;
    void* p_bf16; /* bf16 mirror of p or NULL */
    size_t n;     /* elements, multiple of 4 */
    float lr, beta1, beta2, eps, weight_decay, grad_scale;
    int32_t step;         /* 1-based */
    int32_t correct_bias; /* HF default: 1 */
} mmb_adamw_args;
int mmb_adamw(const mmb_adamw_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMBERT_SM100_H */
