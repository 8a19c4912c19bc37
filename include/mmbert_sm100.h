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

#define MMB_VERSION 100 /* round 1 */

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
    MMB_EPI_DGELU_BF16 = 5      /* C(bf16) = acc * gelu_erf'(aux[m,n]); aux is an INPUT (pre-act)    */
};

typedef struct mmb_gemm_args {
    const void* A;
    const void* B;
    void* C;
    void* aux;         /* see epilogue; row stride = ldaux */
    const float* bias; /* [N] f32 or NULL */
    int64_t lda, ldb, ldc, ldaux;
    int32_t M, N, K;
    int32_t a_major, b_major;
    int32_t epilogue;
    int32_t split_k; /* >= 1; > 1 only with MMB_EPI_ATOMIC_ADD_F32 */
    float alpha;
    /* debug overrides for bring-up (0 = use built-in values) */
    int32_t dbg_flags;
} mmb_gemm_args;

int mmb_gemm(const mmb_gemm_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * dropout + residual + LayerNorm (warp-per-row, HBM-bound).
 *   out = LayerNorm_eps( dropout_p(y) + res ) * gamma + beta ;  mean/rstd saved for backward.
 * Replaces BertSelfOutput.forward (modeling_bert.py:295-297), BertOutput.forward (:353-355) and, with
 * res == NULL and p_drop == 0, the LayerNorm of BertPredictionHeadTransform (:484).  The dense bias is
 * added by the producing GEMM's epilogue.  Dropout masks come from a counter-based generator keyed by
 * (seed, rng_stream, row * H + col): backward regenerates them, nothing is stored.
 */
typedef struct mmb_drln_fwd_args {
    const void* y;   /* [M,H] bf16 dense output (bias included) */
    const void* res; /* [M,H] bf16 residual or NULL */
    const float* gamma;
    const float* beta;
    void* out;   /* [M,H] bf16 */
    float* mean; /* [M] */
    float* rstd; /* [M] */
    int32_t M, H;
    float eps;
    float p_drop;
    uint64_t seed;
    uint32_t rng_stream;
} mmb_drln_fwd_args;
int mmb_dropout_residual_ln_fwd(const mmb_drln_fwd_args* a, void* stream);

/* Backward of the above (autograd of modeling_bert.py:295-297 / :353-355).
 *   dz = LayerNormBackward(g1 + g2);  d_res = dz;  d_y = dz * mask / (1-p)
 *   dgamma += sum_rows (g1+g2) * xhat;  dbeta += sum_rows (g1+g2);  dbias += sum_rows d_y   (fp32 atomics)
 */
typedef struct mmb_drln_bwd_args {
    const void* g1; /* [M,H] bf16 gradient of out */
    const void* g2; /* [M,H] bf16 second gradient of out (residual use by the next block) or NULL */
    const void* y;
    const void* res; /* or NULL */
    const float* mean;
    const float* rstd;
    const float* gamma;
    void* d_y;     /* [M,H] bf16 */
    void* d_res;   /* [M,H] bf16 or NULL */
    float* dgamma; /* [H] accumulated */
    float* dbeta;  /* [H] accumulated */
    float* dbias;  /* [H] accumulated, or NULL */
    int32_t M, H;
    float p_drop;
    uint64_t seed;
    uint32_t rng_stream;
} mmb_drln_bwd_args;
int mmb_dropout_residual_ln_bwd(const mmb_drln_bwd_args* a, void* stream);

/* out[n] += sum_m X[m,n] — bias gradients of the dense layers whose dY is produced by a GEMM / attention
 * kernel (autograd of nn.Linear bias, modeling_bert.py:179-181, :340). */
typedef struct mmb_colsum_args {
    const void* X; /* [M,N] bf16 */
    float* out;    /* [N] f32, accumulated */
    int64_t ld;
    int32_t M, N;
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
 * the same layout (autograd of the same lines); dsum [heads, rows] f32 is caller-provided scratch.
 */
typedef struct mmb_attn_args {
    const void* qkv;        /* [rows, 3H] bf16 */
    void* ctx;              /* [rows, H] bf16 (output of fwd, input of bwd) */
    float* lse;             /* [heads, rows] */
    const float* keybias;   /* [rows] */
    const int32_t* cu_seqlens; /* [nseq + 1] row offsets */
    const void* dctx;       /* bwd: [rows, H] bf16 */
    void* dqkv;             /* bwd: [rows, 3H] bf16 */
    float* dsum;            /* bwd scratch: [heads, rows] */
    int32_t H, nheads, nseq, max_seqlen, total_rows;
    float p_drop;
    uint64_t seed;
    uint32_t rng_stream;
} mmb_attn_args;
int mmb_attn_fwd(const mmb_attn_args* a, void* stream);
int mmb_attn_bwd(const mmb_attn_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMBERT_SM100_H */
