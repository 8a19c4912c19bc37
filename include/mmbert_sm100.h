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

#ifdef __cplusplus
}
#endif
#endif /* MMBERT_SM100_H */
