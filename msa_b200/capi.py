"""ctypes binding of libmmbert_sm100.so (C ABI: include/mmbert_sm100.h).

This is the only place the shared library is loaded.  There is NO fallback: if the library is missing
or the device is not sm_100 the import of a compute entry point raises.  torch is used here only for
device pointers (``tensor.data_ptr()``) and the current CUDA stream handle.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmmbert_sm100.so")

MMB_OK, MMB_EINVAL, MMB_EARCH, MMB_ECUDA = 0, -1, -2, -3
MAJOR_K, MAJOR_MN = 0, 1
EPI_STORE_BF16, EPI_GELU_BF16, EPI_RELU_BF16, EPI_STORE_F32, EPI_ATOMIC_ADD_F32, EPI_DGELU_BF16 = range(6)

_lib = None


class MMBError(RuntimeError):
    pass


def lib():
    """Loads the shared library once.  Raises if it has not been built (``python -m msa_b200.build``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MMBError(
                f"{LIB_PATH} not found: build it with `python -m msa_b200.build` "
                "(there is no CPU / PyTorch fallback for the MMBert hot path)")
        L = ctypes.CDLL(LIB_PATH)
        L.mmb_last_error.restype = ctypes.c_char_p
        L.mmb_version.restype = ctypes.c_int
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != MMB_OK:
        msg = lib().mmb_last_error().decode("utf-8", "replace")
        raise MMBError(f"{what}: status {rc}: {msg}")


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def call(name, args_struct):
    """Invokes ``int mmb_<name>(const args*, void* stream)`` on torch's current stream."""
    fn = getattr(lib(), "mmb_" + name)
    check(fn(ctypes.byref(args_struct), stream_ptr()), "mmb_" + name)


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("A", ctypes.c_void_p), ("B", ctypes.c_void_p), ("C", ctypes.c_void_p), ("aux", ctypes.c_void_p),
        ("bias", ctypes.c_void_p),
        ("lda", ctypes.c_int64), ("ldb", ctypes.c_int64), ("ldc", ctypes.c_int64), ("ldaux", ctypes.c_int64),
        ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32),
        ("a_major", ctypes.c_int32), ("b_major", ctypes.c_int32),
        ("epilogue", ctypes.c_int32), ("split_k", ctypes.c_int32),
        ("alpha", ctypes.c_float), ("dbg_flags", ctypes.c_int32),
    ]


def gemm(A, B, C, M, N, K, *, a_major=MAJOR_K, b_major=MAJOR_K, epilogue=EPI_STORE_BF16, bias=None, aux=None,
         split_k=1, alpha=1.0, lda=None, ldb=None, ldc=None, ldaux=None, dbg_flags=0):
    """C[M,N] = epilogue(alpha * A·Bᵀ).  A/B are 2-D bf16 CUDA tensors (stride(1) == 1); see mmb_gemm."""
    a = GemmArgs()
    a.A, a.B, a.C, a.aux, a.bias = A.data_ptr(), B.data_ptr(), C.data_ptr(), \
        (aux.data_ptr() if aux is not None else 0), (bias.data_ptr() if bias is not None else 0)
    a.lda = A.stride(0) if lda is None else lda
    a.ldb = B.stride(0) if ldb is None else ldb
    a.ldc = C.stride(0) if ldc is None else ldc
    a.ldaux = (aux.stride(0) if aux is not None else 0) if ldaux is None else ldaux
    a.M, a.N, a.K = M, N, K
    a.a_major, a.b_major, a.epilogue, a.split_k = a_major, b_major, epilogue, split_k
    a.alpha, a.dbg_flags = alpha, dbg_flags
    call("gemm", a)


class DrlnFwdArgs(ctypes.Structure):
    _fields_ = [("y", ctypes.c_void_p), ("res", ctypes.c_void_p), ("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p),
                ("out", ctypes.c_void_p), ("mean", ctypes.c_void_p), ("rstd", ctypes.c_void_p),
                ("M", ctypes.c_int32), ("H", ctypes.c_int32), ("eps", ctypes.c_float), ("p_drop", ctypes.c_float),
                ("seed", ctypes.c_uint64), ("rng_stream", ctypes.c_uint32)]


class DrlnBwdArgs(ctypes.Structure):
    _fields_ = [("g1", ctypes.c_void_p), ("g2", ctypes.c_void_p), ("y", ctypes.c_void_p), ("res", ctypes.c_void_p),
                ("mean", ctypes.c_void_p), ("rstd", ctypes.c_void_p), ("gamma", ctypes.c_void_p),
                ("d_y", ctypes.c_void_p), ("d_res", ctypes.c_void_p), ("dgamma", ctypes.c_void_p),
                ("dbeta", ctypes.c_void_p), ("dbias", ctypes.c_void_p),
                ("M", ctypes.c_int32), ("H", ctypes.c_int32), ("p_drop", ctypes.c_float),
                ("seed", ctypes.c_uint64), ("rng_stream", ctypes.c_uint32)]


class ColsumArgs(ctypes.Structure):
    _fields_ = [("X", ctypes.c_void_p), ("out", ctypes.c_void_p), ("ld", ctypes.c_int64),
                ("M", ctypes.c_int32), ("N", ctypes.c_int32)]


class AttnArgs(ctypes.Structure):
    _fields_ = [("qkv", ctypes.c_void_p), ("ctx", ctypes.c_void_p), ("lse", ctypes.c_void_p),
                ("keybias", ctypes.c_void_p), ("cu_seqlens", ctypes.c_void_p), ("dctx", ctypes.c_void_p),
                ("dqkv", ctypes.c_void_p), ("dsum", ctypes.c_void_p),
                ("H", ctypes.c_int32), ("nheads", ctypes.c_int32), ("nseq", ctypes.c_int32),
                ("max_seqlen", ctypes.c_int32), ("total_rows", ctypes.c_int32), ("p_drop", ctypes.c_float),
                ("seed", ctypes.c_uint64), ("rng_stream", ctypes.c_uint32)]


def _p(t):
    return 0 if t is None else t.data_ptr()


def drln_fwd(y, res, gamma, beta, out, mean, rstd, eps, p_drop=0.0, seed=0, rng_stream=0):
    a = DrlnFwdArgs(_p(y), _p(res), _p(gamma), _p(beta), _p(out), _p(mean), _p(rstd), y.shape[0], y.shape[1],
                    eps, p_drop, seed, rng_stream)
    call("dropout_residual_ln_fwd", a)


def drln_bwd(g1, g2, y, res, mean, rstd, gamma, d_y, d_res, dgamma, dbeta, dbias, p_drop=0.0, seed=0, rng_stream=0):
    a = DrlnBwdArgs(_p(g1), _p(g2), _p(y), _p(res), _p(mean), _p(rstd), _p(gamma), _p(d_y), _p(d_res), _p(dgamma),
                    _p(dbeta), _p(dbias), y.shape[0], y.shape[1], p_drop, seed, rng_stream)
    call("dropout_residual_ln_bwd", a)


def colsum(X, out):
    a = ColsumArgs(_p(X), _p(out), X.stride(0), X.shape[0], X.shape[1])
    call("colsum_bf16", a)


def attn_args(qkv, ctx, lse, keybias, cu_seqlens, H, nheads, max_seqlen, dctx=None, dqkv=None, dsum=None,
              p_drop=0.0, seed=0, rng_stream=0):
    return AttnArgs(_p(qkv), _p(ctx), _p(lse), _p(keybias), _p(cu_seqlens), _p(dctx), _p(dqkv), _p(dsum),
                    H, nheads, cu_seqlens.numel() - 1, max_seqlen, qkv.shape[0], p_drop, seed, rng_stream)
