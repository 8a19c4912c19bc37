"""ctypes binding of libmmbert_sm100.so (C ABI: include/mmbert_sm100.h).

This is the only place the shared library is loaded.  There is NO fallback: if the library is missing
or the device is not sm_100 the import of a compute entry point raises.  torch is used here only for
device pointers (``tensor.data_ptr()``) and the current CUDA stream handle.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMB_LIB") or os.path.join(_HERE, "lib", "libmmbert_sm100.so")     # MMB_LIB: A/B builds only

MMB_OK, MMB_EINVAL, MMB_EARCH, MMB_ECUDA = 0, -1, -2, -3
MAJOR_K, MAJOR_MN = 0, 1
(EPI_STORE_BF16, EPI_GELU_BF16, EPI_RELU_BF16, EPI_STORE_F32, EPI_ATOMIC_ADD_F32, EPI_DGELU_BF16, EPI_GELU_GRAD_BF16,
 EPI_MUL_AUX_BF16, EPI_CE_STATS) = range(9)

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



# ---------------------------------------------------------------------------------------------------
# The argument structures are generated from include/mmbert_sm100.h, so the Python side cannot drift
# from the C ABI.  Only the header's own conventions are parsed (plain scalars, pointers, fixed arrays).
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mmbert_sm100.h")
_SCALARS = {"int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "uint32_t": ctypes.c_uint32,
            "uint64_t": ctypes.c_uint64, "float": ctypes.c_float, "size_t": ctypes.c_size_t, "int": ctypes.c_int}


def _parse_header(path=HEADER_PATH):
    import re
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    structs = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        fields = []
        for stmt in m.group(2).split(";"):
            stmt = " ".join(stmt.split())
            if not stmt:
                continue
            mm = re.match(r"^((?:const\s+)?\w+)\s*(.*)$", stmt)
            base, decls = mm.group(1).replace("const ", "").strip(), mm.group(2)
            for d in decls.split(","):
                d = d.strip()
                is_ptr = d.startswith("*")
                d = d.lstrip("* ").strip()
                am = re.match(r"^(\w+)\s*\[(\d+)\]$", d)
                name, count = (am.group(1), int(am.group(2))) if am else (d, None)
                ctype = ctypes.c_void_p if is_ptr else _SCALARS[base]
                fields.append((name, ctype * count if count else ctype))
        structs[m.group(3)] = fields
    funcs = re.findall(r"^\s*(?:int|size_t|long long|const char\*)\s+(mmb_\w+)\s*\(", src, flags=re.M)
    return structs, funcs


_STRUCT_FIELDS, DECLARED_FUNCTIONS = _parse_header()


def _make_struct(cname):
    return type(cname, (ctypes.Structure,), {"_fields_": _STRUCT_FIELDS[cname]})


GemmArgs = _make_struct("mmb_gemm_args")
DrlnFwdArgs = _make_struct("mmb_drln_fwd_args")
DrlnBwdArgs = _make_struct("mmb_drln_bwd_args")
ColsumArgs = _make_struct("mmb_colsum_args")
AttnArgs = _make_struct("mmb_attn_args")
AttnScheduleArgs = _make_struct("mmb_attn_schedule_args")
PackArgs = _make_struct("mmb_pack_args")
EmbedArgs = _make_struct("mmb_embed_args")
CeArgs = _make_struct("mmb_ce_args")
CeSparseArgs = _make_struct("mmb_ce_sparse_args")
HeadsArgs = _make_struct("mmb_heads_args")
AdamwArgs = _make_struct("mmb_adamw_args")
MlmMaskArgs = _make_struct("mmb_mlm_mask_args")
LinearF32Args = _make_struct("mmb_linear_f32_args")
AttnF32Args = _make_struct("mmb_attn_f32_args")
ACT_NONE, ACT_TANH, ACT_RELU, ACT_GELU = range(4)

DT_F32, DT_F64, DT_I64, DT_I32, DT_U8 = range(5)
_DT = {torch.float32: DT_F32, torch.float64: DT_F64, torch.int64: DT_I64, torch.int32: DT_I32, torch.uint8: DT_U8,
       torch.bool: DT_U8}


def dtype_code(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise MMBError(f"unsupported input dtype {t.dtype}")


def _p(t):
    return 0 if t is None else t.data_ptr()


def fill(struct, **kw):
    """Sets fields of a ctypes struct; tensors become device pointers, lists fill fixed arrays."""
    for k, v in kw.items():
        if isinstance(v, (list, tuple)):
            arr = getattr(struct, k)
            for i, x in enumerate(v):
                arr[i] = _p(x) if (torch.is_tensor(x) or x is None) else x
        elif torch.is_tensor(v) or v is None:
            setattr(struct, k, _p(v))
        else:
            setattr(struct, k, v)
    return struct


def gemm_args(A, B, C, M, N, K, *, a_major=MAJOR_K, b_major=MAJOR_K, epilogue=EPI_STORE_BF16, bias=None, aux=None,
              aux2=None, split_k=1, alpha=1.0, lda=None, ldb=None, ldc=None, ldaux=None, dbg_flags=0, colsum=None,
              row_live=None, dead_rows_zeroed=0):
    a = GemmArgs()
    return fill(a, A=A, B=B, C=C, aux=aux, aux2=aux2, bias=bias, colsum=colsum, row_live=row_live,
                dead_rows_zeroed=dead_rows_zeroed,
                lda=A.stride(0) if lda is None else lda, ldb=B.stride(0) if ldb is None else ldb,
                ldc=(C.stride(0) if C is not None else 0) if ldc is None else ldc,
                ldaux=((aux.stride(0) if aux is not None else 0) if ldaux is None else ldaux),
                M=M, N=N, K=K, a_major=a_major, b_major=b_major, epilogue=epilogue, split_k=split_k, alpha=alpha,
                dbg_flags=dbg_flags)


def gemm(A, B, C, M, N, K, **kw):
    """C[M,N] = epilogue(alpha * A·Bᵀ).  A/B are 2-D bf16 CUDA tensors (stride(1) == 1); see mmb_gemm."""
    call("gemm", gemm_args(A, B, C, M, N, K, **kw))


def drln_fwd_args(y, res, gamma, beta, out, mean, rstd, eps, p_drop=0.0, seed=0, rng_stream=0, out_f32=None, row_list=None):
    return fill(DrlnFwdArgs(), y=y, res=res, gamma=gamma, beta=beta, out=out, out_f32=out_f32, mean=mean, rstd=rstd,
                M=y.shape[0],
                H=y.shape[1], eps=eps, p_drop=p_drop, seed=seed, rng_stream=rng_stream, row_list=row_list)


def drln_fwd(*a, **kw):
    call("dropout_residual_ln_fwd", drln_fwd_args(*a, **kw))


def drln_bwd_args(g1, g2, y, res, mean, rstd, gamma, d_y, d_res, dgamma, dbeta, dbias, p_drop=0.0, seed=0,
                  rng_stream=0, row_list=None, dead_rows_zeroed=0):
    return fill(DrlnBwdArgs(), g1=g1, g2=g2, y=y, res=res, mean=mean, rstd=rstd, gamma=gamma, d_y=d_y, d_res=d_res,
                dgamma=dgamma, dbeta=dbeta, dbias=dbias, M=y.shape[0], H=y.shape[1], p_drop=p_drop, seed=seed,
                rng_stream=rng_stream, row_list=row_list, dead_rows_zeroed=dead_rows_zeroed)


def drln_bwd(*a, **kw):
    call("dropout_residual_ln_bwd", drln_bwd_args(*a, **kw))


def colsum_args(X, out, row_list=None):
    return fill(ColsumArgs(), X=X, out=out, ld=X.stride(0), M=X.shape[0], N=X.shape[1], row_list=row_list)


def colsum(X, out):
    call("colsum_bf16", colsum_args(X, out))


def attn_bwd_workspace(total_rows, nheads, device):
    """Scratch tensor for mmb_attn_bwd (size from mmb_attn_bwd_workspace_bytes)."""
    L = lib()
    L.mmb_attn_bwd_workspace_bytes.restype = ctypes.c_size_t
    L.mmb_attn_bwd_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int]
    return torch.empty(L.mmb_attn_bwd_workspace_bytes(total_rows, nheads) // 4, device=device, dtype=torch.float32)


def attn_args(qkv, ctx, lse, keybias, cu_seqlens, H, nheads, max_seqlen, dctx=None, dqkv=None, bwd_ws=None,
              p_drop=0.0, seed=0, rng_stream=0, flags=0, kv_end=None, work=None, row_list=None):
    return fill(AttnArgs(), qkv=qkv, ctx=ctx, lse=lse, keybias=keybias, cu_seqlens=cu_seqlens, dctx=dctx, dqkv=dqkv,
                bwd_ws=bwd_ws, kv_end=kv_end, H=H, nheads=nheads, nseq=cu_seqlens.numel() - 1, max_seqlen=max_seqlen,
                total_rows=qkv.shape[0], p_drop=p_drop, seed=seed, rng_stream=rng_stream, flags=flags, work=work,
                row_list=row_list)


def attn_schedule_buffer(nseq, nheads, max_seqlen, device):
    """int32 tensor [records, 4] for mmb_attn_schedule (size from mmb_attn_schedule_bytes)."""
    L = lib()
    L.mmb_attn_schedule_bytes.restype = ctypes.c_size_t
    L.mmb_attn_schedule_bytes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return torch.empty(L.mmb_attn_schedule_bytes(nseq, nheads, max_seqlen) // 16, 4, device=device, dtype=torch.int32)


def row_list_buffer(rows, device):
    """Zeroed int32 tensor for mmb_attn_schedule_args.row_list (size from mmb_row_list_ints)."""
    L = lib()
    L.mmb_row_list_ints.restype = ctypes.c_size_t
    L.mmb_row_list_ints.argtypes = [ctypes.c_int]
    return torch.zeros(L.mmb_row_list_ints(rows), device=device, dtype=torch.int32)


def attn_schedule_args(cu_seqlens, kv_end, work, nheads, max_seqlen, row_label=None, row_list=None):
    return fill(AttnScheduleArgs(), cu_seqlens=cu_seqlens, kv_end=kv_end, work=work, nseq=cu_seqlens.numel() - 1,
                nheads=nheads, max_seqlen=max_seqlen, row_label=row_label, row_list=row_list)


def cast_bf16(src, dst):
    L = lib()
    L.mmb_cast_bf16.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    check(L.mmb_cast_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), stream_ptr()), "mmb_cast_bf16")


def transpose_f32(src, dst, R, C):
    L = lib()
    L.mmb_transpose_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    check(L.mmb_transpose_f32(src.data_ptr(), dst.data_ptr(), R, C, stream_ptr()), "mmb_transpose_f32")


def launch_count():
    L = lib()
    L.mmb_launch_count.restype = ctypes.c_longlong
    return L.mmb_launch_count()


def ce_stats_floats(M, N):
    L = lib()
    L.mmb_ce_stats_floats.restype = ctypes.c_size_t
    L.mmb_ce_stats_floats.argtypes = [ctypes.c_int, ctypes.c_int]
    return L.mmb_ce_stats_floats(M, N)


def heads_workspace_bytes(B, H):
    L = lib()
    L.mmb_heads_workspace_bytes.restype = ctypes.c_size_t
    L.mmb_heads_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int]
    return L.mmb_heads_workspace_bytes(B, H)
