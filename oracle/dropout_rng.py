"""TEST INFRASTRUCTURE — numpy restatement of the library's counter-based dropout generator (msa_b200/csrc/common.cuh:
``rng_row_key`` / ``fmix32`` / ``attn_drop_qkey`` / ``attn_drop_kkey`` / ``attn_keep`` and ``dropout_threshold``).

The reference draws attention-probability dropout with torch's Philox stream (modeling_bert.py:131 ``nn.Dropout`` on the
softmax output); the CUDA path regenerates its masks from (seed, stream, row, column) instead of storing them, so the two
can only be compared statistically.  This module exists so that tests can (a) check the kernel's mask bit for bit against
the documented generator and (b) test the generator's joint statistics on the CPU with millions of samples.
Only ``tests/`` may import it.
"""
import numpy as np

U32 = np.uint32
_M = 0xFFFFFFFF


def _rotl(x, r):
    x = x.astype(np.uint64)
    return (((x << np.uint64(r)) | (x >> np.uint64(32 - r))) & np.uint64(_M)).astype(U32)


def _mul(a, b):
    return ((a.astype(np.uint64) * np.uint64(b)) & np.uint64(_M)).astype(U32)


def fmix32(h):
    h = h ^ (h >> U32(16))
    h = _mul(h, 0x85ebca6b)
    h = h ^ (h >> U32(13))
    h = _mul(h, 0xc2b2ae35)
    return h ^ (h >> U32(16))


def rng_row_key(seed, stream, row_id):
    """common.cuh:133-142.  ``row_id`` may be an array."""
    row_id = np.asarray(row_id, dtype=U32)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    h = np.full(row_id.shape, ((seed & _M) ^ ((int(stream) * 0x9E3779B9) & _M)) & _M, dtype=U32)
    k = _mul(row_id, 0xcc9e2d51)
    k = _mul(_rotl(k, 15), 0x1b873593)
    h = (_mul(_rotl(h ^ k, 13), 5).astype(np.uint64) + np.uint64(0xe6546b64)).astype(np.uint64) & np.uint64(_M)
    h = h.astype(U32)
    k = np.full(row_id.shape, ((seed >> 32) * 0xcc9e2d51) & _M, dtype=U32)
    k = _mul(_rotl(k, 15), 0x1b873593)
    h = ((_mul(_rotl(h ^ k, 13), 5).astype(np.uint64) + np.uint64(0xe6546b64)) & np.uint64(_M)).astype(U32)
    return fmix32(h ^ U32(12))


def attn_drop_qkey(seed, stream, prob_row):
    return rng_row_key(seed, stream, prob_row) | U32(1)


def attn_drop_kkey(seed, stream, prob_row):
    return rng_row_key(int(seed) ^ 0x9E3779B97F4A7C15, (int(stream) ^ 0x5bd1e995) & _M, prob_row) | U32(1)


def dropout_threshold(p):
    """common.cuh:206-211: 16-bit threshold round(p * 65536), clamped."""
    if p <= 0:
        return 0
    t = p * 65536.0 + 0.5
    return 65535 if t >= 65535.0 else int(t)


def attn_keep_mask(seed, stream, q_rows, k_rows, p):
    """Boolean [len(q_rows), len(k_rows)] keep mask of the attention-probability dropout: ``prob_row`` ids are
    head * total_rows + packed_row (csrc/attn.cu:57, attn_fwd_ws.cu:362)."""
    qk = attn_drop_qkey(seed, stream, q_rows).astype(np.uint64)
    kk = attn_drop_kkey(seed, stream, k_rows).astype(np.uint64)
    prod = (qk[:, None] * kk[None, :]) & np.uint64(_M)
    return prod >= np.uint64(dropout_threshold(p) << 16)
