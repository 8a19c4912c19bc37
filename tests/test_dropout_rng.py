"""Joint statistics of the attention-probability dropout mask.

The mask for (query q, key k) of one (sequence, head) is ``(qkey[q] * kkey[k]) mod 2^32 >= thresh`` with two odd hashed
keys (csrc/common.cuh) — a rank-1 construction: within a 2 x 2 minor the fourth product is determined by the other three.
The reference draws i.i.d. Bernoulli decisions.  What matters for training is that the KEEP DECISIONS are statistically
indistinguishable from independent draws; these tests check that on > 10^6 disjoint 2 x 2 minors built from the real key
generator (numpy restatement in oracle/dropout_rng.py): marginal rate, pairs sharing a key column / a query row, and all
16 minor patterns within 3.5 sigma of independence.  tests/test_kernels_gpu.py checks that the kernels' masks are bit for
bit the ones this generator describes."""
import numpy as np
import pytest

from oracle import dropout_rng as R


def _z(freq, expect, n):
    return (freq - expect) / np.sqrt(expect * (1 - expect) / n)


@pytest.mark.parametrize("p,seed,stream", [(0.1, 12345, (3 << 8) | 1), (0.1, 2 ** 61 + 7, (11 << 8) | 1), (0.5, 99, 1)])
def test_attention_dropout_minors_are_independent(p, seed, stream):
    q = 1.0 - R.dropout_threshold(p) / 65536.0
    pd = 1.0 - q
    total_rows, S, blocks = 73600, 64, 1300                 # 1300 x 1024 = 1.33 M disjoint minors
    counts = np.zeros(16)
    keep = n = same_col = same_row = npairs = 0
    for blk in range(blocks):
        head, row0 = blk % 12, (blk * 173) % (total_rows - S)
        rows = head * total_rows + row0 + np.arange(S)
        m = R.attn_keep_mask(seed, stream, rows, rows, p)
        keep += int(m.sum())
        n += m.size
        a, b, c, d = m[0::2, 0::2], m[0::2, 1::2], m[1::2, 0::2], m[1::2, 1::2]
        idx = (a.astype(np.int64) << 3) | (b.astype(np.int64) << 2) | (c.astype(np.int64) << 1) | d.astype(np.int64)
        counts += np.bincount(idx.ravel(), minlength=16)
        same_col += int((a & c).sum())                       # (q1, k), (q2, k)
        same_row += int((a & b).sum())                       # (q, k1), (q, k2)
        npairs += a.size
    tot = counts.sum()
    assert tot > 1_000_000
    assert abs(_z(keep / n, q, n)) < 3.5
    assert abs(_z(same_col / npairs, q * q, npairs)) < 3.5
    assert abs(_z(same_row / npairs, q * q, npairs)) < 3.5
    worst = 0.0
    for pat in range(16):
        e = np.prod([q if (pat >> i) & 1 else pd for i in range(4)])
        worst = max(worst, abs(_z(counts[pat] / tot, e, tot)))
    assert worst < 3.5, worst
    # conditional on the other three being kept, the fourth is kept at the marginal rate
    kept3 = counts[0b1110] + counts[0b1111]
    assert abs(_z(counts[0b1111] / kept3, q, kept3)) < 3.5


def test_row_dropout_keys_are_distinct_across_streams_and_rows():
    k1 = R.rng_row_key(7, 1, np.arange(100000))
    k2 = R.rng_row_key(7, 2, np.arange(100000))
    assert len(np.unique(k1)) > 99990                        # 32-bit birthday collisions only
    assert float(np.mean(k1 == k2)) < 1e-4
    bits = np.unpackbits(k1.view(np.uint8)).mean()
    assert abs(bits - 0.5) < 2e-3
