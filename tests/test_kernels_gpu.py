"""Per-kernel parity on the B200: every CUDA kernel is called through the C ABI (msa_b200.capi) and compared
with a plain fp32 torch restatement of the same op on the same bf16-rounded inputs.  Tolerances: outputs are
bf16 (8 bits of mantissa -> 2^-8 relative rounding), accumulations are fp32."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

BF16_EPS = 2 ** -8


def _bf(x):
    return x.to(torch.bfloat16)


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-20))


@pytest.mark.parametrize("M,N,K,am,bm", [(300, 200, 136, 0, 0), (1000, 768, 768, 0, 0), (768, 256, 1000, 1, 1),
                                          (512, 768, 3072, 0, 1), (129, 30522, 128, 0, 0)])
def test_gemm_f32_exact(M, N, K, am, bm):
    from msa_b200 import capi
    torch.manual_seed(1)
    A = _bf(torch.randn(M, K, device="cuda"))
    B = _bf(torch.randn(N, K, device="cuda"))
    ref = A.float() @ B.float().t()
    pad = lambda t: torch.nn.functional.pad(t, (0, (-t.shape[1]) % 8))[:, :t.shape[1]]
    A_in = pad(A.t().contiguous()) if am else A
    B_in = pad(B.t().contiguous()) if bm else B
    C = torch.zeros(M, (N + 7) // 8 * 8, device="cuda")[:, :N]
    capi.gemm(A_in, B_in, C, M, N, K, a_major=am, b_major=bm, epilogue=capi.EPI_STORE_F32)
    assert _rel(C, ref) < 1e-5


def test_gemm_epilogues():
    from msa_b200 import capi
    torch.manual_seed(2)
    M, N, K = 520, 3072, 768
    A, B = _bf(torch.randn(M, K, device="cuda") * 0.3), _bf(torch.randn(N, K, device="cuda") * 0.3)
    bias = torch.randn(N, device="cuda")
    pre = A.float() @ B.float().t() + bias
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    aux = torch.empty_like(C)
    capi.gemm(A, B, C, M, N, K, epilogue=capi.EPI_GELU_BF16, bias=bias, aux=aux)
    assert _rel(aux.float(), pre) < 2 * BF16_EPS
    assert _rel(C.float(), torch.nn.functional.gelu(aux.float())) < 2 * BF16_EPS
    capi.gemm(A, B, C, M, N, K, epilogue=capi.EPI_RELU_BF16, bias=bias)
    assert _rel(C.float(), torch.relu(pre)) < 2 * BF16_EPS
    u = aux.float().requires_grad_(True)
    torch.nn.functional.gelu(u).sum().backward()
    capi.gemm(A, B, C, M, N, K, epilogue=capi.EPI_DGELU_BF16, aux=aux)
    assert _rel(C.float(), (A.float() @ B.float().t()) * u.grad) < 2 * BF16_EPS
    # forward that saves gelu'(pre) instead of pre, and the plain-multiply backward that consumes it
    dg = torch.empty_like(C)
    capi.gemm(A, B, C, M, N, K, epilogue=capi.EPI_GELU_GRAD_BF16, bias=bias, aux=dg)
    p32 = pre.clone().requires_grad_(True)
    torch.nn.functional.gelu(p32).sum().backward()
    assert _rel(C.float(), torch.nn.functional.gelu(pre)) < 2 * BF16_EPS
    assert _rel(dg.float(), p32.grad) < 2 * BF16_EPS
    capi.gemm(A, B, C, M, N, K, epilogue=capi.EPI_MUL_AUX_BF16, aux=dg)
    assert _rel(C.float(), (A.float() @ B.float().t()) * dg.float()) < 2 * BF16_EPS
    acc = torch.ones(M, N, device="cuda")
    capi.gemm(A, B, acc, M, N, K, epilogue=capi.EPI_ATOMIC_ADD_F32, split_k=4, alpha=0.5)
    assert _rel(acc, 1 + 0.5 * (A.float() @ B.float().t())) < 1e-5


@pytest.mark.parametrize("M,N,K", [(520, 3072, 768), (1000, 1000, 136), (100, 512, 200), (300, 120, 264)])
def test_gemm_fused_column_sums(M, N, K):
    """mmb_gemm_args.colsum: the bias gradient (column sums of the bf16-ROUNDED output, accumulated) taken in the epilogue
    of the CTA-pair kernel — ragged edges, the multiply and plain epilogues — and by the follow-up launch on the shapes
    the single-CTA kernels serve.  Must equal the column sums of the C the same call wrote."""
    from msa_b200 import capi
    torch.manual_seed(5)
    A, B = _bf(torch.randn(M, K, device="cuda") * 0.3), _bf(torch.randn(N, K, device="cuda") * 0.3)
    bias = torch.randn(N, device="cuda")
    aux = _bf(torch.randn(M, N, device="cuda"))
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for epi, kw in ((capi.EPI_MUL_AUX_BF16, dict(aux=aux)), (capi.EPI_STORE_BF16, dict(bias=bias))):
        cs = torch.full((N,), 3.0, device="cuda")
        capi.gemm(A, B, C, M, N, K, epilogue=epi, colsum=cs, **kw)
        ref = (A.float() @ B.float().t())
        ref = ref * aux.float() if epi == capi.EPI_MUL_AUX_BF16 else ref + bias
        assert _rel(C.float(), ref) < 2 * BF16_EPS
        want = 3.0 + C.float().sum(0)
        assert _rel(cs, want) < 1e-5, epi


@pytest.mark.parametrize("H", [128, 768, 1024])
@pytest.mark.parametrize("with_res", [True, False])
def test_dropout_residual_ln_p0(H, with_res):
    from msa_b200 import capi
    torch.manual_seed(3)
    M = 1037
    y, res = _bf(torch.randn(M, H, device="cuda")), torch.randn(M, H, device="cuda") if with_res else None
    gamma, beta = torch.randn(H, device="cuda"), torch.randn(H, device="cuda")
    out = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
    mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
    out32 = torch.empty(M, H, device="cuda")
    capi.drln_fwd(y, res, gamma, beta, out, mean, rstd, 1e-12, out_f32=out32)
    z = (y.float() + (res if with_res else 0)).requires_grad_(True)
    g_ = gamma.clone().requires_grad_(True)
    b_ = beta.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(z, (H,), g_, b_, 1e-12)
    assert _rel(out.float(), ref) < 2 * BF16_EPS
    assert _rel(out32, ref) < 1e-5           # fp32 residual-stream copy
    assert _rel(mean, z.mean(-1)) < 1e-5
    g1, g2 = _bf(torch.randn(M, H, device="cuda")), torch.randn(M, H, device="cuda")
    ref.backward(g1.float() + g2)
    d_y, d_res = torch.empty_like(out), torch.empty(M, H, device="cuda")
    dgamma, dbeta, dbias = (torch.zeros(H, device="cuda") for _ in range(3))
    capi.drln_bwd(g1, g2, y, res, mean, rstd, gamma, d_y, d_res, dgamma, dbeta, dbias)
    assert _rel(d_y.float(), z.grad) < 2 * BF16_EPS
    assert _rel(d_res, z.grad) < 1e-5
    assert _rel(dgamma, g_.grad) < 1e-4
    assert _rel(dbeta, b_.grad) < 1e-4
    assert _rel(dbias, d_y.float().sum(0)) < 1e-4


def test_dropout_statistics_and_consistency():
    """p > 0: keep-rate, 1/(1-p) scaling and forward/backward mask agreement (RNG streams cannot match torch)."""
    from msa_b200 import capi
    torch.manual_seed(4)
    M, H, p = 2048, 768, 0.1
    y = _bf(torch.ones(M, H, device="cuda"))
    gamma, beta = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
    out = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
    mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
    capi.drln_fwd(y, None, gamma, beta, out, mean, rstd, 1e-5, p_drop=p, seed=77, rng_stream=5)
    # z = mask/(1-p): row mean = keep fraction / (1-p) -> 1 on average
    assert abs(float(mean.mean()) - 1.0) < 5e-3
    # normalised output is negative exactly where the element was dropped
    dropped = out.float() < 0
    rate = float(dropped.float().mean())
    assert abs(rate - p) < 3e-3
    g1 = _bf(torch.ones(M, H, device="cuda"))
    d_y, d_res = torch.empty_like(out), torch.empty(M, H, device="cuda")
    dgamma, dbeta = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    capi.drln_bwd(g1, None, y, None, mean, rstd, gamma, d_y, d_res, dgamma, dbeta, None, p_drop=p, seed=77, rng_stream=5)
    # the backward mask zeroes d_y exactly at the dropped positions
    assert bool(((d_y.float() == 0) | ~dropped).all())
    assert bool((d_y.float()[dropped] == 0).all())
    # a different stream gives a different mask
    out2 = torch.empty_like(out)
    capi.drln_fwd(y, None, gamma, beta, out2, mean, rstd, 1e-5, p_drop=p, seed=77, rng_stream=6)
    assert float(((out2.float() < 0) != dropped).float().mean()) > 0.1


def test_colsum():
    from msa_b200 import capi
    torch.manual_seed(5)
    X = _bf(torch.randn(5000, 2304, device="cuda"))
    out = torch.ones(2304, device="cuda")
    capi.colsum(X, out)
    assert _rel(out, 1 + X.float().sum(0)) < 1e-4


def _attn_ref(qkv, keybias, cu, H, nh):
    d = H // nh
    outs = []
    for i in range(len(cu) - 1):
        a, b = cu[i], cu[i + 1]
        q, k, v = (qkv[a:b, j * H:(j + 1) * H].view(b - a, nh, d).transpose(0, 1) for j in range(3))
        s = q @ k.transpose(1, 2) / math.sqrt(d) + keybias[a:b][None, None, :]
        outs.append((torch.softmax(s, -1) @ v).transpose(0, 1).reshape(b - a, H))
    return torch.cat(outs)


@pytest.mark.parametrize("lens,nh", [([50, 100, 100, 7], 2), ([64, 128, 65], 12), ([550, 3, 201], 4), ([129, 128, 1, 300], 3),
                                     ([2048, 1], 1)])
def test_attention_fwd_bwd(lens, nh):
    from msa_b200 import capi
    torch.manual_seed(6)
    H = nh * 64
    rows = sum(lens)
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    qkv = _bf(torch.randn(rows, 3 * H, device="cuda"))
    keybias = torch.where(torch.rand(rows, device="cuda") < 0.25, -10000.0, 0.0)
    for i in range(len(lens)):
        keybias[cu[i]] = 0.0   # [CLS] is always a valid key
    cu_t = torch.tensor(cu, device="cuda", dtype=torch.int32)
    ctx = torch.empty(rows, H, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(nh, rows, device="cuda")
    capi.call("attn_fwd", capi.attn_args(qkv, ctx, lse, keybias, cu_t, H, nh, max(lens)))
    x = qkv.float().requires_grad_(True)
    ref = _attn_ref(x, keybias, cu, H, nh)
    assert _rel(ctx.float(), ref) < 3 * BF16_EPS
    # the first tcgen05 forward (csrc/attn_tc.cu, kept behind flags bit 2 / MMB_ATTN_FWD=tc for A/B runs): same contract
    ctx_tc, lse_tc = torch.empty_like(ctx), torch.empty_like(lse)
    capi.call("attn_fwd", capi.attn_args(qkv, ctx_tc, lse_tc, keybias, cu_t, H, nh, max(lens), flags=4))
    assert _rel(ctx_tc.float(), ref) < 3 * BF16_EPS and _rel(lse_tc, lse) < 1e-3
    dctx = _bf(torch.randn(rows, H, device="cuda"))
    ref.backward(dctx.float())
    dqkv = torch.zeros(rows, 3 * H, device="cuda", dtype=torch.bfloat16)
    bwd_ws = capi.attn_bwd_workspace(rows, nh, "cuda")
    capi.call("attn_bwd", capi.attn_args(qkv, ctx, lse, keybias, cu_t, H, nh, max(lens), dctx=dctx, dqkv=dqkv, bwd_ws=bwd_ws))
    for j, name in enumerate("QKV"):
        got, want = dqkv[:, j * H:(j + 1) * H].float(), x.grad[:, j * H:(j + 1) * H]
        assert _rel(got, want) < 2e-2, name   # P and dS are rounded to bf16 before the second MMA


def test_attention_many_short_sequences():
    """More sequences than the backward kernel keeps in shared memory (2048): exercises its global-metadata path, and
    work items of one to two rows."""
    from msa_b200 import capi
    torch.manual_seed(12)
    nh, H = 1, 64
    lens = [3, 5, 4, 1, 2, 6] * 360            # 2160 sequences
    rows = sum(lens)
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    qkv = _bf(torch.randn(rows, 3 * H, device="cuda"))
    keybias = torch.zeros(rows, device="cuda")
    cu_t = torch.tensor(cu, device="cuda", dtype=torch.int32)
    ctx = torch.empty(rows, H, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(nh, rows, device="cuda")
    dctx = _bf(torch.randn(rows, H, device="cuda"))
    dqkv = torch.zeros(rows, 3 * H, device="cuda", dtype=torch.bfloat16)
    a = capi.attn_args(qkv, ctx, lse, keybias, cu_t, H, nh, max(lens), dctx=dctx, dqkv=dqkv,
                       bwd_ws=capi.attn_bwd_workspace(rows, nh, "cuda"))
    capi.call("attn_fwd", a)
    capi.call("attn_bwd", a)
    # block-diagonal reference in one shot: additive -inf mask between different sequences
    seq_id = torch.repeat_interleave(torch.arange(len(lens), device="cuda"), torch.tensor(lens, device="cuda"))
    x = qkv.float().requires_grad_(True)
    q, k, v = x[:, :H], x[:, H:2 * H], x[:, 2 * H:]
    s_ = q @ k.t() / 8.0
    s_ = s_.masked_fill(seq_id[:, None] != seq_id[None, :], float("-inf"))
    ref = torch.softmax(s_, -1) @ v
    assert _rel(ctx.float(), ref) < 3 * BF16_EPS
    ref.backward(dctx.float())
    assert _rel(dqkv.float(), x.grad) < 2e-2


def test_attention_all_keys_masked_matches_additive_mask():
    """A sequence whose keys are ALL masked: the reference's additive -10000 (not -inf) makes softmax uniform
    over the real keys; the kernel must reproduce that, not NaN."""
    from msa_b200 import capi
    torch.manual_seed(7)
    nh, H, S = 2, 128, 40
    qkv = _bf(torch.randn(S, 3 * H, device="cuda") * 0.1)
    keybias = torch.full((S,), -10000.0, device="cuda")
    cu_t = torch.tensor([0, S], device="cuda", dtype=torch.int32)
    ctx = torch.empty(S, H, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(nh, S, device="cuda")
    capi.call("attn_fwd", capi.attn_args(qkv, ctx, lse, keybias, cu_t, H, nh, S))
    ref = _attn_ref(qkv.float(), keybias, [0, S], H, nh)
    assert torch.isfinite(ctx.float()).all()
    assert _rel(ctx.float(), ref) < 3 * BF16_EPS


def test_attention_dropout_consistency():
    """With p > 0 the context must equal (P ∘ mask / (1-p)) V for SOME mask with keep-rate 1-p, and the backward
    must use the same mask: checked through linearity in V (ctx is linear in V for a fixed mask)."""
    from msa_b200 import capi
    torch.manual_seed(8)
    nh, H, S, p = 2, 128, 96, 0.5
    qkv = _bf(torch.randn(S, 3 * H, device="cuda") * 0.5)
    keybias = torch.zeros(S, device="cuda")
    cu_t = torch.tensor([0, S], device="cuda", dtype=torch.int32)

    def run(x):
        ctx = torch.empty(S, H, device="cuda", dtype=torch.bfloat16)
        lse = torch.empty(nh, S, device="cuda")
        capi.call("attn_fwd", capi.attn_args(x, ctx, lse, keybias, cu_t, H, nh, S, p_drop=p, seed=9, rng_stream=3))
        return ctx.float(), lse

    # V = one-hot columns recovers the dropped probability matrix column sums: use V = 1 -> ctx = rowsum(P_drop)
    x1 = qkv.clone()
    x1[:, 2 * H:] = 1.0
    c1, _ = run(x1)
    # E[rowsum(P ∘ mask/(1-p))] = 1; with S = 96 keys the spread is wide but the global mean is tight
    assert abs(float(c1.mean()) - 1.0) < 0.05
    assert float(c1.std()) > 0.01     # some dropping actually happened
    c2, _ = run(x1)
    assert torch.equal(c1, c2)        # deterministic in (seed, stream)


def test_attention_dropout_forward_backward_use_the_same_mask():
    """V = identity makes ctx the dropped probability matrix P∘mask/(1-p) itself; the backward must then give
    dV = (P∘mask/(1-p))^T dO with exactly that matrix."""
    from msa_b200 import capi
    torch.manual_seed(11)
    nh, H, S, p = 1, 64, 64, 0.3
    qkv = _bf(torch.randn(S, 3 * H, device="cuda") * 0.5)
    qkv[:, 2 * H:] = torch.eye(S, device="cuda").to(torch.bfloat16)
    keybias = torch.zeros(S, device="cuda")
    cu_t = torch.tensor([0, S], device="cuda", dtype=torch.int32)
    ctx = torch.empty(S, H, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(nh, S, device="cuda")
    capi.call("attn_fwd", capi.attn_args(qkv, ctx, lse, keybias, cu_t, H, nh, S, p_drop=p, seed=5, rng_stream=2))
    Pd = ctx.float()                                   # [q, k]
    drop_rate = float((Pd == 0).float().mean())
    assert abs(drop_rate - p) < 0.03
    dctx = _bf(torch.randn(S, H, device="cuda"))
    dqkv = torch.zeros(S, 3 * H, device="cuda", dtype=torch.bfloat16)
    bwd_ws = capi.attn_bwd_workspace(S, nh, "cuda")
    capi.call("attn_bwd", capi.attn_args(qkv, ctx, lse, keybias, cu_t, H, nh, S, dctx=dctx, dqkv=dqkv, bwd_ws=bwd_ws,
                                         p_drop=p, seed=5, rng_stream=2))
    dV = dqkv[:, 2 * H:].float()
    assert _rel(dV, Pd.t() @ dctx.float()) < 2e-2


@pytest.mark.parametrize("p", [0.1, 0.5])
def test_attention_dropout_mask_is_the_documented_generator(p):
    """V = identity makes ctx the dropped probability matrix itself, so its zero pattern IS the kernel's mask: it must be
    bit for bit the mask of the counter-based generator restated in oracle/dropout_rng.py (whose joint statistics —
    pairs, 2 x 2 minors — are tested on > 10^6 samples in tests/test_dropout_rng.py), for every (sequence, head) of a
    packed batch, and the backward must use the same mask (dV = (P o mask / (1-p))^T dO)."""
    import numpy as np
    from msa_b200 import capi
    from oracle import dropout_rng as R
    torch.manual_seed(31)
    nh, S, nseq, seed, stream = 4, 64, 5, 2 ** 40 + 17, (7 << 8) | 1
    H, rows = nh * 64, nseq * S
    qkv = _bf(torch.randn(rows, 3 * H, device="cuda") * 0.3)
    eye = torch.eye(S, device="cuda").to(torch.bfloat16)
    for h in range(nh):
        qkv[:, 2 * H + h * 64:2 * H + (h + 1) * 64] = eye.repeat(nseq, 1)
    keybias = torch.zeros(rows, device="cuda")
    cu_t = torch.arange(0, rows + 1, S, device="cuda", dtype=torch.int32)
    ctx = torch.empty(rows, H, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(nh, rows, device="cuda")
    dctx = _bf(torch.randn(rows, H, device="cuda"))
    dqkv = torch.zeros(rows, 3 * H, device="cuda", dtype=torch.bfloat16)
    a = capi.attn_args(qkv, ctx, lse, keybias, cu_t, H, nh, S, p_drop=p, seed=seed, rng_stream=stream, dctx=dctx, dqkv=dqkv,
                       bwd_ws=capi.attn_bwd_workspace(rows, nh, "cuda"))
    capi.call("attn_fwd", a)
    capi.call("attn_bwd", a)
    Pd = ctx.float().view(nseq, S, nh, 64)                # [seq, q, head, k]
    for i in range(nseq):
        for h in range(nh):
            ids = h * rows + i * S + np.arange(S)
            want = torch.from_numpy(R.attn_keep_mask(seed, stream, ids, ids, p)).cuda()
            got = Pd[i, :, h, :] != 0
            assert torch.equal(got, want), (i, h, int((got != want).sum()))
            dV = dqkv[i * S:(i + 1) * S, 2 * H + h * 64:2 * H + (h + 1) * 64].float()
            assert _rel(dV, Pd[i, :, h, :].t() @ dctx[i * S:(i + 1) * S, h * 64:(h + 1) * 64].float()) < 2e-2


def test_attention_masked_tail_skipping_is_exact():
    """kv_end lets the kernels skip whole key tiles behind the last unmasked key; results must be identical to the
    full computation (forward context / LSE and all three gradients)."""
    from msa_b200 import capi
    torch.manual_seed(21)
    nh, lens = 2, [550, 300, 40]
    valid = [131, 300, 0]          # seq 0: masked tail from key 131; seq 1: nothing masked; seq 2: everything masked
    H, rows = nh * 64, sum(lens)
    cu = [0, 550, 850, 890]
    qkv = _bf(torch.randn(rows, 3 * H, device="cuda"))
    keybias = torch.zeros(rows, device="cuda")
    keybias[131:550] = -10000.0
    keybias[850:890] = -10000.0
    kv_end = torch.tensor(valid, device="cuda", dtype=torch.int32)
    cu_t = torch.tensor(cu, device="cuda", dtype=torch.int32)
    dctx = _bf(torch.randn(rows, H, device="cuda"))
    res = []
    for kv in (None, kv_end):
        ctx = torch.zeros(rows, H, device="cuda", dtype=torch.bfloat16)
        lse = torch.zeros(nh, rows, device="cuda")
        dqkv = torch.full((rows, 3 * H), 7.0, device="cuda", dtype=torch.bfloat16)
        bwd_ws = capi.attn_bwd_workspace(rows, nh, "cuda")
        a = capi.attn_args(qkv, ctx, lse, keybias, cu_t, H, nh, max(lens), dctx=dctx, dqkv=dqkv, bwd_ws=bwd_ws, kv_end=kv,
                           p_drop=0.1, seed=4, rng_stream=1)
        capi.call("attn_fwd", a)
        capi.call("attn_bwd", a)
        res.append((ctx.float(), lse.clone(), dqkv.float()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert torch.equal(res[0][2], res[1][2])
    assert float(res[1][2][131:550, H:].abs().max()) == 0.0        # dK, dV of masked keys


def _expected_work_lists(lens, kv_end, nh):
    """Restates the record layout documented at mmb_attn_schedule (include/mmbert_sm100.h)."""
    n = len(lens)
    cu = [0]
    for s in lens:
        cu.append(cu[-1] + s)
    eff = [(e if 0 < e < s else s) for s, e in zip(lens, kv_end)]
    tq = [(s + 127) // 128 for s in lens]
    tkv = [(e + 127) // 128 for e in eff]
    rec = lambda i, h, t: [cu[i], lens[i], eff[i], (h << 16) | t]
    q = [rec(i, h, t) for i in sorted(range(n), key=lambda i: (-eff[i], i)) for h in range(nh) for t in range(tq[i])]
    kv = [rec(i, h, t) for i in sorted(range(n), key=lambda i: (-lens[i], i)) for h in range(nh) for t in range(tkv[i])]
    z = [rec(i, h, t) for i in range(n) for h in range(nh) for t in range(tkv[i], tq[i])]
    return q, kv, z


@pytest.mark.parametrize("with_kv_end", [False, True])
def test_attention_schedule_lists(with_kv_end):
    from msa_b200 import capi
    g = torch.Generator().manual_seed(3)
    lens = torch.randint(1, 551, (70,), generator=g).tolist() + [550, 550, 128, 129, 1]
    kv = [int(torch.randint(0, s + 40, (1,), generator=g)) for s in lens] if with_kv_end else [0] * len(lens)
    nh, max_s = 3, max(lens)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), device="cuda", dtype=torch.int32)
    kv_t = torch.tensor(kv, device="cuda", dtype=torch.int32) if with_kv_end else None
    work = capi.attn_schedule_buffer(len(lens), nh, max_s, "cuda")
    work.fill_(-1)
    capi.call("attn_schedule", capi.attn_schedule_args(cu, kv_t, work, nh, max_s))
    w = work.cpu().tolist()
    q, kvl, z = _expected_work_lists(lens, kv, nh)
    cap = len(lens) * nh * ((max_s + 127) // 128)
    assert w[0] == [len(q), len(kvl) + len(z), cap, len(kvl)]
    assert w[1:1 + len(q)] == q
    assert w[1 + cap:1 + cap + len(kvl)] == kvl
    assert w[1 + cap + len(kvl):1 + cap + len(kvl) + len(z)] == z
    steps = [(r[2] + 63) // 64 for r in q]
    assert steps == sorted(steps, reverse=True)          # longest items first


@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_attention_with_work_lists_is_bit_identical(p_drop):
    """The work lists only change WHICH CTA runs an item and when: context, LSE and all three gradients must not
    change by a bit (masked tails, fully masked key tiles and a fully masked sequence included)."""
    from msa_b200 import capi
    g = torch.Generator().manual_seed(8)
    nh = 12
    lens = torch.randint(1, 551, (37,), generator=g).tolist() + [550, 40, 300]
    valid = [int(torch.randint(1, s + 1, (1,), generator=g)) for s in lens[:-3]] + [131, 0, 300]
    H, rows = nh * 64, sum(lens)
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    torch.manual_seed(9)
    qkv = _bf(torch.randn(rows, 3 * H, device="cuda"))
    keybias = torch.zeros(rows, device="cuda")
    for i, (s, v) in enumerate(zip(lens, valid)):
        keybias[cu[i] + v:cu[i] + s] = -10000.0
    kv_end = torch.tensor(valid, device="cuda", dtype=torch.int32)
    cu_t = torch.tensor(cu, device="cuda", dtype=torch.int32)
    dctx = _bf(torch.randn(rows, H, device="cuda"))
    work = capi.attn_schedule_buffer(len(lens), nh, max(lens), "cuda")
    capi.call("attn_schedule", capi.attn_schedule_args(cu_t, kv_end, work, nh, max(lens)))
    res = []
    for wl in (None, work):
        ctx = torch.zeros(rows, H, device="cuda", dtype=torch.bfloat16)
        lse = torch.zeros(nh, rows, device="cuda")
        dqkv = torch.full((rows, 3 * H), 7.0, device="cuda", dtype=torch.bfloat16)
        bwd_ws = capi.attn_bwd_workspace(rows, nh, "cuda")
        a = capi.attn_args(qkv, ctx, lse, keybias, cu_t, H, nh, max(lens), dctx=dctx, dqkv=dqkv, bwd_ws=bwd_ws,
                           kv_end=kv_end, p_drop=p_drop, seed=4, rng_stream=1, work=wl)
        capi.call("attn_fwd", a)
        capi.call("attn_bwd", a)
        res.append((ctx.float(), lse.clone(), dqkv.float()))
    assert torch.isfinite(res[1][2]).all()
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert torch.equal(res[0][2], res[1][2])


@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_attention_backward_skips_the_zero_gradient_query_tail_exactly(p_drop):
    """mmb_attn_schedule with row labels: when no row at or behind kv_end is labelled, the upstream gradient is exactly
    zero there and the backward skips those QUERY rows (include/mmbert_sm100.h).  With a dctx that is zero on those rows,
    all three gradients must be bit-identical to the full sweep; with a label in a tail the schedule must NOT set the
    flag (and the results stay those of the full sweep even for a dctx that is non-zero there)."""
    from msa_b200 import capi
    g = torch.Generator().manual_seed(18)
    nh = 12
    lens = torch.randint(1, 551, (45,), generator=g).tolist() + [550, 40, 300, 129, 550]
    valid = [int(torch.randint(1, s + 1, (1,), generator=g)) for s in lens[:-5]] + [131, 0, 300, 128, 64]
    H, rows = nh * 64, sum(lens)
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    torch.manual_seed(19)
    qkv = _bf(torch.randn(rows, 3 * H, device="cuda"))
    keybias = torch.zeros(rows, device="cuda")
    dctx = _bf(torch.randn(rows, H, device="cuda"))
    row_label = torch.full((rows,), -100, device="cuda", dtype=torch.int32)
    for i, (s, v) in enumerate(zip(lens, valid)):
        keybias[cu[i] + v:cu[i] + s] = -10000.0
        if v > 0:
            dctx[cu[i] + v:cu[i] + s] = 0          # the premise the labels guarantee
            row_label[cu[i] + int(torch.randint(0, v, (1,), generator=g))] = 5
    kv_end = torch.tensor(valid, device="cuda", dtype=torch.int32)
    cu_t = torch.tensor(cu, device="cuda", dtype=torch.int32)

    def run(wl, d):
        ctx = torch.zeros(rows, H, device="cuda", dtype=torch.bfloat16)
        lse = torch.zeros(nh, rows, device="cuda")
        dqkv = torch.full((rows, 3 * H), 7.0, device="cuda", dtype=torch.bfloat16)
        bwd_ws = capi.attn_bwd_workspace(rows, nh, "cuda")
        a = capi.attn_args(qkv, ctx, lse, keybias, cu_t, H, nh, max(lens), dctx=d, dqkv=dqkv, bwd_ws=bwd_ws,
                           kv_end=kv_end, p_drop=p_drop, seed=4, rng_stream=1, work=wl)
        capi.call("attn_fwd", a)
        capi.call("attn_bwd", a)
        return dqkv.float()

    plain = capi.attn_schedule_buffer(len(lens), nh, max(lens), "cuda")
    capi.call("attn_schedule", capi.attn_schedule_args(cu_t, kv_end, plain, nh, max(lens)))
    skip = capi.attn_schedule_buffer(len(lens), nh, max(lens), "cuda")
    capi.call("attn_schedule", capi.attn_schedule_args(cu_t, kv_end, skip, nh, max(lens), row_label=row_label))
    assert (int(plain[0, 2]) >> 30) == 0 and (int(skip[0, 2]) >> 30) == 1
    assert int(skip[0, 2]) & 0x3fffffff == int(plain[0, 2])
    ref, got = run(plain, dctx), run(skip, dctx)
    assert torch.isfinite(got).all()
    assert torch.equal(ref, got)
    # a label behind kv_end: the premise fails, the flag must stay clear
    bad = row_label.clone()
    i = lens.index(550)
    bad[cu[i] + valid[i] + 3] = 9
    capi.call("attn_schedule", capi.attn_schedule_args(cu_t, kv_end, skip, nh, max(lens), row_label=bad))
    assert (int(skip[0, 2]) >> 30) == 0
    dense = _bf(torch.randn(rows, H, device="cuda"))
    assert torch.equal(run(plain, dense), run(skip, dense))


@pytest.mark.gpu
@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_attention_forward_leaves_only_fully_masked_query_tiles_unwritten(p_drop):
    """mmb_attn_args.flags bit 3 (include/mmbert_sm100.h): with the zero-gradient-tail bit of the schedule set, the forward
    skips exactly the 128-query tiles that start at or behind kv_end.  Every row before ceil(kv_end / 128) * 128 must be
    bit-identical to the full forward (context and LSE), the skipped rows must keep their previous contents, the backward
    (whose upstream gradient is zero behind kv_end) must be bit-identical, and without the schedule's bit the flag is inert."""
    from msa_b200 import capi
    g = torch.Generator().manual_seed(28)
    nh = 12
    lens = torch.randint(1, 551, (40,), generator=g).tolist() + [550, 40, 300, 129, 550, 550]
    valid = [int(torch.randint(1, s + 1, (1,), generator=g)) for s in lens[:-6]] + [131, 0, 300, 128, 64, 1]
    H, rows = nh * 64, sum(lens)
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    torch.manual_seed(29)
    qkv = _bf(torch.randn(rows, 3 * H, device="cuda"))
    keybias = torch.zeros(rows, device="cuda")
    dctx = _bf(torch.randn(rows, H, device="cuda"))
    row_label = torch.full((rows,), -100, device="cuda", dtype=torch.int32)
    written = torch.ones(rows, dtype=torch.bool)
    for i, (s, v) in enumerate(zip(lens, valid)):
        keybias[cu[i] + v:cu[i] + s] = -10000.0
        if v > 0:
            dctx[cu[i] + v:cu[i] + s] = 0
            row_label[cu[i] + int(torch.randint(0, v, (1,), generator=g))] = 5
            written[cu[i] + (v + 127) // 128 * 128:cu[i] + s] = False
    assert 0.2 < float((~written).float().mean()) < 0.8
    kv_end = torch.tensor(valid, device="cuda", dtype=torch.int32)
    cu_t = torch.tensor(cu, device="cuda", dtype=torch.int32)
    written = written.cuda()

    def run(wl, flags):
        ctx = torch.full((rows, H), 3.0, device="cuda", dtype=torch.bfloat16)
        lse = torch.full((nh, rows), 5.0, device="cuda")
        dqkv = torch.full((rows, 3 * H), 7.0, device="cuda", dtype=torch.bfloat16)
        bwd_ws = capi.attn_bwd_workspace(rows, nh, "cuda")
        a = capi.attn_args(qkv, ctx, lse, keybias, cu_t, H, nh, max(lens), dctx=dctx, dqkv=dqkv, bwd_ws=bwd_ws,
                           kv_end=kv_end, p_drop=p_drop, seed=4, rng_stream=1, work=wl, flags=flags)
        capi.call("attn_fwd", a)
        capi.call("attn_bwd", a)
        return ctx.float(), lse, dqkv.float()

    plain = capi.attn_schedule_buffer(len(lens), nh, max(lens), "cuda")
    capi.call("attn_schedule", capi.attn_schedule_args(cu_t, kv_end, plain, nh, max(lens)))
    skip = capi.attn_schedule_buffer(len(lens), nh, max(lens), "cuda")
    capi.call("attn_schedule", capi.attn_schedule_args(cu_t, kv_end, skip, nh, max(lens), row_label=row_label))
    assert (int(skip[0, 2]) >> 30) == 1
    ctx0, lse0, dq0 = run(skip, 0)
    ctx1, lse1, dq1 = run(skip, 8)
    assert torch.equal(ctx0[written], ctx1[written]) and torch.equal(lse0[:, written], lse1[:, written])
    assert bool((ctx1[~written] == 3.0).all()) and bool((lse1[:, ~written] == 5.0).all())
    assert bool((ctx0[~written] != 3.0).any())
    assert torch.isfinite(dq1).all() and torch.equal(dq0, dq1)
    # flags bit 4: the caller vouches for zeros in the tiles behind kv_end -> the backward leaves them alone
    _, _, dq3 = run(skip, 8 | 16)
    assert torch.equal(dq3[written], dq1[written]) and bool((dq3[~written] == 7.0).all()) and bool((dq1[~written] == 0).all())
    # without the schedule's verdict (no row labels given) the flag changes nothing
    ctx2, lse2, dq2 = run(plain, 8)
    assert torch.equal(ctx0, ctx2) and torch.equal(lse0, lse2) and torch.equal(dq0, dq2)


def _expected_row_list(lens, kv_end, premise):
    """Restates mmb_attn_schedule_args.row_list (include/mmbert_sm100.h)."""
    live, tile, dead, row0 = [], [], [], 0
    for s, e in zip(lens, kv_end):
        e = (e if 0 < e < s else s) if premise else s
        te = min(s, (e + 127) // 128 * 128)
        live += list(range(row0, row0 + e))
        tile += list(range(row0 + e, row0 + te))
        dead += list(range(row0 + te, row0 + s))
        row0 += s
    return live, tile, dead


@pytest.mark.gpu
def test_schedule_row_list_and_the_row_kernels_that_take_it():
    """The row list of mmb_attn_schedule (live rows | rest of their attention tile | dead rows) against a Python restatement,
    with and without the premise; then the row kernels with the list: LayerNorm forward computes exactly the live rows
    (bit-identical to the full launch) and leaves the others alone, the backward is bit-identical on the live rows and
    writes zeros elsewhere without reading (NaN-filled inputs there), the column sum only reads live rows."""
    from msa_b200 import capi
    g = torch.Generator().manual_seed(31)
    lens = torch.randint(1, 551, (37,), generator=g).tolist() + [550, 128, 129, 1, 300]
    kv = [int(torch.randint(1, s + 1, (1,), generator=g)) for s in lens[:-5]] + [70, 128, 0, 1, 129]
    rows, nh, H = sum(lens), 2, 768
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    cu_t = torch.tensor(cu, device="cuda", dtype=torch.int32)
    kv_t = torch.tensor(kv, device="cuda", dtype=torch.int32)
    row_label = torch.full((rows,), -100, device="cuda", dtype=torch.int32)
    for i, (s, v) in enumerate(zip(lens, kv)):
        row_label[cu[i] + (int(torch.randint(0, v, (1,), generator=g)) if v > 0 else 0)] = 7
    work = capi.attn_schedule_buffer(len(lens), nh, max(lens), "cuda")
    rl = torch.full((4 + 2 * rows,), -1, device="cuda", dtype=torch.int32)

    def schedule(labels):
        rl.fill_(-1)
        capi.call("attn_schedule", capi.attn_schedule_args(cu_t, kv_t, work, nh, max(lens), row_label=labels, row_list=rl))
        return rl.cpu().tolist()

    for premise in (True, False):
        labels = row_label
        if not premise:
            labels = row_label.clone()
            labels[cu[37] + 549] = 3      # a label behind kv_end (sequence 37: 550 rows, kv_end 70)
        got = schedule(labels)
        live, tile, dead = _expected_row_list(lens, kv, premise)
        assert got[:4] == [len(live), len(tile), rows, 1 if premise else 0]
        assert got[4:4 + len(live)] == live
        assert got[4 + len(live):4 + len(live) + len(tile)] == tile
        assert sorted(got[4 + len(live) + len(tile):4 + rows]) == dead
        flags = [0] * rows
        for r in live:
            flags[r] = 1
        assert got[4 + rows:] == flags          # per-row live flags (mmb_gemm_args.row_live)
    got = schedule(row_label)
    live, tile, dead = _expected_row_list(lens, kv, True)
    assert 0.2 < len(live) / rows < 0.8
    is_live = torch.zeros(rows, dtype=torch.bool, device="cuda")
    is_live[torch.tensor(live, device="cuda")] = True

    torch.manual_seed(32)
    y, res = _bf(torch.randn(rows, H, device="cuda")), torch.randn(rows, H, device="cuda")
    gamma, beta = torch.randn(H, device="cuda"), torch.randn(H, device="cuda")

    def fwd(row_list, p):
        out = torch.full((rows, H), 3.0, device="cuda", dtype=torch.bfloat16)
        out32, mean, rstd = torch.full((rows, H), 3.0, device="cuda"), torch.full((rows,), 3.0, device="cuda"), torch.full((rows,), 3.0, device="cuda")
        capi.drln_fwd(y, res, gamma, beta, out, mean, rstd, 1e-12, p_drop=p, seed=5, rng_stream=2, out_f32=out32, row_list=row_list)
        return out, out32, mean, rstd

    for p in (0.0, 0.1):
        full, part = fwd(None, p), fwd(rl, p)
        for a, b in zip(full, part):
            assert torch.equal(a[is_live], b[is_live])
            assert bool((b[~is_live] == 3.0).all())
        out, out32, mean, rstd = full
        g1, g2 = _bf(torch.randn(rows, H, device="cuda")), torch.randn(rows, H, device="cuda")
        g1[~is_live] = 0
        g2[~is_live] = 0

        def bwd(row_list, poison, zeroed=0):
            g1_, g2_, y_, res_ = (t.clone() for t in (g1, g2, y, res))
            if poison:          # nothing may be read on the other rows
                for t in (g1_, g2_, y_, res_):
                    t[~is_live] = float("nan")
            d_y = torch.full((rows, H), 5.0, device="cuda", dtype=torch.bfloat16)
            d_res = torch.full((rows, H), 5.0, device="cuda")
            dgamma, dbeta, dbias = (torch.zeros(H, device="cuda") for _ in range(3))
            capi.drln_bwd(g1_, g2_, y_, res_, mean, rstd, gamma, d_y, d_res, dgamma, dbeta, dbias, p_drop=p, seed=5, rng_stream=2,
                          row_list=row_list, dead_rows_zeroed=zeroed)
            return d_y, d_res, dgamma, dbeta, dbias

        fb, pb = bwd(None, False), bwd(rl, True)
        assert torch.equal(fb[0][is_live], pb[0][is_live]) and torch.equal(fb[1][is_live], pb[1][is_live])
        assert bool((pb[0][~is_live] == 0).all()) and bool((pb[1][~is_live] == 0).all())
        assert bool((fb[0][~is_live] == 0).all())          # the full launch computes the same zeros
        zb = bwd(rl, True, zeroed=1)                        # caller-guaranteed zeros: the other rows are not written at all
        assert torch.equal(zb[0][is_live], pb[0][is_live]) and torch.equal(zb[1][is_live], pb[1][is_live])
        assert bool((zb[0][~is_live] == 5.0).all()) and bool((zb[1][~is_live] == 5.0).all())
        for a, b in zip(fb[2:], pb[2:]):
            assert torch.isfinite(b).all() and _rel(b, a) < 1e-5          # fp32 atomics: order only

    X = _bf(torch.randn(rows, 2304, device="cuda"))
    X[~is_live] = 0
    ref = torch.zeros(2304, device="cuda")
    capi.colsum(X, ref)
    Xp = X.clone()
    Xp[~is_live] = float("nan")
    got_sum = torch.zeros(2304, device="cuda")
    capi.call("colsum_bf16", capi.colsum_args(Xp, got_sum, row_list=rl))
    assert torch.isfinite(got_sum).all() and _rel(got_sum, ref) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("M,N", [(1000, 3072), (520, 768)])
def test_gemm_row_live_hint_skips_only_all_padding_slices(M, N):
    """mmb_gemm_args.row_live: a 32-row epilogue slice without a live row is left unwritten by the GELU epilogues and
    zero-filled by the multiply epilogue (whose column sums then only see the written rows); every slice that holds a
    live row — and every other epilogue — is bit-identical to the call without the hint."""
    from msa_b200 import capi
    torch.manual_seed(6)
    K = 768
    A, B = _bf(torch.randn(M, K, device="cuda") * 0.3), _bf(torch.randn(N, K, device="cuda") * 0.3)
    bias = torch.randn(N, device="cuda")
    live = torch.zeros(M, device="cuda", dtype=torch.int32)
    live[:70] = 1
    live[200:300] = 1          # slices 6..9 hold live rows (rows 192..319), the rest of 96..191 and 320.. are dead
    live[M - 3] = 1            # a live row in the ragged last slice
    slice_live = live.view(-1)[: M // 32 * 32].view(-1, 32).amax(1).repeat_interleave(32)
    slice_live = torch.cat((slice_live, live[M // 32 * 32:].amax().expand(M - M // 32 * 32))).bool()
    assert 0.2 < float(slice_live.float().mean()) < 0.8

    def run(epi, hint, **kw):
        C = torch.full((M, N), 9.0, device="cuda", dtype=torch.bfloat16)
        aux_out = torch.full((M, N), 9.0, device="cuda", dtype=torch.bfloat16) if epi != capi.EPI_MUL_AUX_BF16 else None
        cs = torch.zeros(N, device="cuda")
        if epi == capi.EPI_MUL_AUX_BF16:
            capi.gemm(A, B, C, M, N, K, epilogue=epi, colsum=cs, row_live=live if hint else None, **kw)
        else:
            capi.gemm(A, B, C, M, N, K, epilogue=epi, bias=bias, aux=aux_out, row_live=live if hint else None)
        return C, aux_out, cs

    for epi in (capi.EPI_GELU_GRAD_BF16, capi.EPI_GELU_BF16):
        (c0, a0, _), (c1, a1, _) = run(epi, False), run(epi, True)
        assert torch.equal(c0[slice_live], c1[slice_live]) and torch.equal(a0[slice_live], a1[slice_live])
        assert bool((c1[~slice_live] == 9.0).all()) and bool((a1[~slice_live] == 9.0).all())
    aux = _bf(torch.randn(M, N, device="cuda"))
    poisoned = aux.clone()
    poisoned[~slice_live] = float("nan")           # nothing is read there
    (c0, _, s0), (c1, _, s1) = run(capi.EPI_MUL_AUX_BF16, False, aux=aux), run(capi.EPI_MUL_AUX_BF16, True, aux=poisoned)
    assert torch.equal(c0[slice_live], c1[slice_live])
    assert bool((c1[~slice_live] == 0).all())
    assert torch.isfinite(s1).all() and _rel(s1, c1.float().sum(0)) < 1e-5 and _rel(s0, c0.float().sum(0)) < 1e-5
    c3 = torch.full((M, N), 9.0, device="cuda", dtype=torch.bfloat16)       # dead_rows_zeroed: those slices are not written
    capi.gemm(A, B, c3, M, N, K, epilogue=capi.EPI_MUL_AUX_BF16, aux=poisoned, row_live=live, dead_rows_zeroed=1)
    assert torch.equal(c3[slice_live], c1[slice_live]) and bool((c3[~slice_live] == 9.0).all())
    # the plain store leaves all-dead slices unwritten as well (not when it also takes column sums) ...
    c2, c4 = (torch.full((M, N), 9.0, device="cuda", dtype=torch.bfloat16) for _ in range(2))
    capi.gemm(A, B, c2, M, N, K, epilogue=capi.EPI_STORE_BF16, bias=bias, row_live=live)
    want = A.float() @ B.float().t() + bias
    assert _rel(c2.float()[slice_live], want[slice_live]) < 2 * BF16_EPS and bool((c2[~slice_live] == 9.0).all())
    cs = torch.zeros(N, device="cuda")
    capi.gemm(A, B, c4, M, N, K, epilogue=capi.EPI_STORE_BF16, bias=bias, row_live=live, colsum=cs)
    assert _rel(c4.float(), want) < 2 * BF16_EPS and _rel(cs, c4.float().sum(0)) < 1e-5
    # ... and an epilogue that does not take the hint computes every row
    capi.gemm(A, B, c2, M, N, K, epilogue=capi.EPI_RELU_BF16, bias=bias, row_live=live)
    assert _rel(c2.float(), torch.relu(want)) < 2 * BF16_EPS
