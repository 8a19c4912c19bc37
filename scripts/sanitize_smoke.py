"""Small launches of every warp-specialised kernel (CTA-pair GEMM with each bf16 epilogue and the fused column sums,
single-CTA GEMM, attention forward / backward with work lists and dropout, LayerNorm forward / backward, fused cross
entropy) for compute-sanitizer:
    compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_smoke.py
Shapes are tiny (the tools slow these kernels by 10-1000x); results are checked against torch so that a tool-induced
timing change that breaks a protocol shows up as a wrong answer, not only as a report."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msa_b200 import capi

torch.manual_seed(0)
dev = "cuda"
bf = lambda t: t.to(torch.bfloat16)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


# ---- GEMM
M, N, K = 600, 512, 320
A, B = bf(torch.randn(M, K, device=dev) * 0.3), bf(torch.randn(N, K, device=dev) * 0.3)
bias = torch.randn(N, device=dev)
ref = A.float() @ B.float().t()
C, aux = torch.empty(M, N, device=dev, dtype=torch.bfloat16), torch.empty(M, N, device=dev, dtype=torch.bfloat16)
cs = torch.zeros(N, device=dev)
capi.gemm(A, B, C, M, N, K, bias=bias, colsum=cs)
assert rel(C.float(), ref + bias) < 1e-2 and rel(cs, C.float().sum(0)) < 1e-4
capi.gemm(A, B, C, M, N, K, epilogue=capi.EPI_GELU_GRAD_BF16, bias=bias, aux=aux)
assert rel(C.float(), torch.nn.functional.gelu(ref + bias)) < 1e-2
capi.gemm(A, B, C, M, N, K, epilogue=capi.EPI_MUL_AUX_BF16, aux=aux)
assert rel(C.float(), ref * aux.float()) < 1e-2
acc = torch.zeros(N, K, device=dev)
X = bf(torch.randn(M, N, device=dev) * 0.3)
capi.gemm(X, A, acc, N, K, M, a_major=capi.MAJOR_MN, b_major=capi.MAJOR_MN, epilogue=capi.EPI_ATOMIC_ADD_F32, split_k=2)
assert rel(acc, X.float().t() @ A.float()) < 1e-3
C1 = torch.empty(100, 96, device=dev, dtype=torch.bfloat16)
capi.gemm(A[:100], B[:96], C1, 100, 96, K)            # single-CTA kernel
assert rel(C1.float(), ref[:100, :96]) < 1e-2
# fused cross entropy epilogue
V = 520
W = bf(torch.randn(V, K, device=dev) * 0.3)
lab = torch.full((M,), -100, device=dev, dtype=torch.int32)
lab[::7] = torch.randint(0, V, (len(lab[::7]),), device=dev, dtype=torch.int32)
stats = torch.zeros(capi.ce_stats_floats(M, V), device=dev)
dl = torch.zeros(M, 520, device=dev, dtype=torch.bfloat16)
capi.gemm(A, W, dl, M, V, K, epilogue=capi.EPI_CE_STATS, aux=lab, aux2=stats, ldaux=0)
G = (V + 127) // 128
logits = A.float() @ W.float().t()
rows = (lab != -100).nonzero()[:, 0]
m = stats.view(2 * G + 2, M)[0:2 * G:2, rows].max(0).values
assert rel(m, logits[rows].max(1).values) < 1e-2
# ---- attention
nh, lens, valid = 2, [200, 70, 129], [150, 70, 0]
H, rows_ = nh * 64, sum(lens)
cu = torch.tensor([0, 200, 270, 399], device=dev, dtype=torch.int32)
qkv = bf(torch.randn(rows_, 3 * H, device=dev))
keybias = torch.zeros(rows_, device=dev)
keybias[150:200] = -10000.0
kv_end = torch.tensor(valid, device=dev, dtype=torch.int32)
row_label = torch.full((rows_,), -100, device=dev, dtype=torch.int32)
row_label[3] = 5
work = capi.attn_schedule_buffer(3, nh, 200, dev)
capi.call("attn_schedule", capi.attn_schedule_args(cu, kv_end, work, nh, 200, row_label=row_label))
ctx = torch.zeros(rows_, H, device=dev, dtype=torch.bfloat16)
lse = torch.zeros(nh, rows_, device=dev)
dctx = bf(torch.randn(rows_, H, device=dev))
dctx[150:200] = 0
dqkv = torch.zeros(rows_, 3 * H, device=dev, dtype=torch.bfloat16)
ws = capi.attn_bwd_workspace(rows_, nh, dev)
for p in (0.0, 0.1):
    a = capi.attn_args(qkv, ctx, lse, keybias, cu, H, nh, 200, dctx=dctx, dqkv=dqkv, bwd_ws=ws, kv_end=kv_end, p_drop=p, seed=3,
                       rng_stream=1, work=work)
    capi.call("attn_fwd", a)
    capi.call("attn_bwd", a)
    assert torch.isfinite(ctx.float()).all() and torch.isfinite(dqkv.float()).all()
q, k, v = (qkv[:200, i * H:i * H + 64].float() for i in range(3))
pr = torch.softmax(q @ k.t() / 8 + keybias[:200][None, :], -1)
capi.call("attn_fwd", capi.attn_args(qkv, ctx, lse, keybias, cu, H, nh, 200, kv_end=kv_end, work=work))
assert rel(ctx[:200, :64].float(), pr @ v) < 2e-2
# ---- LayerNorm
Mr, Hh = 300, 768
y, g1 = bf(torch.randn(Mr, Hh, device=dev)), bf(torch.randn(Mr, Hh, device=dev))
res, g2 = torch.randn(Mr, Hh, device=dev), torch.randn(Mr, Hh, device=dev)
gamma, beta = torch.randn(Hh, device=dev), torch.randn(Hh, device=dev)
out, d_y = torch.empty_like(y), torch.empty_like(y)
out32, d_res = torch.empty_like(res), torch.empty_like(res)
mean, rstd = torch.empty(Mr, device=dev), torch.empty(Mr, device=dev)
dg, db, dbias = torch.zeros(Hh, device=dev), torch.zeros(Hh, device=dev), torch.zeros(Hh, device=dev)
capi.drln_fwd(y, res, gamma, beta, out, mean, rstd, 1e-12, p_drop=0.1, seed=5, rng_stream=2, out_f32=out32)
capi.drln_bwd(g1, g2, y, res, mean, rstd, gamma, d_y, d_res, dg, db, dbias, p_drop=0.1, seed=5, rng_stream=2)
assert torch.isfinite(d_res).all() and torch.isfinite(dg).all()
# ---- padding rows: the schedule's row list and every kernel that takes it / its per-row flags
rl = torch.zeros(4 + 2 * rows_, device=dev, dtype=torch.int32)
capi.call("attn_schedule", capi.attn_schedule_args(cu, kv_end, work, nh, 200, row_label=row_label, row_list=rl))
hdr = rl[:4].tolist()
assert hdr == [150 + 70 + 129, 50, rows_, 1], hdr            # sequence 0: 150 live + 50 in its second tile; 1 and 2 all live
live = rl[4 + rows_:].bool()
ctx_s, lse_s, dq_s = torch.full_like(ctx, 3.0), torch.full_like(lse, 5.0), torch.full_like(dqkv, 7.0)
for flags in (8, 8 | 16):
    a = capi.attn_args(qkv, ctx_s, lse_s, keybias, cu, H, nh, 200, dctx=dctx, dqkv=dq_s, bwd_ws=ws, kv_end=kv_end, p_drop=0.1, seed=3,
                       rng_stream=1, work=work, flags=flags, row_list=rl)
    capi.call("attn_fwd", a)
    capi.call("attn_bwd", a)
    assert torch.isfinite(ctx_s.float()).all() and torch.isfinite(dq_s.float()).all()
Mr2 = rows_
y2, g12 = bf(torch.randn(Mr2, Hh, device=dev)), bf(torch.randn(Mr2, Hh, device=dev))
res2, g22 = torch.randn(Mr2, Hh, device=dev), torch.randn(Mr2, Hh, device=dev)
out2, dy2 = torch.zeros_like(y2), torch.zeros_like(y2)
out322, dres2 = torch.zeros_like(res2), torch.zeros_like(res2)
mean2, rstd2 = torch.zeros(Mr2, device=dev), torch.zeros(Mr2, device=dev)
capi.drln_fwd(y2, res2, gamma, beta, out2, mean2, rstd2, 1e-12, p_drop=0.1, seed=5, rng_stream=2, out_f32=out322, row_list=rl)
for zeroed in (0, 1):
    capi.drln_bwd(g12, g22, y2, res2, mean2, rstd2, gamma, dy2, dres2, dg, db, dbias, p_drop=0.1, seed=5, rng_stream=2, row_list=rl,
                  dead_rows_zeroed=zeroed)
assert bool((out2[~live] == 0).all()) and bool((dres2[~live] == 0).all()) and torch.isfinite(dres2).all()
csl = torch.zeros(Hh, device=dev)
capi.call("colsum_bf16", capi.colsum_args(dy2, csl, row_list=rl))
assert rel(csl, dy2.float().sum(0)) < 1e-4
Ar = bf(torch.randn(rows_, K, device=dev) * 0.3)
Cr, auxr = torch.zeros(rows_, N, device=dev, dtype=torch.bfloat16), torch.zeros(rows_, N, device=dev, dtype=torch.bfloat16)
refr = Ar.float() @ B.float().t()
capi.gemm(Ar, B, Cr, rows_, N, K, epilogue=capi.EPI_GELU_GRAD_BF16, bias=bias, aux=auxr, row_live=rl[4 + rows_:])
assert rel(Cr.float()[live], torch.nn.functional.gelu(refr + bias)[live]) < 1e-2
for zeroed in (0, 1):
    capi.gemm(Ar, B, Cr, rows_, N, K, epilogue=capi.EPI_MUL_AUX_BF16, aux=auxr, colsum=cs, row_live=rl[4 + rows_:], dead_rows_zeroed=zeroed)
assert rel(Cr.float()[live], (refr * auxr.float())[live]) < 1e-2
capi.gemm(Ar, B, Cr, rows_, N, K, bias=bias, row_live=rl[4 + rows_:])
assert rel(Cr.float()[live], (refr + bias)[live]) < 1e-2
torch.cuda.synchronize()
print("sanitize_smoke: all launches done, results correct")
