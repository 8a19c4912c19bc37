"""Unit parity of the non-GEMM stages on IDENTICAL inputs (so ReLU gates / rounding upstream cannot differ):
the fused heads, the fused embeddings and the vocabulary cross entropy, forward and backward, against plain torch
restatements of the reference lines they replace.  fp32 accumulations -> tight tolerances."""
import pytest
import torch

from msa_b200 import synth
from msa_b200.params import BertShape, seeded_state_dict
from oracle import mmbert_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _model(shape, dataset, sd):
    from msa_b200.api import MMBertForPretraining
    m = MMBertForPretraining(shape)
    m.bert.set_joint_embeddings(dataset)
    m.bert.jointEmbeddings.dropout.p = 0.0
    m.load_state_dict(sd)
    return m.cuda().train()


def _small_shape(**kw):
    return BertShape(hidden_size=128, num_hidden_layers=1, num_attention_heads=2, intermediate_size=256, vocab_size=300,
                     max_position_embeddings=64, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **kw)


@pytest.mark.parametrize("num_labels", [7, 1])
def test_heads_forward_backward_exact(num_labels):
    """Feeds the heads the encoder output the CUDA path itself produced and restates
    MMBertForPretraining.py:295-302,406-443 + MMBertEmbedding.py:21-32 in fp64 torch on that same tensor."""
    shape = _small_shape()
    sd = seeded_state_dict(shape, "mosi", seed=21)
    m = _model(shape, "mosi", sd)
    m.num_labels = num_labels
    m.set_alpha_beta(0.7, 0.4)
    B, T, L = 6, 10, 10
    batch = synth.make_batch(B, T, L, L, 47, 74, vocab_size=300, seed=4, min_len=4)
    out, logits = m(**synth.tree_to(batch, "cuda"))
    plan = next(iter(m._plans.values()))
    out[0].backward()
    torch.cuda.synchronize()
    # ---- torch restatement on the same seq_out ([CLS] rows) and the same MLM statistics
    cu = plan.cu.cpu().tolist()
    x0 = plan.seq_out.float().cpu().double()[cu[:-1]].requires_grad_(True)          # [3B, H]
    P = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()
         if k.split(".")[0] in ("attn", "vt", "vv", "vs", "classifier1_1", "classifier1_2", "cpc_zt", "cpc_zv", "cpc_za")
         or k.startswith(("bert.pooler", "cls.align"))}
    lin = lambda x, n: x @ P[n + ".weight"].t() + P[n + ".bias"]
    pooled = torch.tanh(lin(x0, "bert.pooler.dense"))
    p_t, p_v, p_s = pooled[:B], pooled[B:2 * B], pooled[2 * B:]
    al_v, al_s = lin(x0[B:2 * B], "cls.align"), lin(x0[2 * B:], "cls.align")
    ce = torch.nn.functional.cross_entropy
    ap = (ce(al_v, batch["ap_label"][0]) + ce(al_s, batch["ap_label"][1])) / 2
    score = lambda p, v: lin(torch.relu(lin(torch.cat((p, p), 1), "attn")), v)
    cat = torch.cat((p_t * score(p_t, "vt"), p_v * score(p_v, "vv"), p_s * score(p_s, "vs")), 1)
    temp = lin(cat, "classifier1_1")
    lg = lin(temp, "classifier1_2")

    def cpc(name, x, y):
        xp = lin(y, name + ".net")
        xp = xp / xp.norm(dim=1, keepdim=True)
        x = x / x.norm(dim=1, keepdim=True)
        return -((x * xp).sum(-1) - torch.logsumexp(x @ xp.t(), -1)).mean()

    nce = cpc("cpc_zt", p_t, temp) + cpc("cpc_zv", p_v, temp) + cpc("cpc_za", p_s, temp)
    lo = torch.tanh(lg) if num_labels == 1 else lg
    label = ((lo.view(-1) - batch["sentiment"].double()) ** 2).mean()
    mlm = float(plan.losses[1])
    joint = 0.7 * mlm + ap + label - 0.4 * nce
    joint.backward()
    assert rel_err(plan.losses[2], ap) < 1e-5 and rel_err(plan.losses[3], label) < 1e-5
    assert rel_err(plan.losses[4], nce) < 1e-5 and rel_err(plan.losses[0], joint) < 1e-5
    assert rel_err(logits, lo) < 1e-5
    assert rel_err(out[10], al_v) < 1e-5 and rel_err(out[12], al_s) < 1e-5
    named = dict(m.named_parameters())
    for n, t in P.items():
        assert rel_err(named[n].grad, t.grad, floor=1e-7) < 2e-4, n


def test_embeddings_forward_exact_and_backward():
    """x0 of the CUDA path vs BertEmbeddings + JointEmbeddings restated in torch (oracle.bert_pass with 0 layers), at
    UR-FUNNY's frame dims (371 / 81: the padded-leading-dimension tensor-core projection wgrad), forward AND backward:
    mmb_embed_bwd is run alone on a known upstream gradient and every gradient it produces is compared with autograd
    through the restatement (word / position / token-type tables, both LayerNorms, Wv / Ws and their biases)."""
    from msa_b200.engine import Plan
    shape = _small_shape()
    shape.num_hidden_layers = 1
    sd = seeded_state_dict(shape, "ur_funny", seed=22)
    m = _model(shape, "ur_funny", sd)
    B, T, Lv, La = 3, 9, 14, 5
    batch = synth.make_batch(B, T, Lv, La, 371, 81, vocab_size=300, seed=6, min_len=4)
    batch["token_type_ids"][0][:, 3:] = 1          # exercise token type 1 on the text pass
    m(**synth.tree_to(batch, "cuda"))
    plan = next(iter(m._plans.values()))
    ocfg = O.Cfg(128, 0, 2, 256, 300, 64)
    names = ("bert.embeddings.word_embeddings.weight", "bert.embeddings.position_embeddings.weight",
             "bert.embeddings.token_type_embeddings.weight", "bert.embeddings.LayerNorm.weight", "bert.embeddings.LayerNorm.bias",
             "bert.jointEmbeddings.LayerNorm.weight", "bert.jointEmbeddings.LayerNorm.bias",
             "bert.jointEmbeddings.Wv.weight", "bert.jointEmbeddings.Wv.bias",
             "bert.jointEmbeddings.Ws.weight", "bert.jointEmbeddings.Ws.bias")
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    for n in names:
        sd64[n] = sd64[n].clone().requires_grad_(True)
    ids_t, vis, aud, ids_v, ids_s = batch["input_ids"]
    m_t, (m_tv, m_v), (m_ts, m_s) = batch["attention_mask"]
    ref = torch.cat([O.bert_pass(sd64, ocfg, ids_t, m_t, batch["token_type_ids"][0])[0].reshape(-1, 128),
                     O.bert_pass(sd64, ocfg, ids_v, m_tv, None, vis, m_v)[0].reshape(-1, 128),
                     O.bert_pass(sd64, ocfg, ids_s, m_ts, None, aud, m_s)[0].reshape(-1, 128)])
    # padding rows (behind the last unmasked key of their sequence) may be left alone by the plan: mmb_embed_args.row_live
    live = torch.ones(plan.M, dtype=torch.bool, device="cuda") if plan.row_live is None else plan.row_live.bool()
    assert 0.3 < float(live.float().mean()) <= 1.0
    # fp32 residual-stream copy: only the bf16 rounding of the frame projection separates it from the oracle
    assert rel_err(plan.x32[0][live], ref.detach()[live.cpu()]) < 6e-3
    assert rel_err(plan.x[0].float()[live], ref.detach()[live.cpu()]) < 8e-3
    # ---- backward: dL/dx0 = g (bf16 part) + g32 (fp32 residual-stream part)
    torch.manual_seed(7)
    g = torch.randn(plan.M, 128, device="cuda").to(torch.bfloat16)
    g32 = torch.randn(plan.M, 128, device="cuda") * 0.5
    g[~live] = 0            # the premise of the padding-row skip: no gradient reaches a padding row
    g32[~live] = 0
    m._prepare_grads()
    m._store.grad.zero_()
    plan.GA.copy_(g)
    plan.GB.copy_(g32)
    Plan.run([plan.bwd[-1]])                        # mmb_embed_bwd alone
    torch.cuda.synchronize()
    (ref * (g.float() + g32).cpu().double()).sum().backward()
    named = dict(m.named_parameters())
    for n in names:
        want = sd64[n].grad
        tol = 2e-2 if n.endswith(("Wv.weight", "Ws.weight")) else 1e-2     # projection wgrad: bf16 frames x bf16 dpre
        assert rel_err(named[n].grad, want, floor=1e-6) < tol, (n, rel_err(named[n].grad, want, floor=1e-6))
    assert float(named["bert.embeddings.word_embeddings.weight"].grad[0].abs().max()) == 0.0     # padding_idx row


def test_vocab_cross_entropy_exact():
    """mmb_ce_fwd / mmb_ce_bwd vs torch.nn.functional.cross_entropy(ignore_index=-100) on the same bf16 logits."""
    from msa_b200 import capi
    torch.manual_seed(3)
    B, T, L, V = 5, 7, 7, 30522
    Vp = (V + 7) // 8 * 8
    rows = B * (3 * T + 2 * L)
    logits = (torch.randn(rows, Vp, device="cuda") * 3).to(torch.bfloat16)
    labs = [torch.full((B, T), -100, dtype=torch.int64), torch.full((B, T + L), -100, dtype=torch.int64),
            torch.full((B, T + L), -100, dtype=torch.int64)]
    g = torch.Generator().manual_seed(1)
    for t in labs:
        sel = torch.rand(t.shape, generator=g) < 0.2
        t[sel] = torch.randint(0, V, (int(sel.sum()),), generator=g)
        t[0, 0] = V - 1                       # the last vocabulary column (tail handling)
    labs_d = [t.cuda() for t in labs]
    count = torch.tensor([int((t != -100).sum()) for t in labs] + [0], device="cuda", dtype=torch.int32)
    row_lse, loss_sum = torch.zeros(rows, device="cuda"), torch.zeros(4, device="cuda")
    dlogits = torch.full((rows, Vp), 7.0, device="cuda", dtype=torch.bfloat16)
    gs = torch.tensor([0.5], device="cuda")
    a = capi.fill(capi.CeArgs(), logits=logits, dlogits=dlogits, labels=labs_d, label_count=count, row_lse=row_lse,
                  loss_sum=loss_sum, gscale=gs, coef=1.0 / 3, V=V, ldl=Vp, B=B, T=T, L=[L, L], dense=1)
    capi.call("ce_fwd", a)
    capi.call("ce_bwd", a)
    x = logits[:, :V].float().requires_grad_(True)
    flat = torch.cat([t.reshape(-1) for t in labs_d])
    bounds = [0, B * T, B * T + B * (T + L), rows]
    losses = [torch.nn.functional.cross_entropy(x[bounds[i]:bounds[i + 1]], flat[bounds[i]:bounds[i + 1]]) for i in range(3)]
    for i in range(3):
        assert rel_err(loss_sum[i] / count[i], losses[i]) < 1e-5
    (0.5 * sum(losses) / 3).backward()
    assert rel_err(dlogits[:, :V].float(), x.grad, floor=1e-9) < 2 ** -7      # bf16 output
    assert float(dlogits[:, V:].float().abs().max()) == 0.0                    # padding columns are zeroed
