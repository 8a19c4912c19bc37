"""TEST / BENCH INFRASTRUCTURE — the MMBert hot path written the way a PyTorch user would run the reference on a GPU:
the same arithmetic as oracle/mmbert_oracle.py (pinned against the reference's golden vectors; this file is checked
against that oracle in tests/test_oracle.py) but on torch's FUSED library ops — ``F.linear`` (cuBLASLt),
``F.scaled_dot_product_attention`` (what transformers 5.x dispatches BertSelfAttention to by default),
``F.layer_norm``, ``F.gelu``, ``F.cross_entropy``, ``F.dropout`` — so that it can be timed on the B200 in fp32
(TF32 off / on) and under ``torch.autocast(bfloat16)`` as the "PyTorch / cuBLASLt / SDPA bar" of BASELINE.md §4.5.

/root/reference is Python on top of transformers and does not exist on the GPU box, so this port stands in for it
there; only bench.py's ``gpu_torch_baseline`` leg and tests/ may import it — never the product path.
Reference lines: MMBertForPretraining.py:216-285, 292-302, 367-449; MMBertEmbedding.py:57-72;
transformers modeling_bert.py:72-112, 168-207, 294-298, 339-342, 352-356, 481-501.
"""
import torch
import torch.nn.functional as F

from . import mmbert_oracle as O


def _lin(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _ln(x, sd, name, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def bert_pass(sd, cfg, ids, mask, token_type, frames=None, frame_mask=None, p=(0.0, 0.0, 0.0), training=False):
    H, nh = cfg.hidden_size, cfg.num_attention_heads
    d = H // nh
    B, T = ids.shape
    p_h, p_a, p_j = p if training else (0.0, 0.0, 0.0)
    w = sd["bert.embeddings.word_embeddings.weight"]
    if frames is not None:
        token_type = torch.zeros_like(ids)
    e = F.embedding(ids.long(), w, padding_idx=0) + sd["bert.embeddings.token_type_embeddings.weight"][token_type.long()] \
        + sd["bert.embeddings.position_embeddings.weight"][:T][None]
    x = F.dropout(_ln(e, sd, "bert.embeddings.LayerNorm", cfg.layer_norm_eps), p_h, training)
    ext = O._ext_mask(mask, w.dtype)
    if frames is not None:
        which = "Wv" if frames.shape[-1] == sd["bert.jointEmbeddings.Wv.weight"].shape[1] else "Ws"
        pe = torch.relu(_lin(frames.float().to(w.dtype), sd, "bert.jointEmbeddings." + which))
        x = F.dropout(_ln(torch.cat((x, pe.to(x.dtype)), dim=1), sd, "bert.jointEmbeddings.LayerNorm", 1e-5), p_j, training)
        ext = torch.cat((ext, O._ext_mask(frame_mask, w.dtype)), dim=-1)
    S = x.shape[1]
    for i in range(cfg.num_hidden_layers):
        pre = f"bert.encoder.layer.{i}."
        q = _lin(x, sd, pre + "attention.self.query").view(B, S, nh, d).transpose(1, 2)
        k = _lin(x, sd, pre + "attention.self.key").view(B, S, nh, d).transpose(1, 2)
        v = _lin(x, sd, pre + "attention.self.value").view(B, S, nh, d).transpose(1, 2)
        ctx = F.scaled_dot_product_attention(q, k, v, attn_mask=ext.to(q.dtype), dropout_p=p_a)
        ctx = ctx.transpose(1, 2).reshape(B, S, H)
        a = _ln(F.dropout(_lin(ctx, sd, pre + "attention.output.dense"), p_h, training) + x, sd,
                pre + "attention.output.LayerNorm", cfg.layer_norm_eps)
        h = F.gelu(_lin(a, sd, pre + "intermediate.dense"))
        x = _ln(F.dropout(_lin(h, sd, pre + "output.dense"), p_h, training) + a, sd, pre + "output.LayerNorm",
                cfg.layer_norm_eps)
    return x


def lm_head(sd, cfg, seq):
    t = _ln(F.gelu(_lin(seq, sd, "cls.predictions.transform.dense")), sd, "cls.predictions.transform.LayerNorm",
            cfg.layer_norm_eps)
    return F.linear(t, sd["bert.embeddings.word_embeddings.weight"], sd["cls.predictions.bias"])


def forward(sd, cfg, input_ids, token_type_ids, attention_mask, masked_labels, ap_label, sentiment, alpha=1.0, beta=1.0,
            p=(0.1, 0.1, 0.5), training=False):
    """((13-tuple), logits) like MMBertForPretraining.forward; ``sd`` holds the parameters in the compute dtype / device."""
    ids_t, vis, aud, ids_v, ids_s = input_ids
    m_t, (m_tv, m_v), (m_ts, m_s) = attention_mask
    lab_t, lab_v, lab_s = masked_labels
    V = cfg.vocab_size
    seq_t = bert_pass(sd, cfg, ids_t, m_t, token_type_ids[0], p=p, training=training)
    seq_v = bert_pass(sd, cfg, ids_v, m_tv, None, vis, m_v, p=p, training=training)
    seq_s = bert_pass(sd, cfg, ids_s, m_ts, None, aud, m_s, p=p, training=training)
    pred_t, pred_v, pred_s = lm_head(sd, cfg, seq_t), lm_head(sd, cfg, seq_v), lm_head(sd, cfg, seq_s)
    up = lambda t: t.float() if t.dtype in (torch.bfloat16, torch.float16) else t      # autocast leaves the logits in bf16
    ce = lambda pred, lab: F.cross_entropy(up(pred.reshape(-1, V)), lab.reshape(-1), ignore_index=-100)
    mlm = (ce(pred_t, lab_t) + ce(pred_v, lab_v) + ce(pred_s, lab_s)) / 3.0
    x0 = torch.cat((seq_t[:, 0], seq_v[:, 0], seq_s[:, 0]), dim=0)
    ap, label, nce, out_logits, rel_t, al_v, al_s = O.heads(sd, cfg, x0, ap_label[0], ap_label[1], sentiment, dtype=x0.dtype)
    joint = alpha * mlm + ap + label - beta * nce
    return (joint, None, None, None, ap, label, nce, pred_t, rel_t, pred_v, al_v, pred_s, al_s), out_logits
