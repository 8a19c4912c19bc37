"""TEST INFRASTRUCTURE — CPU restatement ("port") of the reference's MMBert hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module, and only as the checker (or the timed CPU baseline) — never as the product path.

It restates, as plain functional torch on a ``state_dict`` (no nn.Module, no transformers import), the
arithmetic of:
  * MMBertForPretraining.forward            /root/reference/MMBertForPretraining.py:392-449
  * MMBertForPretraining.get_outputs        :367-390      (CE with ignore_index -100, mean over valid)
  * MMBertModel.forward + extended mask     :216-285, :57-154  ((1-m)*-10000, frame mask = feature 0 only)
  * MMBertPreTrainingHeads.forward          :292-302
  * JointEmbeddings.forward, CPC.forward    /root/reference/MMBertEmbedding.py:57-72, :21-32
  * transformers 5.5.0 modeling_bert.py (third-party, not vendored in the reference; the reference pins
    transformers==2.8.0 but only imports that exist in >=4.x): BertEmbeddings :72-112, eager attention
    :115-140, BertSelfAttention :168-207, BertSelfOutput :294-298, BertIntermediate :339-342 (exact erf
    GELU), BertOutput :352-356, BertPooler :462-468, BertPredictionHeadTransform :481-485,
    BertLMPredictionHead :498-501 (decoder weight tied to the word embeddings, :733-736).

Pinning: the reference ships no tests or golden vectors, so this restatement is pinned
against the live reference executed in the build container: ``tests/golden/*.npz`` are produced by
``tests/golden/make_golden.py`` from the unmodified reference, and ``tests/test_oracle.py`` checks this file
against them (all 13 outputs + logits, every parameter gradient and the set of parameters without
gradient) — and, where /root/reference is present, against the live reference directly.

Dropout: the parity arithmetic is the p=0 / eval() one (RNG streams cannot match).  For STATISTICAL comparisons of
training runs (K-step sentiment-MAE test) ``forward(..., dropout=Dropout(...))`` applies torch dropout at the five places
the reference does: BertEmbeddings :111, JointEmbeddings MMBertEmbedding.py:70 (p = 0.5), attention probabilities
modeling_bert.py:131, BertSelfOutput :296, BertOutput :354.
"""
import math

import torch
import torch.nn.functional as F


class Cfg:
    """The few BertConfig fields the path reads."""

    def __init__(self, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                 vocab_size=30522, max_position_embeddings=512, layer_norm_eps=1e-12, num_labels=7):
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.vocab_size = vocab_size
        self.max_position_embeddings = max_position_embeddings
        self.layer_norm_eps = layer_norm_eps
        self.num_labels = num_labels


class Dropout:
    """Dropout probabilities + a torch generator (training-mode statistics; see the module docstring)."""

    def __init__(self, hidden=0.1, attn=0.1, joint=0.5, seed=0):
        self.hidden, self.attn, self.joint = hidden, attn, joint
        self.gen = torch.Generator().manual_seed(seed)

    def __call__(self, x, p):
        if p <= 0.0:
            return x
        keep = (torch.rand(x.shape, generator=self.gen) >= p).to(x.dtype)
        return x * keep / (1.0 - p)


def _ln(x, w, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _lin(x, sd, name):
    return x @ sd[name + ".weight"].t() + sd[name + ".bias"]


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _ext_mask(m, dtype):
    # MMBertForPretraining.py:74-77 (3-D frame mask -> feature 0 only), :152-153
    if m.dim() == 3:
        m = m[:, :, 0]
    return (1.0 - m.to(dtype))[:, None, None, :] * -10000.0


def _cross_entropy(logits, labels):
    # torch.nn.CrossEntropyLoss(): ignore_index=-100, mean over non-ignored rows (NaN when there are none)
    valid = labels != -100
    lse = torch.logsumexp(logits, dim=-1)
    picked = logits.gather(-1, labels.clamp(min=0)[:, None])[:, 0]
    return ((lse - picked) * valid).sum() / valid.sum()


def bert_pass(sd, cfg, ids, mask, token_type, frames=None, frame_mask=None, dtype=torch.float64, dropout=None):
    """MMBertModel.forward for one pass -> (sequence_output [B,S,H], pooled [B,H])."""
    H, nh = cfg.hidden_size, cfg.num_attention_heads
    d = H // nh
    B, T = ids.shape
    joint = frames is not None
    if joint:
        token_type = torch.zeros_like(ids)          # :223
    # nn.Embedding(vocab, H, padding_idx=config.pad_token_id == 0): lookups of id 0 contribute NO gradient to
    # row 0 (modeling_bert.py:58); the tied decoder's gradient into row 0 is kept.
    e = F.embedding(ids.long(), sd["bert.embeddings.word_embeddings.weight"], padding_idx=0) \
        + sd["bert.embeddings.token_type_embeddings.weight"][token_type.long()] \
        + sd["bert.embeddings.position_embeddings.weight"][:T][None]
    x = _ln(e, sd["bert.embeddings.LayerNorm.weight"], sd["bert.embeddings.LayerNorm.bias"], cfg.layer_norm_eps)
    drop = dropout if dropout is not None else (lambda t, p: t)
    p_h, p_a, p_j = (dropout.hidden, dropout.attn, dropout.joint) if dropout is not None else (0.0, 0.0, 0.0)
    x = drop(x, p_h)
    ext = _ext_mask(mask, dtype)
    if joint:
        which = "Wv" if frames.shape[-1] == sd["bert.jointEmbeddings.Wv.weight"].shape[1] else "Ws"
        f = frames.float().to(dtype)                 # MMBertEmbedding.py:62: pair_ids.float()
        p = torch.relu(_lin(f, sd, "bert.jointEmbeddings." + which))
        x = torch.cat((x, p), dim=1)
        x = drop(_ln(x, sd["bert.jointEmbeddings.LayerNorm.weight"], sd["bert.jointEmbeddings.LayerNorm.bias"], 1e-5), p_j)
        ext = torch.cat((ext, _ext_mask(frame_mask, dtype)), dim=-1)
    S = x.shape[1]
    for i in range(cfg.num_hidden_layers):
        pre = f"bert.encoder.layer.{i}."
        q = _lin(x, sd, pre + "attention.self.query").view(B, S, nh, d).transpose(1, 2)
        k = _lin(x, sd, pre + "attention.self.key").view(B, S, nh, d).transpose(1, 2)
        v = _lin(x, sd, pre + "attention.self.value").view(B, S, nh, d).transpose(1, 2)
        s = (q @ k.transpose(2, 3)) * (d ** -0.5) + ext
        ctx = (drop(torch.softmax(s, dim=-1), p_a) @ v).transpose(1, 2).reshape(B, S, H)
        a = _ln(drop(_lin(ctx, sd, pre + "attention.output.dense"), p_h) + x,
                sd[pre + "attention.output.LayerNorm.weight"], sd[pre + "attention.output.LayerNorm.bias"],
                cfg.layer_norm_eps)
        h = _gelu(_lin(a, sd, pre + "intermediate.dense"))
        x = _ln(drop(_lin(h, sd, pre + "output.dense"), p_h) + a,
                sd[pre + "output.LayerNorm.weight"], sd[pre + "output.LayerNorm.bias"], cfg.layer_norm_eps)
    pooled = torch.tanh(_lin(x[:, 0], sd, "bert.pooler.dense"))
    return x, pooled


def lm_head(sd, cfg, seq):
    t = _gelu(_lin(seq, sd, "cls.predictions.transform.dense"))
    t = _ln(t, sd["cls.predictions.transform.LayerNorm.weight"], sd["cls.predictions.transform.LayerNorm.bias"],
            cfg.layer_norm_eps)
    return t @ sd["bert.embeddings.word_embeddings.weight"].t() + sd["cls.predictions.bias"]


def cpc(sd, name, x, y):
    xp = _lin(y, sd, name + ".net")
    xp = xp / xp.norm(dim=1, keepdim=True)
    x = x / x.norm(dim=1, keepdim=True)
    pos = (x * xp).sum(-1)
    neg = torch.logsumexp(x @ xp.t(), dim=-1)
    return -(pos - neg).mean()


def heads(sd, cfg, x0, ap_v, ap_s, sentiment, dtype=torch.float64):
    """Everything the reference computes from the three [CLS] rows of a sample: pooler (modeling_bert.py:462-468),
    align / seq_relationship (MMBertForPretraining.py:295-302), fusion head (:406-415), CPC x3 (:422-425) and the
    alignment / label losses (:386-388, :427-443).  ``x0`` = [3B, H]: sequence_output[:, 0] of the text, text+visual and
    text+speech passes.  Returns (ap_loss, label_loss, nce, out_logits, rel_t, al_v, al_s)."""
    B = x0.shape[0] // 3
    pooled = torch.tanh(_lin(x0, sd, "bert.pooler.dense"))
    p_t, p_v, p_s = pooled[:B], pooled[B:2 * B], pooled[2 * B:]
    rel_t = _lin(p_t, sd, "cls.seq_relationship")
    al_v = _lin(x0[B:2 * B], sd, "cls.align")
    al_s = _lin(x0[2 * B:], sd, "cls.align")
    ap = (_cross_entropy(al_v, ap_v.reshape(-1)) + _cross_entropy(al_s, ap_s.reshape(-1))) / 2.0

    def score(p, vname):
        return _lin(torch.relu(_lin(torch.cat((p, p), dim=1), sd, "attn")), sd, vname)

    pooled3 = torch.cat((p_t * score(p_t, "vt"), p_v * score(p_v, "vv"), p_s * score(p_s, "vs")), dim=1)
    temp = _lin(pooled3, sd, "classifier1_1")
    logits = _lin(temp, sd, "classifier1_2")
    nce = cpc(sd, "cpc_zt", p_t, temp) + cpc(sd, "cpc_zv", p_v, temp) + cpc(sd, "cpc_za", p_s, temp)
    out_logits = logits
    if cfg.num_labels in (1, 7):                     # :431-436 regression
        if cfg.num_labels == 1:
            out_logits = torch.tanh(logits)
        label = ((out_logits.reshape(-1) - sentiment.reshape(-1).to(dtype)) ** 2).mean()
    else:                                            # :437-442 classification on the [B, 1] classifier output
        label = F.cross_entropy(logits, sentiment)
        out_logits = torch.argmax(torch.sigmoid(logits), dim=1)
    return ap, label, nce, out_logits, rel_t, al_v, al_s


def forward(sd, cfg, input_ids, token_type_ids, attention_mask, masked_labels, ap_label, sentiment,
            alpha=1.0, beta=1.0, dtype=torch.float64, dropout=None, cls_values=None):
    """MMBertForPretraining.forward -> ((13-tuple), logits), same structure as the reference.
    ``cls_values`` ([3B, H], test aid): the heads are EVALUATED at these [CLS] rows (a candidate's own encoder output)
    while their gradient still flows into this restatement's encoder (straight-through).  The fusion head's
    relu(attn(.)) gates (MMBertForPretraining.py:407-409) then take the candidate's side of every near-zero
    pre-activation, so that a gate flipped by bf16 rounding — a discontinuity of the model, not an error of the
    candidate — does not show up as a gradient difference in every encoder parameter."""
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    ids_t, vis, aud, ids_v, ids_s = input_ids
    m_t, (m_tv, m_v), (m_ts, m_s) = attention_mask
    lab_t, lab_v, lab_s = masked_labels
    ap_v, ap_s = ap_label
    V = cfg.vocab_size

    seq_t, _ = bert_pass(sd, cfg, ids_t, m_t, token_type_ids[0], dtype=dtype, dropout=dropout)
    seq_v, _ = bert_pass(sd, cfg, ids_v, m_tv, None, vis, m_v, dtype=dtype, dropout=dropout)
    seq_s, _ = bert_pass(sd, cfg, ids_s, m_ts, None, aud, m_s, dtype=dtype, dropout=dropout)
    pred_t, pred_v, pred_s = lm_head(sd, cfg, seq_t), lm_head(sd, cfg, seq_v), lm_head(sd, cfg, seq_s)
    mlm = (_cross_entropy(pred_t.reshape(-1, V), lab_t.reshape(-1))
           + _cross_entropy(pred_v.reshape(-1, V), lab_v.reshape(-1))
           + _cross_entropy(pred_s.reshape(-1, V), lab_s.reshape(-1))) / 3.0
    x0 = torch.cat((seq_t[:, 0], seq_v[:, 0], seq_s[:, 0]), dim=0)
    if cls_values is not None:
        x0 = x0 + (cls_values.to(dtype) - x0).detach()
    ap, label, nce, out_logits, rel_t, al_v, al_s = heads(sd, cfg, x0, ap_v, ap_s, sentiment, dtype=dtype)
    joint = alpha * mlm + ap + label - beta * nce
    return (joint, None, None, None, ap, label, nce, pred_t, rel_t, pred_v, al_v, pred_s, al_s), out_logits


HEAD_PARAM_PREFIXES = ("attn.", "bert.pooler.", "vt.", "vv.", "vs.", "classifier1_", "cpc_z", "cls.align.")


def heads_backward(sd, cfg, x0, ap_label, sentiment, beta=1.0, dtype=torch.float64):
    """Gradients of the head parameters (and of ``x0``) of ``ap + label - beta * nce`` for GIVEN [CLS] rows ``x0``
    [3B, H]: lets a test feed the candidate's own encoder output, so that ReLU gates near zero cannot flip between the
    two sides.  Returns ((ap, label, nce, out_logits), grads) with grads keyed by parameter name plus 'x0'."""
    leaves = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in sd.items()
              if k.startswith(HEAD_PARAM_PREFIXES) and v.is_floating_point()}
    full = dict(leaves)
    for k in ("cls.seq_relationship.weight", "cls.seq_relationship.bias"):
        full[k] = sd[k].detach().to(dtype)
    x = x0.detach().to(dtype).clone().requires_grad_(True)
    ap, label, nce, out_logits, _, _, _ = heads(full, cfg, x, ap_label[0], ap_label[1], sentiment, dtype=dtype)
    (ap + label - beta * nce).backward()
    grads = {k: v.grad for k, v in leaves.items()}
    grads["x0"] = x.grad
    return (ap.detach(), label.detach(), nce.detach(), out_logits.detach()), grads


# parameters that the reference leaves with ``grad is None`` after backward
NO_GRAD_PARAMS = ("bert.jointEmbeddings.W_cv.weight", "bert.jointEmbeddings.W_cv.bias",
                  "bert.jointEmbeddings.W_cs.weight", "bert.jointEmbeddings.W_cs.bias",
                  "cls.seq_relationship.weight", "cls.seq_relationship.bias")
# state_dict aliases of tied parameters (canonical name first)
TIED = {"cls.predictions.decoder.weight": "bert.embeddings.word_embeddings.weight",
        "cls.predictions.decoder.bias": "cls.predictions.bias"}


def forward_backward(sd, cfg, batch, alpha=1.0, beta=1.0, dtype=torch.float64, dropout=None, cls_values=None):
    """Runs forward + ``joint_loss.backward()`` on leaf copies of ``sd``.
    Returns (outputs, logits, grads) where grads maps canonical parameter names to gradients
    (None for parameters the path does not touch)."""
    leaves = {}
    for k, v in sd.items():
        if k in TIED or not v.is_floating_point():
            continue
        leaves[k] = v.detach().to(dtype).clone().requires_grad_(True)
    full = dict(leaves)
    for alias, canon in TIED.items():
        full[alias] = leaves[canon]
    out, logits = forward(full, cfg, alpha=alpha, beta=beta, dtype=dtype, dropout=dropout, cls_values=cls_values, **batch)
    out[0].backward()
    grads = {k: (v.grad if v.grad is not None else None) for k, v in leaves.items()}
    return out, logits, grads


def mask_tokens_rules(orig_ids, new_ids, labels, special_ids, mask_id):
    """Checker for an MLM-masking implementation against the rules of model_utils.mask_tokens (model_utils.py:6-39):
    returns a dict of violation counts (all zero = conforming) and the observed rates.
      * labels == original id at selected positions, -100 elsewhere (:29)
      * special tokens are never selected (:16-22)
      * a position's id changes only if it is selected, and only to mask_id (:31-33); the rest keep their token
    """
    orig_ids, new_ids, labels = orig_ids.cpu(), new_ids.cpu(), labels.cpu()
    special = torch.zeros_like(orig_ids, dtype=torch.bool)
    for s_ in special_ids:
        special |= orig_ids == s_
    selected = labels != -100
    changed = new_ids != orig_ids
    n_cand = int((~special).sum())
    return {
        "label_mismatch": int((labels[selected] != orig_ids[selected]).sum()),
        "special_selected": int((selected & special).sum()),
        "changed_unselected": int((changed & ~selected).sum()),
        "changed_not_to_mask": int((changed & (new_ids != mask_id)).sum()),
        "select_rate": float(selected.sum()) / max(n_cand, 1),
        "replace_rate": float((selected & (new_ids == mask_id) & (orig_ids != mask_id)).sum()) / max(int(selected.sum()), 1),
    }

