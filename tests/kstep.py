"""K-step training comparison (BASELINE.json: "bf16 path within ... 1e-2 absolute on sentiment MAE after a fixed number
of steps"): the CUDA model trained through msa_b200.trainer_fast.train_epoch — the reference's loop incl. its
``(step + 1) & accumulation`` stepping rule (trainer.py:96) — with the fused AdamW and the reference's dropout rates,
against the CPU oracle trained by a restatement of the same loop (oracle dropout = torch RNG, HF AdamW rule in torch).
Dropout streams cannot match, so the comparison is statistical: several dropout seeds per side, eval-mode MAE at the end.

Used by tests/test_model_gpu.py (small) and scripts/kstep_mae.py (bert-base width, recorded under profiles/)."""
import types

import torch

from msa_b200 import synth
from msa_b200.params import seeded_state_dict
from oracle import mmbert_oracle as O


def make_batches(dataset, n, B, T, L, vocab, seed=500):
    dv, da = synth.DATASET_DIMS[dataset]
    return [synth.make_batch(B, T, L, L, dv, da, vocab_size=vocab, seed=seed + i, min_len=5, mlm=False) for i in range(n)]


def to_collate(b):
    """forward kwargs -> the tuple structure of model_utils.collate (:117-142) that trainer.train_epoch unpacks."""
    ids_t, vis, aud, ids_v, ids_s = b["input_ids"]
    m_t, (m_tv, m_v), (m_ts, m_s) = b["attention_mask"]
    tt = b["token_type_ids"]
    text = (ids_t, None, tt[0], m_t, b["sentiment"])
    visual = (ids_v, vis, b["ap_label"][0], tt[1], m_v, None)
    speech = (ids_s, aud, b["ap_label"][1], tt[2], m_s, None)
    return text, visual, speech, (m_tv, m_ts), None, None


def from_ids(b):
    """mlm off (trainer.py:45-47 else branch): labels are the ids themselves, duplicated for the joint passes (:50,:53)."""
    ids_t, _, _, ids_v, ids_s = b["input_ids"]
    out = dict(b)
    out["masked_labels"] = (ids_t, torch.cat((ids_v, ids_v), -1), torch.cat((ids_s, ids_s), -1))
    return out


def mae(logits, sentiment):
    return float((logits.reshape(-1).double().cpu() - sentiment.double()).abs().mean())


def train_cuda(ocfg, dataset, sd, batches, eval_batches, order_seed, drop_seed, lr, p=(0.1, 0.1, 0.5)):
    from msa_b200 import trainer_fast
    from msa_b200.api import MMBertForPretraining
    from msa_b200.optim import FusedAdamW
    from msa_b200.params import BertShape
    shape = BertShape(ocfg.hidden_size, ocfg.num_hidden_layers, ocfg.num_attention_heads, ocfg.intermediate_size,
                      ocfg.vocab_size, ocfg.max_position_embeddings, hidden_dropout_prob=p[0], attention_probs_dropout_prob=p[1])
    m = MMBertForPretraining(shape)
    m.bert.set_joint_embeddings(dataset)
    m.bert.jointEmbeddings.dropout.p = p[2]
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    opt = FusedAdamW(m, lr=lr)
    sched = types.SimpleNamespace(step=lambda: None)
    args = types.SimpleNamespace(train_batch_size=1, gradient_accumulation_step=1, mlm=False, mlm_probability=0.15)
    data = [to_collate(b) for b in batches]
    order = torch.randperm(len(data), generator=torch.Generator().manual_seed(order_seed)).tolist()
    torch.manual_seed(drop_seed)                       # forward() draws its dropout seeds from torch's generator
    trainer_fast.train_epoch(args, m, [data[i] for i in order], opt, sched, None, collate_fn=lambda items: items[0],
                             device=torch.device("cuda"), faithful_stepping=True, sampler_shuffle=False)
    m.eval()
    out = []
    with torch.no_grad():
        for b in eval_batches:
            _, logits = m(**synth.tree_to(from_ids(b), "cuda"))
            out.append(mae(logits, b["sentiment"]))
    return sum(out) / len(out), order


def train_oracle(ocfg, sd, batches, eval_batches, order, drop_seed, lr, p=(0.1, 0.1, 0.5), wd=0.01, dtype=torch.float64):
    b1, b2, eps = 0.9, 0.999, 1e-6
    params = {k: v.to(dtype).clone() for k, v in sd.items() if k not in O.TIED}
    mom = {k: torch.zeros_like(v) for k, v in params.items()}
    var = {k: torch.zeros_like(v) for k, v in params.items()}
    acc = {k: None for k in params}
    drop = O.Dropout(p[0], p[1], p[2], seed=drop_seed) if max(p) > 0 else None
    t = 0

    def full():
        f = dict(params)
        for alias, canon in O.TIED.items():
            f[alias] = params[canon]
        return f

    for step, i in enumerate(order):
        _, _, grads = O.forward_backward(full(), ocfg, from_ids(batches[i]), dtype=dtype, dropout=drop)
        for k, g in grads.items():
            if g is not None:
                acc[k] = g if acc[k] is None else acc[k] + g           # loss.backward() accumulates (trainer.py:83)
        if ((step + 1) & 1) == 0:                                       # trainer.py:96 with accumulation 1
            t += 1
            s = lr * (1 - b2 ** t) ** 0.5 / (1 - b1 ** t)
            for k, g in acc.items():
                if g is None:
                    continue
                mom[k] = mom[k] * b1 + (1 - b1) * g
                var[k] = var[k] * b2 + (1 - b2) * g * g
                params[k] = params[k] - s * mom[k] / (var[k].sqrt() + eps)
                if not ("bias" in k or "LayerNorm.weight" in k):
                    params[k] = params[k] - lr * wd * params[k]
            acc = {k: None for k in params}                              # optimizer.zero_grad()
    out = []
    with torch.no_grad():
        f = full()
        for b in eval_batches:
            _, logits = O.forward(f, ocfg, dtype=dtype, **from_ids(b))
            out.append(mae(logits, b["sentiment"]))
    return sum(out) / len(out)


def run(ocfg, dataset, K, B, T, L, seeds, lr=5e-4, weight_seed=31, p=(0.1, 0.1, 0.5), dtype=torch.float64):
    sd = seeded_state_dict(ocfg, dataset, seed=weight_seed, std=0.02)
    batches = make_batches(dataset, K, B, T, L, ocfg.vocab_size)
    eval_batches = batches[:2] + make_batches(dataset, 2, B, T, L, ocfg.vocab_size, seed=900)
    with torch.no_grad():
        mae0 = sum(mae(O.forward(sd, ocfg, **from_ids(b))[1], b["sentiment"]) for b in eval_batches) / len(eval_batches)
    gpu, ref = [], []
    for s in range(seeds):
        mg, order = train_cuda(ocfg, dataset, sd, batches, eval_batches, order_seed=77, drop_seed=1000 + s, lr=lr, p=p)
        gpu.append(mg)
        ref.append(train_oracle(ocfg, sd, batches, eval_batches, order, drop_seed=2000 + s, lr=lr, p=p, dtype=dtype))
    tg, tr = torch.tensor(gpu, dtype=torch.float64), torch.tensor(ref, dtype=torch.float64)
    return dict(mae_initial=mae0, mae_cuda=gpu, mae_oracle=ref, mean_cuda=float(tg.mean()), mean_oracle=float(tr.mean()),
                gap=abs(float(tg.mean()) - float(tr.mean())),
                std_cuda=float(tg.std()) if seeds > 1 else 0.0, std_oracle=float(tr.std()) if seeds > 1 else 0.0,
                K=K, optimizer_steps=K // 2, lr=lr, dropout=list(p), seeds=seeds,
                shape=dict(hidden=ocfg.hidden_size, layers=ocfg.num_hidden_layers, vocab=ocfg.vocab_size, B=B, T=T, L=L))
