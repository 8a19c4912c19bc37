"""Seed-to-seed statistics of the training-mode loss and of a few gradients, CUDA path vs the CPU oracle (torch Philox
dropout), same weights and batch: the two dropout implementations must give the same mean AND the same spread.
usage: python scripts/dropout_variance.py [nseeds]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msa_b200 import synth
from msa_b200.api import MMBertForPretraining
from msa_b200.params import BertShape, seeded_state_dict
from oracle import mmbert_oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
P = tuple(float(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (0.1, 0.1, 0.5)     # hidden, attention, joint
ocfg = O.Cfg(num_hidden_layers=2)
sd = seeded_state_dict(ocfg, "mosi", seed=31, std=0.02)
batch = synth.make_batch(6, 16, 16, 16, 47, 74, vocab_size=ocfg.vocab_size, seed=500, min_len=5)
names = ("classifier1_2.weight", "bert.encoder.layer.0.intermediate.dense.weight", "bert.encoder.layer.1.attention.self.query.weight",
         "bert.jointEmbeddings.Wv.weight", "bert.embeddings.position_embeddings.weight")


def stats(xs):
    t = torch.tensor(xs, dtype=torch.float64)
    return float(t.mean()), float(t.std())


shape = BertShape(768, 2, 12, 3072, ocfg.vocab_size, 512, hidden_dropout_prob=P[0], attention_probs_dropout_prob=P[1])
m = MMBertForPretraining(shape)
m.bert.set_joint_embeddings("mosi")
m.bert.jointEmbeddings.dropout.p = P[2]
m.load_state_dict(sd, strict=True)
m = m.cuda().train()
db = synth.tree_to(batch, "cuda")
res = {"cuda": {"loss": [], **{k: [] for k in names}}, "oracle": {"loss": [], **{k: [] for k in names}}}
for s in range(n):
    torch.manual_seed(1000 + s)
    for p in m.parameters():
        p.grad = None
    out, _ = m(**db)
    out[0].backward()
    res["cuda"]["loss"].append(float(out[0]))
    g = dict(m.named_parameters())
    for k in names:
        res["cuda"][k].append(float(g[k].grad.double().norm()))
    drop = O.Dropout(P[0], P[1], P[2], seed=2000 + s)
    o, _, grads = O.forward_backward(sd, ocfg, batch, dtype=torch.float32, dropout=drop)
    res["oracle"]["loss"].append(float(o[0]))
    for k in names:
        res["oracle"][k].append(float(grads[k].double().norm()))
print(f"# dropout (hidden, attention, joint) = {P}, {n} seeds")
for key in ("loss",) + names:
    mc, sc = stats(res["cuda"][key])
    mo, so = stats(res["oracle"][key])
    print(f"{key:55s} cuda mean {mc:10.5f} std {sc:9.5f} | oracle mean {mo:10.5f} std {so:9.5f} | std ratio {sc / max(so, 1e-12):6.2f}")
