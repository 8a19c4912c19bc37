"""Stage-by-stage comparison of the CUDA path with the oracle (embeddings, every layer, heads) for bring-up."""
import copy
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msa_b200 import synth
from msa_b200.params import seeded_state_dict
from oracle import mmbert_oracle as O
from tests.helpers import rel_err, load_golden, expand_recipe
from tests.test_model_gpu import _build

name = sys.argv[1] if len(sys.argv) > 1 else "tiny_mosi_aligned"
recipe, g = load_golden(name)
ocfg, sd, batch = expand_recipe(recipe)
m = _build(ocfg, recipe["dataset"], sd).train()
m.set_alpha_beta(recipe["alpha"], recipe["beta"])
out, logits = m(**synth.tree_to(batch, "cuda"))
torch.cuda.synchronize()
plan = next(iter(m._plans.values()))
B, T, Lv, La = plan.B, plan.T, plan.Lv, plan.La
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
ids_t, vis, aud, ids_v, ids_s = batch["input_ids"]
m_t, (m_tv, m_v), (m_ts, m_s) = batch["attention_mask"]
b1, b2 = B * T, B * T + B * (T + Lv)
for k in range(ocfg.num_hidden_layers + 1):
    c = copy.copy(ocfg); c.num_hidden_layers = k
    xs = [O.bert_pass(sd64, c, ids_t, m_t, batch["token_type_ids"][0])[0],
          O.bert_pass(sd64, c, ids_v, m_tv, None, vis, m_v)[0],
          O.bert_pass(sd64, c, ids_s, m_ts, None, aud, m_s)[0]]
    ref = torch.cat([x.reshape(-1, ocfg.hidden_size) for x in xs])
    got = plan.x[k].float().cpu()
    print(f"x[{k}] rel_err {rel_err(got, ref):.4e}  per-pass",
          [f"{rel_err(got[a:b], ref[a:b]):.3e}" for a, b in ((0, b1), (b1, b2), (b2, plan.M))])
kb = plan.keybias.cpu()
ref_kb = torch.cat([((1 - m_t) * -10000).reshape(-1),
                    torch.cat(((1 - m_tv) * -10000, (1 - m_v[:, :, 0]) * -10000), 1).reshape(-1),
                    torch.cat(((1 - m_ts.double()) * -10000, (1 - m_s[:, :, 0].double()) * -10000), 1).reshape(-1)])
print("keybias equal:", bool((kb.double() == ref_kb).all()), "label_count", plan.label_count.cpu().tolist(),
      "cu", plan.cu.cpu().tolist()[:5])
ref_out, ref_logits = O.forward(sd, ocfg, alpha=recipe["alpha"], beta=recipe["beta"], **batch)
names = ("joint", None, None, None, "ap", "label", "nce", "pred_t", "rel_t", "pred_v", "align_v", "pred_s", "align_s")
for n, a, b in zip(names, out, ref_out):
    if n: print(f"{n:8s} rel_err {rel_err(a.detach().float(), b.detach()):.4e}")
print("logits rel_err", rel_err(logits.float(), ref_logits.detach()), "losses", plan.losses.cpu().tolist())
out[0].backward(); torch.cuda.synchronize()
_, _, ref_grads = O.forward_backward(sd, ocfg, batch, alpha=recipe["alpha"], beta=recipe["beta"])
rows = []
for n, p in m.named_parameters():
    if p.grad is None: rows.append((n, None)); continue
    rows.append((n, rel_err(p.grad, ref_grads[n], floor=1e-4)))
for n, e in rows: print(f"grad {n:60s} {'None' if e is None else f'{e:.3e}'}")
