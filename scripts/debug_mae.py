"""Bring-up aid: K-step training of the CUDA path in lockstep with the fp64 oracle (per-step MAE, parameter drift)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msa_b200 import synth
from msa_b200.optim import FusedAdamW
from msa_b200.params import seeded_state_dict
from oracle import mmbert_oracle as O
from tests.test_model_gpu import _build

ocfg = O.Cfg(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, vocab_size=512,
             max_position_embeddings=64)
sd = seeded_state_dict(ocfg, "mosi", seed=31, std=0.03)
batch = synth.make_batch(6, 12, 12, 12, 47, 74, vocab_size=512, seed=13, min_len=5)
K, lr, wd, b1, b2, eps = int(sys.argv[1]), float(sys.argv[2]), 0.01, 0.9, 0.999, 1e-6
m = _build(ocfg, "mosi", sd).train()
opt = FusedAdamW(m, lr=lr, weight_decay=wd)
dbatch = synth.tree_to(batch, "cuda")
params = {k: v.double().clone() for k, v in sd.items() if k not in O.TIED}
mom = {k: torch.zeros_like(v) for k, v in params.items()}
var = {k: torch.zeros_like(v) for k, v in params.items()}
sent = batch["sentiment"].double()
for t in range(1, K + 1):
    out, logits = m(**dbatch)
    out[0].mean().backward()
    full = dict(params)
    for alias, canon in O.TIED.items():
        full[alias] = params[canon]
    ro, rl, grads = O.forward_backward(full, ocfg, batch)
    mae_g = float((logits.view(-1).double().cpu() - sent).abs().mean())
    mae_r = float((rl.view(-1).detach() - sent).abs().mean())
    # gradient agreement at this step (parameters differ slightly from step 2 on)
    worst = []
    named = dict(m.named_parameters())
    for k, g in grads.items():
        if g is None or named[k].grad is None:
            continue
        gg = named[k].grad.double().cpu()
        worst.append((float((gg - g).abs().max() / g.abs().max().clamp_min(1e-9)), k))
    worst.sort(reverse=True)
    print(f"step {t}: loss gpu {float(out[0]):.5f} ref {float(ro[0]):.5f}  mae gpu {mae_g:.5f} ref {mae_r:.5f}  worst grads",
          [(f"{e:.2e}", k) for e, k in worst[:4]])
    opt.step(); opt.zero_grad()
    step = lr * (1 - b2 ** t) ** 0.5 / (1 - b1 ** t)
    for k, g in grads.items():
        if g is None:
            continue
        mom[k] = mom[k] * b1 + (1 - b1) * g
        var[k] = var[k] * b2 + (1 - b2) * g * g
        params[k] = params[k] - step * mom[k] / (var[k].sqrt() + eps)
        if not ("bias" in k or "LayerNorm.weight" in k):
            params[k] = params[k] - lr * wd * params[k]
    drift = []
    for k, p in m.named_parameters():
        if k in params:
            d = float((p.detach().double().cpu() - params[k]).abs().max())
            drift.append((d, k))
    drift.sort(reverse=True)
    print("   param drift", [(f"{d:.2e}", k) for d, k in drift[:5]])
m.eval()
with torch.no_grad():
    _, le = m(**dbatch)
m.train()
_, lt = m(**dbatch)
print("eval vs train logits", le.view(-1).tolist(), lt.view(-1).tolist())
