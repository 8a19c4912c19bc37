"""Generates tests/golden/*.npz by running the UNMODIFIED reference (via oracle/ref_loader.py) on seeded
weights (msa_b200.params.seeded_state_dict) and seeded synthetic batches (msa_b200.synth.make_batch).
Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

Each fixture stores only the recipe (config, seeds, shapes) and the reference's results: the 13-tuple
outputs + logits in eval mode (fp32), and — from a train-mode run with every dropout probability set to
0 — the loss and every parameter gradient (fp32), plus the names of parameters whose grad is None.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from msa_b200 import synth  # noqa: E402
from msa_b200.params import seeded_state_dict, TIED  # noqa: E402

FIXTURES = {
    # name: (dataset, cfg kwargs, B, T, Lv, La, weight seed, data seed, alpha, beta)
    "tiny_mosi_aligned": ("mosi", dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                       intermediate_size=256, vocab_size=256, max_position_embeddings=32),
                          3, 8, 8, 8, 11, 21, 1.0, 1.0),
    "tiny_mosei_unaligned": ("mosei", dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                           intermediate_size=256, vocab_size=256, max_position_embeddings=32),
                             4, 10, 70, 23, 12, 22, 0.7, 0.3),
    # oracle-only fixtures (tests/helpers.py: GOLDEN_ORACLE_ONLY): the wide UR-FUNNY frame dims (371 / 81), and a
    # single-sample batch (CPC's in-batch negatives degenerate to one row; every sequence length at its minimum)
    "tiny_ur_funny": ("ur_funny", dict(hidden_size=128, num_hidden_layers=1, num_attention_heads=2,
                                       intermediate_size=256, vocab_size=256, max_position_embeddings=32),
                      3, 9, 9, 9, 13, 23, 1.0, 0.5),
    "tiny_mosi_single": ("mosi", dict(hidden_size=64, num_hidden_layers=1, num_attention_heads=1,
                                      intermediate_size=128, vocab_size=256, max_position_embeddings=32),
                         1, 6, 6, 6, 14, 24, 1.0, 1.0),
}
OUT_NAMES = ("joint_loss", None, None, None, "ap_loss", "label_loss", "nce", "pred_t", "rel_t", "pred_v",
             "align_v", "pred_s", "align_s")


def run(name):
    from transformers import BertConfig
    dataset, ckw, B, T, Lv, La, wseed, dseed, alpha, beta = FIXTURES[name]
    cfg = BertConfig(**ckw)
    cfg.hidden_dropout_prob = 0.0
    cfg.attention_probs_dropout_prob = 0.0
    model = ref_loader.build_model(cfg, dataset)
    sd = seeded_state_dict(cfg, dataset, seed=wseed)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in m or "token_type_ids" in m for m in missing), (missing, unexpected)
    model.set_alpha_beta(alpha, beta)
    model.bert.jointEmbeddings.dropout.p = 0.0
    dv, da = synth.DATASET_DIMS[dataset]
    batch = synth.make_batch(B, T, Lv, La, dv, da, vocab_size=cfg.vocab_size, seed=dseed, min_len=4)
    arrays = {}
    model.eval()
    with torch.no_grad():
        out, logits = model(**batch)
    for n, o in zip(OUT_NAMES, out):
        if n is None:
            assert o is None
        else:
            arrays["eval." + n] = o.detach().numpy()
    arrays["eval.logits"] = logits.detach().numpy()
    model.train()
    out, logits = model(**batch)
    out[0].mean().backward()
    arrays["train.joint_loss"] = out[0].detach().numpy()
    none_grads = []
    for n, p in model.named_parameters():
        if p.grad is None:
            none_grads.append(n)
        else:
            arrays["grad." + n] = p.grad.detach().numpy()
    recipe = dict(name=name, dataset=dataset, cfg=ckw, B=B, T=T, Lv=Lv, La=La, weight_seed=wseed, data_seed=dseed,
                  alpha=alpha, beta=beta, min_len=4, none_grads=none_grads, torch=torch.__version__)
    arrays["recipe"] = np.frombuffer(json.dumps(recipe).encode(), dtype=np.uint8)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz")
    np.savez_compressed(path, **arrays)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB; joint", float(out[0]), "none_grads", none_grads)


if __name__ == "__main__":
    for n in (sys.argv[1:] or FIXTURES):
        run(n)
