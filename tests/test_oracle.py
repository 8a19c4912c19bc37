"""Pins the CPU oracle (oracle/mmbert_oracle.py) against the reference: the committed golden vectors
(generated from the unmodified reference by tests/golden/make_golden.py) and, when /root/reference is
present, the live reference itself."""
import pytest
import torch

from oracle import mmbert_oracle as O
from oracle import ref_loader
from tests.helpers import GOLDEN, GOLDEN_ORACLE_ONLY, OUT_NAMES, expand_recipe, load_golden, rel_err

TOL = 2e-5   # fp64 oracle vs fp32 reference: fp32 rounding of the reference is the only difference


@pytest.mark.parametrize("name", GOLDEN + GOLDEN_ORACLE_ONLY)
def test_oracle_forward_matches_golden(name):
    recipe, g = load_golden(name)
    cfg, sd, batch = expand_recipe(recipe)
    with torch.no_grad():
        out, logits = O.forward(sd, cfg, alpha=recipe["alpha"], beta=recipe["beta"], **batch)
    for n, o in zip(OUT_NAMES, out):
        if n is None:
            assert o is None
            continue
        assert tuple(o.shape) == tuple(g["eval." + n].shape), n
        # scalar losses: absolute floor (with one sample the CPC terms are exactly 0 and the reference holds -4e-9)
        assert rel_err(o, g["eval." + n], floor=1e-3 if o.dim() == 0 else 1e-30) < TOL, n
    assert rel_err(logits, g["eval.logits"]) < TOL


@pytest.mark.parametrize("name", GOLDEN + GOLDEN_ORACLE_ONLY)
def test_oracle_backward_matches_golden(name):
    recipe, g = load_golden(name)
    cfg, sd, batch = expand_recipe(recipe)
    out, _, grads = O.forward_backward(sd, cfg, batch, alpha=recipe["alpha"], beta=recipe["beta"])
    assert rel_err(out[0].detach(), g["train.joint_loss"]) < TOL
    none = sorted(k for k, v in grads.items() if v is None)
    assert none == sorted(recipe["none_grads"]) == sorted(O.NO_GRAD_PARAMS)
    for k, v in grads.items():
        if v is None:
            continue
        # fp32 reference gradients carry ~1e-5 rounding noise; key-bias gradients are identically zero in exact
        # arithmetic (softmax shift invariance), hence the absolute floor
        assert rel_err(v, g["grad." + k], floor=1e-5) < 3e-4, k


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
def test_oracle_matches_live_reference_bert_base_shape():
    """bert-base width (hidden 768, 12 heads) at 2 layers, MOSI dims, fp32 reference vs fp64 oracle."""
    from transformers import BertConfig
    from msa_b200 import synth
    from msa_b200.params import seeded_state_dict
    ckw = dict(hidden_size=768, num_hidden_layers=2, num_attention_heads=12, intermediate_size=3072,
               vocab_size=2048, max_position_embeddings=64)
    cfg = BertConfig(**ckw)
    model = ref_loader.build_model(cfg, "mosi").eval()
    sd = seeded_state_dict(cfg, "mosi", seed=3)
    model.load_state_dict(sd, strict=False)
    batch = synth.make_batch(2, 12, 12, 12, 47, 74, vocab_size=2048, seed=5, min_len=4)
    with torch.no_grad():
        ref_out, ref_logits = model(**batch)
        out, logits = O.forward(sd, O.Cfg(**ckw), **batch)
    for a, b in zip(out, ref_out):
        if b is None:
            assert a is None
        else:
            assert rel_err(a, b) < TOL
    assert rel_err(logits, ref_logits) < TOL


def test_fused_torch_baseline_equals_the_oracle():
    """oracle/torch_baseline.py (the port bench.py times on the GPU as the PyTorch / cuBLASLt / SDPA bar) is the same
    arithmetic as the pinned oracle: fp64, eval mode, every output to 1e-12."""
    from msa_b200 import synth
    from msa_b200.params import seeded_state_dict
    from oracle import torch_baseline as TB
    cfg = O.Cfg(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, vocab_size=512,
                max_position_embeddings=64)
    sd = seeded_state_dict(cfg, "mosei", seed=3, std=0.05)
    batch = synth.make_batch(3, 10, 24, 17, 35, 74, vocab_size=512, seed=5, min_len=4)
    with torch.no_grad():
        o1, l1 = O.forward(sd, cfg, alpha=0.7, beta=0.3, **batch)
        o2, l2 = TB.forward({k: v.double() for k, v in sd.items()}, cfg, alpha=0.7, beta=0.3, **batch)
    for a, b in zip(o1, o2):
        assert (a is None) == (b is None)
        if a is not None:
            assert rel_err(b, a) < 1e-12
    assert rel_err(l2, l1) < 1e-12
