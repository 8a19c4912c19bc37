"""Parameter inventory of MMBertForPretraining: names, shapes and order exactly as the reference's
``named_parameters()`` / ``state_dict()`` expose them (MMBertForPretraining.py:304-347,
MMBertEmbedding.py:34-55, transformers modeling_bert.py BertEmbeddings/BertLayer/BertPooler/heads).

``train.py:76-91`` groups parameters for weight decay by substring of these names and checkpoints are
``state_dict()`` dumps (trainer.py:269), so the names are part of the drop-in boundary.
"""
from collections import OrderedDict

import numpy as np
import torch

from .synth import DATASET_DIMS

# state_dict aliases of tied parameters: alias -> canonical
TIED = OrderedDict([
    ("cls.predictions.decoder.weight", "bert.embeddings.word_embeddings.weight"),
    ("cls.predictions.decoder.bias", "cls.predictions.bias"),
])
# parameters the hot path never touches: they must finish backward with ``.grad is None``
NO_GRAD = ("bert.jointEmbeddings.W_cv.weight", "bert.jointEmbeddings.W_cv.bias",
           "bert.jointEmbeddings.W_cs.weight", "bert.jointEmbeddings.W_cs.bias",
           "cls.seq_relationship.weight", "cls.seq_relationship.bias")


def param_shapes(cfg, dataset):
    """OrderedDict name -> shape, in the reference's named_parameters() order."""
    H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    dv, da = DATASET_DIMS[dataset]
    s = OrderedDict()
    s["bert.embeddings.word_embeddings.weight"] = (V, H)
    s["bert.embeddings.position_embeddings.weight"] = (cfg.max_position_embeddings, H)
    s["bert.embeddings.token_type_embeddings.weight"] = (getattr(cfg, "type_vocab_size", 2), H)
    s["bert.embeddings.LayerNorm.weight"] = (H,)
    s["bert.embeddings.LayerNorm.bias"] = (H,)
    for i in range(cfg.num_hidden_layers):
        p = f"bert.encoder.layer.{i}."
        for n in ("query", "key", "value"):
            s[p + f"attention.self.{n}.weight"] = (H, H)
            s[p + f"attention.self.{n}.bias"] = (H,)
        s[p + "attention.output.dense.weight"] = (H, H)
        s[p + "attention.output.dense.bias"] = (H,)
        s[p + "attention.output.LayerNorm.weight"] = (H,)
        s[p + "attention.output.LayerNorm.bias"] = (H,)
        s[p + "intermediate.dense.weight"] = (I, H)
        s[p + "intermediate.dense.bias"] = (I,)
        s[p + "output.dense.weight"] = (H, I)
        s[p + "output.dense.bias"] = (H,)
        s[p + "output.LayerNorm.weight"] = (H,)
        s[p + "output.LayerNorm.bias"] = (H,)
    s["bert.pooler.dense.weight"] = (H, H)
    s["bert.pooler.dense.bias"] = (H,)
    s["bert.jointEmbeddings.W_cv.weight"] = (H, dv + H)
    s["bert.jointEmbeddings.W_cv.bias"] = (H,)
    s["bert.jointEmbeddings.W_cs.weight"] = (H, da + H)
    s["bert.jointEmbeddings.W_cs.bias"] = (H,)
    s["bert.jointEmbeddings.Wv.weight"] = (H, dv)
    s["bert.jointEmbeddings.Wv.bias"] = (H,)
    s["bert.jointEmbeddings.Ws.weight"] = (H, da)
    s["bert.jointEmbeddings.Ws.bias"] = (H,)
    s["bert.jointEmbeddings.LayerNorm.weight"] = (H,)
    s["bert.jointEmbeddings.LayerNorm.bias"] = (H,)
    s["cls.predictions.bias"] = (V,)
    s["cls.predictions.transform.dense.weight"] = (H, H)
    s["cls.predictions.transform.dense.bias"] = (H,)
    s["cls.predictions.transform.LayerNorm.weight"] = (H,)
    s["cls.predictions.transform.LayerNorm.bias"] = (H,)
    s["cls.seq_relationship.weight"] = (2, H)
    s["cls.seq_relationship.bias"] = (2,)
    s["cls.align.weight"] = (2, H)
    s["cls.align.bias"] = (2,)
    s["classifier1_1.weight"] = (H, 3 * H)
    s["classifier1_1.bias"] = (H,)
    s["classifier1_2.weight"] = (1, H)   # num_labels is 7 at construction -> Linear(H, 1) (:309-314)
    s["classifier1_2.bias"] = (1,)
    s["attn.weight"] = (H, 2 * H)
    s["attn.bias"] = (H,)
    for n in ("vt", "vs", "vv"):
        s[n + ".weight"] = (1, H)
        s[n + ".bias"] = (1,)
    for n in ("cpc_zt", "cpc_zv", "cpc_za"):
        s[n + ".net.weight"] = (H, H)
        s[n + ".net.bias"] = (H,)
    return s


def seeded_state_dict(cfg, dataset, seed=0, std=0.05):
    """Deterministic, platform-independent random weights (numpy PCG64): the golden fixtures store
    only the seed.  Linear/embedding weights ~ N(0, std), LayerNorm weights ~ 1 + 0.1 N(0,1), biases ~ 0.02 N(0,1).
    Returns a state_dict including the tied aliases."""
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    for name, shape in param_shapes(cfg, dataset).items():
        x = rng.standard_normal(shape).astype(np.float32)
        if "LayerNorm.weight" in name:
            x = 1.0 + 0.1 * x
        elif name.endswith("bias"):
            x = 0.02 * x
        else:
            x = std * x
        sd[name] = torch.from_numpy(x)
    for alias, canon in TIED.items():
        sd[alias] = sd[canon]
    return sd


class BertShape:
    """The BertConfig fields the path reads (defaults = bert-base-uncased, the reference's model).  Any object
    with these attributes — e.g. ``transformers.BertConfig`` — can be passed to MMBertForPretraining instead."""

    def __init__(self, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                 vocab_size=30522, max_position_embeddings=512, type_vocab_size=2, layer_norm_eps=1e-12,
                 hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, initializer_range=0.02):
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.vocab_size = vocab_size
        self.max_position_embeddings = max_position_embeddings
        self.type_vocab_size = type_vocab_size
        self.layer_norm_eps = layer_norm_eps
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.initializer_range = initializer_range


def forward_gflop_per_sample(shape, T, Lv, La, Dv, Da):
    """Algorithmic forward GFLOP of ONE sample (= the reference's three passes), dense-faithful: SURVEY.md §8d."""
    H, I, V, N = shape.hidden_size, shape.intermediate_size, shape.vocab_size, shape.num_hidden_layers
    passes = (T, T + Lv, T + La)
    tok = sum(passes)
    enc = N * tok * (8 * H * H + 4 * H * I)
    attn = N * sum(4 * S * S * H for S in passes)
    mlm = tok * (2 * H * H + 2 * H * V)
    proj = 2 * H * (Lv * Dv + La * Da)
    small = 30 * H * H + 20 * H
    return (enc + attn + mlm + proj + small) / 1e9


def train_gflop_per_sample(shape, workload):
    """train = 3 x forward (backward = 2 x forward, no recompute credit)."""
    dv, da = workload.dims
    return 3.0 * forward_gflop_per_sample(shape, workload.T, workload.Lv, workload.La, dv, da)
