"""Shared test helpers: golden fixture loading and the recipe -> (cfg, state_dict, batch) expansion."""
import json
import os

import numpy as np
import torch

from msa_b200 import synth
from msa_b200.params import seeded_state_dict
from oracle import mmbert_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# four fixtures generated from the unmodified reference (tests/golden/make_golden.py): MOSI-aligned, MOSEI-unaligned,
# UR-FUNNY's wide frame dims (371 / 81) and a one-sample batch (the in-batch CPC terms degenerate to 0).  All four pin
# the CPU oracle AND are GPU parity cases.
GOLDEN = ("tiny_mosi_aligned", "tiny_mosei_unaligned", "tiny_ur_funny", "tiny_mosi_single")
GOLDEN_ORACLE_ONLY = ()
OUT_NAMES = ("joint_loss", None, None, None, "ap_loss", "label_loss", "nce", "pred_t", "rel_t", "pred_v",
             "align_v", "pred_s", "align_s")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    recipe = json.loads(bytes(z["recipe"]).decode())
    arrays = {k: torch.from_numpy(z[k]) for k in z.files if k != "recipe"}
    return recipe, arrays


def expand_recipe(recipe):
    cfg = O.Cfg(**recipe["cfg"])
    sd = seeded_state_dict(cfg, recipe["dataset"], seed=recipe["weight_seed"])
    dv, da = synth.DATASET_DIMS[recipe["dataset"]]
    batch = synth.make_batch(recipe["B"], recipe["T"], recipe["Lv"], recipe["La"], dv, da,
                             vocab_size=cfg.vocab_size, seed=recipe["data_seed"], min_len=recipe["min_len"])
    return cfg, sd, batch


def rel_err(a, b, floor=1e-30):
    """max |a-b| / max(max|b|, floor): the 'relative on logits / loss' measure used by every parity test."""
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(floor))
