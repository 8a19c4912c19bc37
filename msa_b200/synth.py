"""Seeded synthetic batches with the exact tuple layout, dtypes and statistics that the reference's
``trainer.train_epoch`` feeds to ``MMBertForPretraining.forward`` (trainer.py:45-72, model_utils.py:15-32,
117-142, train.py:101-133, MMBertDataset.py:89-92,145-156; summarised in SURVEY.md §8d).

Host-side logic only (CPU tensors); used by the tests, the bench and the smoke run.
"""
from dataclasses import dataclass

import torch

# dataset -> (visual dim, speech dim)   (reference config.py:12-17)
DATASET_DIMS = {"mosi": (47, 74), "mosei": (35, 74), "ur_funny": (371, 81)}
PAD, CLS, SEP, MASK = 0, 101, 102, 103


@dataclass
class Workload:
    """A named shape of BASELINE.json's configs."""
    name: str
    dataset: str
    T: int
    Lv: int
    La: int
    batch: int

    @property
    def dims(self):
        return DATASET_DIMS[self.dataset]

    @property
    def positions(self):  # encoder positions per sample: the reference's three passes
        return 3 * self.T + self.Lv + self.La


WORKLOADS = {
    # configs[0]/[1]: MOSI-aligned, text 50 + audio 50x74 + visual 50x47
    "mosi_aligned_b32": Workload("mosi_aligned_b32", "mosi", 50, 50, 50, 32),
    "mosi_aligned_b64": Workload("mosi_aligned_b64", "mosi", 50, 50, 50, 64),
    # configs[2]: CMU-MOSEI unaligned, text 50 + audio 500x74 + visual 500x35
    "mosei_unaligned_b64": Workload("mosei_unaligned_b64", "mosei", 50, 500, 500, 64),
    # configs[3]: UR-FUNNY-shaped
    "ur_funny_b64": Workload("ur_funny_b64", "ur_funny", 50, 50, 50, 64),
}


def _mask_tokens(ids, gen, mlm_probability=0.15):
    """model_utils.mask_tokens (:15-32): 15 % of non-special tokens are selected, 80 % of those become
    [MASK]; labels hold the original id at selected positions and -100 elsewhere."""
    ids = ids.clone()
    labels = ids.clone()
    special = (ids == PAD) | (ids == CLS) | (ids == SEP)
    prob = torch.full(ids.shape, mlm_probability)
    prob.masked_fill_(special, 0.0)
    selected = torch.bernoulli(prob, generator=gen).bool()
    # the reference loss is NaN when a pass has no labelled position: guarantee one per batch
    if not selected.any():
        cand = (~special).nonzero()
        if len(cand):
            selected[cand[0, 0], cand[0, 1]] = True
    labels[~selected] = -100
    replaced = torch.bernoulli(torch.full(ids.shape, 0.8), generator=gen).bool() & selected
    ids[replaced] = MASK
    return ids, labels


def make_batch(B, T, Lv, La, Dv, Da, vocab_size=30522, seed=1234, min_len=None, mlm=True):
    """Returns the keyword arguments of ``MMBertForPretraining.forward`` as CPU tensors."""
    g = torch.Generator().manual_seed(seed)
    min_len = min(10, T) if min_len is None else min_len
    lo = 1000 if vocab_size > 2000 else MASK + 1
    text = torch.randint(lo, vocab_size, (B, T), generator=g)
    lens = torch.randint(min_len, T + 1, (B,), generator=g)
    text[:, 0] = CLS
    for b in range(B):
        n = int(lens[b])
        text[b, n - 1] = SEP
        text[b, n:] = PAD
    vis = torch.randn(B, Lv, Dv, generator=g, dtype=torch.float64)
    aud = torch.randn(B, La, Da, generator=g, dtype=torch.float64)
    for frames, L in ((vis, Lv), (aud, La)):
        if L == T:  # aligned: word-level features, SEP row and padding rows are zero (train.py:115-127)
            flen = lens - 1
        else:       # unaligned: independent frame count
            flen = torch.randint(min(50, L), L + 1, (B,), generator=g)
        for b in range(B):
            frames[b, int(flen[b]):] = 0
    if mlm:
        ids_t, lab_t = _mask_tokens(text, g)
        ids_v, lab_v = _mask_tokens(text, g)
        ids_s, lab_s = _mask_tokens(text, g)
    else:
        ids_t = ids_v = ids_s = text
        lab_t = lab_v = lab_s = text
    neg_v = torch.full((B, Lv), -100, dtype=torch.int64)
    neg_s = torch.full((B, La), -100, dtype=torch.int64)
    # trainer.py:50,53 duplicates the text labels onto the frame half (only shape-valid when L == T)
    lab_v2 = torch.cat((lab_v, lab_v if Lv == T else neg_v), dim=-1)
    lab_s2 = torch.cat((lab_s, lab_s if La == T else neg_s), dim=-1)
    return dict(
        input_ids=(ids_t, vis, aud, ids_v, ids_s),
        token_type_ids=(torch.zeros(B, T, dtype=torch.int64), torch.zeros(B, T, dtype=torch.int64),
                        torch.zeros(B, T, dtype=torch.int64)),
        attention_mask=((text != PAD).double(),
                        (torch.ones(B, T, dtype=torch.float64), (vis != 0).double()),   # collate typo: all ones
                        (torch.ones(B, T, dtype=torch.int64), (aud != 0).long())),
        masked_labels=(lab_t, lab_v2, lab_s2),
        ap_label=(torch.randint(0, 2, (B,), generator=g), torch.randint(0, 2, (B,), generator=g)),
        sentiment=torch.empty(B).uniform_(-3, 3, generator=g),
    )


def make_workload_batch(w: Workload, seed=1234, batch=None, vocab_size=30522):
    dv, da = w.dims
    return make_batch(batch or w.batch, w.T, w.Lv, w.La, dv, da, vocab_size=vocab_size, seed=seed)


def tree_to(obj, device, non_blocking=False):
    """Moves a nested tuple/dict of tensors to a device (the ``.to(DEVICE)`` calls of trainer.py:49-72)."""
    if torch.is_tensor(obj):
        return obj.to(device, non_blocking=non_blocking)
    if isinstance(obj, tuple):
        return tuple(tree_to(o, device, non_blocking) for o in obj)
    if isinstance(obj, dict):
        return {k: tree_to(v, device, non_blocking) for k, v in obj.items()}
    return obj


def tree_map(fn, obj):
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, tuple):
        return tuple(tree_map(fn, o) for o in obj)
    if isinstance(obj, dict):
        return {k: tree_map(fn, v) for k, v in obj.items()}
    return obj


def tree_bytes(obj):
    if torch.is_tensor(obj):
        return obj.numel() * obj.element_size()
    if isinstance(obj, (tuple, list)):
        return sum(tree_bytes(o) for o in obj)
    if isinstance(obj, dict):
        return sum(tree_bytes(v) for v in obj.values())
    return 0
