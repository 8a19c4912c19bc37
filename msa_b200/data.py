"""On-device input preparation for the MMBert step: drop-ins for the pieces of the reference's
host loop that produce the model's inputs.

* ``mask_tokens(inputs, tokenizer, args)`` — same name, arguments and return value as ``model_utils.mask_tokens``
  (model_utils.py:6-39), but one kernel launch on the device tensor instead of Python lists, two ``torch.bernoulli`` calls
  and boolean-index writes with host round trips.  ``inputs`` is modified in place, like the reference.
* ``mask_step(text_ids, twv_ids, tws_ids, ...)`` — the three calls of trainer.py:45-47 plus the label duplication of
  :50,53 (``cat((labels, labels), -1)``), returning the ``masked_labels`` tuple the model takes.

The random stream is the library's counter-based generator (not torch's), seeded from ``torch`` so that
``torch.manual_seed`` still makes runs reproducible.
"""
import torch

from . import capi

BERT_SPECIAL_IDS = (0, 100, 101, 102, 103)       # [PAD] [UNK] [CLS] [SEP] [MASK]  (bert-base-uncased)


def _special_ids(tokenizer):
    ids = getattr(tokenizer, "all_special_ids", None)
    return tuple(int(i) for i in ids) if ids else BERT_SPECIAL_IDS


def mask_ids_(ids, labels, *, prob=0.15, replace_prob=0.8, mask_id=103, special=BERT_SPECIAL_IDS, seed=None, rng_stream=0,
              labels_dup=None):
    """In-place masking of a CUDA int64 [B,T] tensor through mmb_mlm_mask; fills ``labels`` (and ``labels_dup`` [B,2T])."""
    if ids.device.type != "cuda" or ids.dtype != torch.int64 or not ids.is_contiguous() or ids.dim() != 2:
        raise capi.MMBError("mask_tokens needs a contiguous CUDA int64 [B,T] tensor (there is no CPU path)")
    if len(special) > 8:
        raise capi.MMBError("at most 8 special token ids")
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    B, T = ids.shape
    a = capi.fill(capi.MlmMaskArgs(), ids=ids, labels=labels, labels_dup=labels_dup, special=list(special), n_special=len(special),
                  B=B, T=T, mask_id=int(mask_id), prob=float(prob), replace_prob=float(replace_prob), seed=seed,
                  rng_stream=int(rng_stream))
    capi.call("mlm_mask", a)
    return ids, labels


def mask_tokens(inputs, tokenizer, args):
    """Drop-in for model_utils.mask_tokens(inputs, tokenizer, args): returns (inputs, labels), inputs modified in place."""
    if getattr(tokenizer, "mask_token", None) is None:
        raise ValueError("This tokenizer does not have a mask token which is necessary for masked language modeling. "
                         "Remove the --mlm flag")
    if inputs.device.type != "cuda":
        inputs = inputs.cuda()            # the reference moves its tensors to DEVICE right after (trainer.py:60)
    labels = torch.empty_like(inputs)
    mask_id = tokenizer.convert_tokens_to_ids(tokenizer.mask_token)
    return mask_ids_(inputs, labels, prob=args.mlm_probability, mask_id=mask_id, special=_special_ids(tokenizer))


def mask_step(text_ids, twv_ids, tws_ids, Lv=None, La=None, *, prob=0.15, mask_id=103, special=BERT_SPECIAL_IDS, seed=None):
    """trainer.py:45-53 in three launches: masks the three id tensors in place and returns
    ``(text_labels [B,T], visual_labels [B,2T], speech_labels [B,2T])`` (the reference's aligned case Lv == La == T).
    For unaligned frame counts pass Lv / La: the frame half of the joint labels is then -100 (no frame targets)."""
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    B, T = text_ids.shape
    lab_t = torch.empty_like(text_ids)
    lab_v1, lab_s1 = torch.empty_like(text_ids), torch.empty_like(text_ids)
    aligned_v, aligned_s = Lv in (None, T), La in (None, T)
    dup_v = torch.empty(B, 2 * T, dtype=torch.int64, device=text_ids.device) if aligned_v else None
    dup_s = torch.empty(B, 2 * T, dtype=torch.int64, device=text_ids.device) if aligned_s else None
    mask_ids_(text_ids, lab_t, prob=prob, mask_id=mask_id, special=special, seed=seed, rng_stream=0)
    mask_ids_(twv_ids, lab_v1, prob=prob, mask_id=mask_id, special=special, seed=seed, rng_stream=1, labels_dup=dup_v)
    mask_ids_(tws_ids, lab_s1, prob=prob, mask_id=mask_id, special=special, seed=seed, rng_stream=2, labels_dup=dup_s)

    def joint(lab, dup, L):
        if dup is not None:
            return dup
        out = torch.full((B, T + L), -100, dtype=torch.int64, device=lab.device)
        out[:, :T] = lab
        return out

    return lab_t, joint(lab_v1, dup_v, Lv), joint(lab_s1, dup_s, La)
