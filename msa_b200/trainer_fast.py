"""Sync-free drop-in for the reference's training-epoch loop.

``train_epoch(args, model, traindata, optimizer, scheduler, tokenizer)`` has the signature, batch unpacking, model call
and return value of ``trainer.train_epoch`` (trainer.py:13-101), with the three things that serialise host and device
every step removed:

* the two executed ``.item()`` read-backs per step (trainer.py:85, :93) — losses are accumulated on the device and read
  once at the end of the epoch;
* MLM masking through Python lists and boolean-index writes (model_utils.py:6-39, three times per step) — one kernel
  launch per id tensor (``msa_b200.data.mask_tokens``);
* pageable, blocking host->device copies (trainer.py:49-64) — pinned buffers and ``non_blocking`` copies.

What is deliberately kept: the optimizer stepping rule ``(step + 1) & args.gradient_accumulation_step == 0``
(trainer.py:96 — a bitwise AND, so with the default 1 the optimizer steps on every second batch) unless
``faithful_stepping=False`` selects the usual modulo rule; and the returned tuple, including the reference's quirk that
the fifth element is the LAST step's alignment loss divided by the step count (:101).
"""
import torch
from torch.utils.data import DataLoader, RandomSampler

from . import data


def _to_dev(t, device):
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    if t.device.type == "cpu" and device.type == "cuda":
        t = t.pin_memory()
    return t.to(device, non_blocking=True)


def unpack_batch(batch, device, tokenizer=None, args=None):
    """trainer.py:41-64: collate output -> keyword arguments of MMBertForPretraining.forward (tensors on ``device``)."""
    text_batch, visual_batch, speech_batch, attention_batch = batch[0], batch[1], batch[2], batch[3]
    text_ids, twv_ids, tws_ids = (_to_dev(x, device) for x in (text_batch[0], visual_batch[0], speech_batch[0]))
    if args is not None and getattr(args, "mlm", False):
        # mask_tokens modifies its input in place (like the reference): work on device copies
        text_ids, twv_ids, tws_ids = text_ids.clone(), twv_ids.clone(), tws_ids.clone()
        text_ids, text_lab = data.mask_tokens(text_ids, tokenizer, args)
        twv_ids, vis_lab = data.mask_tokens(twv_ids, tokenizer, args)
        tws_ids, sp_lab = data.mask_tokens(tws_ids, tokenizer, args)
    else:
        text_lab, vis_lab, sp_lab = text_ids, twv_ids, tws_ids                     # trainer.py:45-47, else branch
    visual_inputs, speech_inputs = _to_dev(visual_batch[1], device), _to_dev(speech_batch[1], device)
    vis_lab = torch.cat((vis_lab, vis_lab), dim=-1)                                  # trainer.py:50
    sp_lab = torch.cat((sp_lab, sp_lab), dim=-1)                                     # trainer.py:53
    return dict(
        input_ids=(text_ids, visual_inputs, speech_inputs, twv_ids, tws_ids),
        token_type_ids=(_to_dev(text_batch[2], device), _to_dev(visual_batch[3], device), _to_dev(speech_batch[3], device)),
        attention_mask=(_to_dev(text_batch[3], device),
                        (_to_dev(attention_batch[0], device), _to_dev(visual_batch[4], device)),
                        (_to_dev(attention_batch[1], device), _to_dev(speech_batch[4], device))),
        masked_labels=(text_lab, vis_lab, sp_lab),
        ap_label=(_to_dev(visual_batch[2], device), _to_dev(speech_batch[2], device)),
        sentiment=_to_dev(text_batch[-1], device),
    )


def train_epoch(args, model, traindata, optimizer, scheduler, tokenizer, *, collate_fn=None, device=None,
                faithful_stepping=True):
    if collate_fn is None:
        import model_utils                      # the reference's module (on PYTHONPATH in the drop-in setting)
        collate_fn = model_utils.collate
    if device is None:
        device = next(model.parameters()).device
    loader = DataLoader(traindata, sampler=RandomSampler(traindata), batch_size=args.train_batch_size, collate_fn=collate_fn)
    sums = torch.zeros(5, device=device, dtype=torch.float64)      # train, text, visual, speech, label
    n, ap_last = 0, None
    model.train()
    for step, batch in enumerate(loader):
        outputs, _ = model(**unpack_batch(batch, device, tokenizer, args))
        loss = outputs[0]
        loss.mean().backward()
        with torch.no_grad():
            sums[0] += loss.detach().mean()
            for i in (1, 2, 3):                                    # always None in the reference (:394, :445)
                if outputs[i] is not None:
                    sums[i] += outputs[i].detach().mean()
            sums[4] += outputs[5].detach().mean()
            ap_last = outputs[4]
        n += 1
        accum = args.gradient_accumulation_step
        do_step = ((step + 1) & accum) == 0 if faithful_stepping else ((step + 1) % accum) == 0
        if do_step:
            optimizer.step()
            scheduler.step()
            optimizer.zero_grad()
    if n == 0:
        raise ValueError("empty training set")
    s = (sums / n).tolist()                                        # the epoch's only device->host read
    return s[0], s[1], s[2], s[3], (ap_last / n if ap_last is not None else None), s[4]
