"""On-device MLM masking (row N3) against the rules of the reference's model_utils.mask_tokens (oracle restatement)."""
import types

import pytest
import torch

from msa_b200 import data, synth
from oracle import mmbert_oracle as O

pytestmark = pytest.mark.gpu


class _Tok:          # the few tokenizer attributes model_utils.mask_tokens touches
    mask_token = "[MASK]"
    _pad_token = "[PAD]"
    pad_token_id = 0
    all_special_ids = [100, 102, 0, 101, 103]

    def convert_tokens_to_ids(self, tok):
        return {"[MASK]": 103}[tok]


def _ids(B=256, T=50, seed=3):
    return synth.make_batch(B, T, T, T, 47, 74, seed=seed, mlm=False)["input_ids"][0]


def test_mask_tokens_follows_the_reference_rules():
    ids0 = _ids()
    ids = ids0.clone().cuda()
    torch.manual_seed(11)
    out, labels = data.mask_tokens(ids, _Tok(), types.SimpleNamespace(mlm_probability=0.15))
    assert out.data_ptr() == ids.data_ptr()            # in place, like the reference
    r = O.mask_tokens_rules(ids0, out, labels, _Tok.all_special_ids, 103)
    assert r["label_mismatch"] == r["special_selected"] == r["changed_unselected"] == r["changed_not_to_mask"] == 0, r
    assert abs(r["select_rate"] - 0.15) < 0.015 and abs(r["replace_rate"] - 0.8) < 0.03, r
    # deterministic in the torch seed, different otherwise
    ids2 = ids0.clone().cuda()
    torch.manual_seed(11)
    out2, labels2 = data.mask_tokens(ids2, _Tok(), types.SimpleNamespace(mlm_probability=0.15))
    assert torch.equal(out, out2) and torch.equal(labels, labels2)
    ids3 = ids0.clone().cuda()
    torch.manual_seed(12)
    _, labels3 = data.mask_tokens(ids3, _Tok(), types.SimpleNamespace(mlm_probability=0.15))
    assert not torch.equal(labels, labels3)


def test_mask_step_builds_the_label_tuple_of_the_trainer():
    ids0 = _ids(B=64, T=50, seed=5)
    t, v, s = ids0.clone().cuda(), ids0.clone().cuda(), ids0.clone().cuda()
    lab_t, lab_v, lab_s = data.mask_step(t, v, s, seed=7)
    assert lab_t.shape == (64, 50) and lab_v.shape == (64, 100) and lab_s.shape == (64, 100)
    assert torch.equal(lab_v[:, :50], lab_v[:, 50:]) and torch.equal(lab_s[:, :50], lab_s[:, 50:])   # trainer.py:50,53
    assert not torch.equal(lab_t, lab_v[:, :50])        # three independent draws (trainer.py:45-47)
    for orig, new, lab in ((ids0, t, lab_t), (ids0, v, lab_v[:, :50]), (ids0, s, lab_s[:, :50])):
        r = O.mask_tokens_rules(orig, new, lab.contiguous(), data.BERT_SPECIAL_IDS, 103)
        assert r["label_mismatch"] == r["special_selected"] == r["changed_unselected"] == r["changed_not_to_mask"] == 0, r
    # unaligned frames: the frame half carries no targets
    t2, v2, s2 = ids0.clone().cuda(), ids0.clone().cuda(), ids0.clone().cuda()
    _, lab_v2, _ = data.mask_step(t2, v2, s2, Lv=77, La=50, seed=7)
    assert lab_v2.shape == (64, 127) and bool((lab_v2[:, 50:] == -100).all())


def test_masked_batch_runs_through_the_model():
    from tests.test_model_gpu import _build
    from msa_b200.params import seeded_state_dict
    ocfg = O.Cfg(hidden_size=128, num_hidden_layers=1, num_attention_heads=2, intermediate_size=256, vocab_size=30522,
                 max_position_embeddings=64)
    sd = seeded_state_dict(ocfg, "mosi", seed=1, std=0.03)
    m = _build(ocfg, "mosi", sd).train()
    batch = synth.tree_to(synth.make_batch(8, 16, 16, 16, 47, 74, seed=2, min_len=6, mlm=False), "cuda")
    ids = batch["input_ids"]
    labels = data.mask_step(ids[0], ids[3], ids[4], seed=1)
    batch["masked_labels"] = labels
    out, _ = m(**batch)
    assert torch.isfinite(out[0])
    out[0].backward()


def test_cpu_tensors_and_bad_arguments_fail_loudly():
    from msa_b200 import capi
    ids = _ids(B=4, T=8)
    with pytest.raises(capi.MMBError):
        data.mask_ids_(ids, torch.empty_like(ids))          # CPU tensor: no fallback
    with pytest.raises(ValueError):
        data.mask_tokens(ids.cuda(), types.SimpleNamespace(mask_token=None), types.SimpleNamespace(mlm_probability=0.15))


def test_sync_free_train_epoch_on_the_cuda_model():
    """msa_b200.trainer_fast.train_epoch (row N1) drives the CUDA model with on-device masking and the fused AdamW:
    losses come back finite, the optimizer stepped on every second batch (trainer.py:96), parameters moved."""
    from msa_b200 import trainer_fast
    from msa_b200.optim import FusedAdamW
    from msa_b200.params import seeded_state_dict
    from tests.test_model_gpu import _build
    ocfg = O.Cfg(hidden_size=128, num_hidden_layers=1, num_attention_heads=2, intermediate_size=256, vocab_size=30522,
                 max_position_embeddings=64)
    sd = seeded_state_dict(ocfg, "mosi", seed=1, std=0.03)
    m = _build(ocfg, "mosi", sd)
    N, T = 12, 16
    full = synth.make_batch(N, T, T, T, 47, 74, seed=4, min_len=6, mlm=False)

    def collate(idx):            # the tuple structure of model_utils.collate (:117-142), sliced from one synthetic batch
        i = torch.tensor(idx)
        ids, vis, aud = full["input_ids"][0][i], full["input_ids"][1][i], full["input_ids"][2][i]
        m_t, (m_tv, m_v), (m_ts, m_s) = full["attention_mask"]
        tt = full["token_type_ids"][0][i]
        text = (ids, None, tt, m_t[i], full["sentiment"][i])
        visual = (ids.clone(), vis, full["ap_label"][0][i], tt, m_v[i], None)
        speech = (ids.clone(), aud, full["ap_label"][1][i], tt, m_s[i], None)
        return text, visual, speech, (m_tv[i], m_ts[i]), None, None

    args = types.SimpleNamespace(train_batch_size=4, gradient_accumulation_step=1, mlm=True, mlm_probability=0.15)
    opt = FusedAdamW(m, lr=1e-3)
    steps = []
    orig_step = opt.step
    opt.step = lambda *a, **k: (steps.append(1), orig_step(*a, **k))[1]
    sched = types.SimpleNamespace(step=lambda: None)
    w0 = m.classifier1_1.weight.detach().clone()
    torch.manual_seed(3)
    out = trainer_fast.train_epoch(args, m, list(range(N)), opt, sched, _Tok(), collate_fn=collate)
    assert len(steps) == 1                                   # 3 batches: only (step+1) = 2 satisfies (step+1) & 1 == 0
    assert all(isinstance(v, float) and v == v for v in (out[0], out[5]))
    assert torch.isfinite(out[4]).all()
    assert not torch.equal(w0, m.classifier1_1.weight.detach())


def test_eval_epoch_on_the_cuda_model_keeps_every_batch_predictions():
    """trainer_fast.eval_epoch over THREE batches of the real model: the model's outputs live in buffers of a cached
    launch plan that the next batch overwrites, so predictions must be handed out as copies — the collected [N, 1] array
    must equal per-batch forward passes (an alias would repeat the last batch), and so must the loss averages."""
    from msa_b200 import trainer_fast
    from msa_b200.params import seeded_state_dict
    from tests.test_model_gpu import _build
    ocfg = O.Cfg(hidden_size=128, num_hidden_layers=1, num_attention_heads=2, intermediate_size=256, vocab_size=30522,
                 max_position_embeddings=64)
    sd = seeded_state_dict(ocfg, "mosi", seed=1, std=0.03)
    m = _build(ocfg, "mosi", sd)
    N, T, bs = 12, 16, 4
    full = synth.make_batch(N, T, T, T, 47, 74, seed=4, min_len=6, mlm=False)

    def collate(idx):
        i = torch.tensor(idx)
        ids, vis, aud = full["input_ids"][0][i], full["input_ids"][1][i], full["input_ids"][2][i]
        m_t, (m_tv, m_v), (m_ts, m_s) = full["attention_mask"]
        tt = full["token_type_ids"][0][i]
        text = (ids, None, tt, m_t[i], full["sentiment"][i])
        visual = (ids.clone(), vis, full["ap_label"][0][i], tt, m_v[i], None)
        speech = (ids.clone(), aud, full["ap_label"][1][i], tt, m_s[i], None)
        return text, visual, speech, (m_tv[i], m_ts[i]), None, None

    args = types.SimpleNamespace(val_batch_size=bs, mlm=False, mlm_probability=0.15)
    torch.manual_seed(5)
    out = trainer_fast.eval_epoch(args, m, list(range(N)), _Tok(), collate_fn=collate)
    preds, labels = out[6], out[7]
    assert preds.shape == (N, 1) and labels.shape == (N,)
    # direct evaluation, sample by sample lookup through the sentiment values (unique floats identify the sample)
    m.eval()
    want = {}
    losses = []
    with torch.no_grad():
        for b0 in range(0, N, bs):
            kw = trainer_fast.unpack_batch(collate(list(range(b0, b0 + bs))), "cuda", _Tok(), args)
            o, lg = m(**kw)
            for s_, l_ in zip(full["sentiment"][b0:b0 + bs].tolist(), lg.view(-1).tolist()):
                want[round(s_, 6)] = l_
    got = {round(float(s_), 6): float(p_) for s_, p_ in zip(labels, preds[:, 0])}
    assert len(got) == N                                            # every sample once (no repeated last batch)
    # batches are drawn in another order, so CPC's in-batch negatives differ; the regression head itself does not
    # depend on the other samples of a batch
    for k, v in want.items():
        assert abs(got[k] - v) < 1e-4, (k, got[k], v)
    assert len(set(round(v, 5) for v in got.values())) > N // 2     # not one value repeated


def test_staged_prefetcher_reuses_pinned_buffers_for_pageable_batches():
    from msa_b200.trainer_fast import DevicePrefetcher
    batches = [{"x": torch.full((300, 33), float(i)), "y": (torch.arange(5) + i,)} for i in range(9)]
    pf = DevicePrefetcher(iter(batches), "cuda", stage=True)
    got = [(float(d["x"].mean()), int(d["y"][0][0])) for d in pf]
    assert got == [(float(i), i) for i in range(9)]
    assert pf.h2d_bytes == 9 * (300 * 33 * 4 + 5 * 8)
    assert all(t.is_pinned() for t in (pf._pin[0]["x"], pf._pin[1]["x"], pf._pin[2]["x"]))


def test_prefetcher_and_deferred_reads_deliver_every_batch_in_order():
    """DevicePrefetcher reuses two device buffer sets while copies run ahead on a copy stream: with a slow consumer
    every batch must still arrive intact and in order (a buffer overwritten early would show the next batch's values),
    including a last batch of another shape; DeferredScalars must hand back one value per push, in order."""
    from msa_b200.trainer_fast import DeferredScalars, DevicePrefetcher
    batches = [{"a": (torch.full((1000, 257), float(i)).pin_memory(), (torch.arange(64) + i).pin_memory()),
                "b": torch.tensor([i]).pin_memory()} for i in range(7)]
    batches.append({"a": (torch.full((10, 257), 7.0), torch.arange(3) + 7), "b": torch.tensor([7])})   # short, pageable
    want = [float(b["a"][0].sum() + b["a"][1].sum() + b["b"].sum()) for b in batches]
    big = torch.randn(4096, 4096, device="cuda")
    reader, got = DeferredScalars("cuda"), []
    for dev in DevicePrefetcher(iter(batches), "cuda"):
        assert dev["a"][0].is_cuda
        for _ in range(6):
            big @ big                     # keeps the consumer stream busy: the copies run ahead of it
        v = reader.push(dev["a"][0].sum() + dev["a"][1].sum() + dev["b"].sum())
        if v is not None:
            got.append(v)
    got += reader.flush()
    assert got == want
    assert list(DevicePrefetcher(iter([]), "cuda")) == []
