"""msa_b200.trainer_fast.train_epoch against a line-by-line restatement of trainer.train_epoch (trainer.py:13-101) on a
recording fake model (CPU: the loop is host logic; the CUDA model is exercised by tests/test_data_gpu.py)."""
import types

import torch

from msa_b200 import trainer_fast


class _FakeModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.tensor(0.5))
        self.calls = []

    def forward(self, input_ids, token_type_ids, attention_mask, masked_labels, ap_label, sentiment):
        self.calls.append((input_ids, token_type_ids, attention_mask, masked_labels, ap_label, sentiment))
        loss = (self.w * sentiment.float().mean() + input_ids[0].float().mean() * 1e-3) ** 2
        ap = self.w.detach() * 2 + len(self.calls)
        label = loss.detach() * 0.25
        return (loss, None, None, None, ap, label, None), None


class _Opt:
    def __init__(self):
        self.steps = self.zeros = 0

    def step(self):
        self.steps += 1

    def zero_grad(self):
        self.zeros += 1


def _collate(items):          # same tuple structure as model_utils.collate (:117-142)
    B, T, D = len(items), 6, 3
    ids = torch.stack([torch.full((T,), i + 1) for i in items])
    frames = torch.ones(B, T, D, dtype=torch.float64)
    text = (ids, torch.zeros(B, dtype=torch.int64), torch.zeros(B, T, dtype=torch.int64), torch.ones(B, T, dtype=torch.float64),
            torch.tensor([float(i) for i in items]))
    vis = (ids + 1, frames, torch.ones(B, dtype=torch.int64), torch.zeros(B, T, dtype=torch.int64), frames.clone(), None)
    sp = (ids + 2, frames * 2, torch.zeros(B, dtype=torch.int64), torch.zeros(B, T, dtype=torch.int64), frames.long(), None)
    return text, vis, sp, (torch.ones(B, T, dtype=torch.float64), torch.ones(B, T, dtype=torch.int64)), None, None


def test_train_epoch_matches_the_reference_loop():
    args = types.SimpleNamespace(train_batch_size=2, gradient_accumulation_step=1, mlm=False, mlm_probability=0.15)
    data = list(range(10))            # 5 batches
    torch.manual_seed(0)
    model, opt, sched = _FakeModel(), _Opt(), _Opt()
    out = trainer_fast.train_epoch(args, model, data, opt, sched, tokenizer=None, collate_fn=_collate, device=torch.device("cpu"))
    # the `&` rule of trainer.py:96 with accumulation 1: steps on batches 2 and 4 only
    assert opt.steps == 2 and sched.steps == 2 and opt.zeros == 2
    assert len(model.calls) == 5
    ids, tt, am, labels, ap, sent = model.calls[0]
    assert len(ids) == 5 and len(labels) == 3 and labels[1].shape[-1] == 2 * ids[0].shape[-1]     # cat((l, l), -1)
    assert am[1][1].dtype == torch.float64 and am[2][1].dtype == torch.int64                      # collate's dtypes kept
    # restatement with per-step .item() (same sampler order through the same seed)
    torch.manual_seed(0)
    ref_model = _FakeModel()
    from torch.utils.data import DataLoader, RandomSampler
    tl = ll = 0.0
    nb = 0
    for step, batch in enumerate(DataLoader(data, sampler=RandomSampler(data), batch_size=2, collate_fn=_collate)):
        o, _ = ref_model(**trainer_fast.unpack_batch(batch, torch.device("cpu")))
        o[0].mean().backward()
        tl += o[0].mean().item()
        ll += o[5].mean().item()
        ap_last = o[4]
        nb += 1
    assert abs(out[0] - tl / nb) < 1e-9 and abs(out[5] - ll / nb) < 1e-9
    assert out[1] == out[2] == out[3] == 0.0
    assert torch.allclose(out[4], ap_last / nb)                    # the reference's "last ap_loss / steps" quirk
    # a data-parallel model (GradReducer attached): gradients are exchanged only by the backward that precedes
    # optimizer.step(); the micro-batches in between accumulate locally
    model3 = _FakeModel()
    model3._reducer = types.SimpleNamespace(sync=True)
    seen = []
    model3.register_forward_pre_hook(lambda m, a: seen.append(m._reducer.sync))
    trainer_fast.train_epoch(args, model3, data, _Opt(), _Opt(), None, collate_fn=_collate, device=torch.device("cpu"))
    assert seen == [False, True, False, True, False] and model3._reducer.sync is True
    # modulo stepping on request
    model2, opt2 = _FakeModel(), _Opt()
    trainer_fast.train_epoch(args, model2, data, opt2, _Opt(), None, collate_fn=_collate, device=torch.device("cpu"),
                             faithful_stepping=False)
    assert opt2.steps == 5


def test_deferred_scalars_return_every_value_in_order():
    from msa_b200.trainer_fast import DeferredScalars
    r = DeferredScalars("cpu", depth=2)
    got = [r.push(torch.tensor(float(i))) for i in range(5)]
    assert got == [None, None, 0.0, 1.0, 2.0] and r.flush() == [3.0, 4.0]
    assert r.push(torch.tensor(9.0)) is None and r.flush() == [9.0] and r.flush() == []


class _EvalModel(_FakeModel):
    def forward(self, **kw):
        out, _ = super().forward(**kw)
        return out, kw["sentiment"].float().reshape(-1, 1) * 0.5 - 1.0


def test_eval_epoch_matches_the_reference_loop():
    import numpy as np
    args = types.SimpleNamespace(val_batch_size=4, mlm=False, mlm_probability=0.15)
    data = list(range(10))            # batches of 4, 4, 2
    torch.manual_seed(1)
    model = _EvalModel()
    out = trainer_fast.eval_epoch(args, model, data, tokenizer=None, collate_fn=_collate, device=torch.device("cpu"))
    assert not model.training and len(out) == 8 and len(model.calls) == 3
    # restatement of trainer.py:127-194 with its per-batch host reads
    torch.manual_seed(1)
    ref = _EvalModel().eval()
    from torch.utils.data import DataLoader, RandomSampler
    dl = ll = 0.0
    preds, labels, nb = [], [], 0
    with torch.no_grad():
        for batch in DataLoader(data, sampler=RandomSampler(data), batch_size=4, collate_fn=_collate):
            o, logits = ref(**trainer_fast.unpack_batch(batch, torch.device("cpu")))
            dl += o[0].mean().item()
            ll += o[5].mean().item()
            ap_last = o[4]
            preds.extend(logits.numpy())
            labels.extend(batch[0][-1].numpy())
            nb += 1
    assert abs(out[0] - dl / nb) < 1e-9 and abs(out[5] - ll / nb) < 1e-9 and out[1] == out[2] == out[3] == 0.0
    assert torch.allclose(out[4], ap_last / nb)
    assert out[6].shape == (10, 1) and out[7].shape == (10,)
    assert np.array_equal(out[6], np.array(preds)) and np.array_equal(out[7], np.array(labels))


def test_score_functions_match_sklearn():
    import numpy as np
    from sklearn.metrics import accuracy_score, f1_score
    rng = np.random.default_rng(0)
    y = rng.uniform(-3, 3, 200).astype(np.float32)
    p = (y + rng.normal(0, 1.5, 200)).astype(np.float32).reshape(-1, 1)
    acc, mae, f = trainer_fast.test_MSE_score_model(p, y)
    assert abs(mae - np.mean(np.absolute(p - y))) < 1e-7             # the reference's expression, broadcasting included
    assert abs(f - f1_score(y >= 0, p >= 0, average="weighted")) < 1e-12
    assert abs(acc - accuracy_score(y >= 0, p >= 0)) < 1e-12
    yc, pc = rng.integers(0, 7, 300), rng.integers(0, 6, 300)
    acc, mae, f = trainer_fast.test_CE_score_model(pc, yc)
    assert abs(f - f1_score(yc, pc, average="weighted")) < 1e-12 and abs(acc - accuracy_score(yc, pc)) < 1e-12
    assert abs(mae - np.mean(np.abs(pc - yc))) < 1e-12


def _dist_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    args = types.SimpleNamespace(train_batch_size=2, gradient_accumulation_step=1, mlm=False, mlm_probability=0.15)
    model, opt = _FakeModel(), _Opt()
    out = trainer_fast.train_epoch(args, model, list(range(12)), opt, _Opt(), None, collate_fn=_collate,
                                   device=torch.device("cpu"), epoch=3)
    seen = sorted(int(v) for c in model.calls for v in c[5].tolist())          # sentiment == sample index
    local = sum(float(((0.5 * c[5].float().mean() + c[0][0].float().mean() * 1e-3) ** 2)) for c in model.calls)
    q.put((rank, out[0], seen, local, len(model.calls)))
    dist.destroy_process_group()


def test_train_epoch_shards_the_data_and_averages_losses_over_ranks_gloo():
    import os
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    (_, loss0, seen0, local0, n0), (_, loss1, seen1, local1, n1) = res
    assert sorted(seen0 + seen1) == list(range(12)) and not set(seen0) & set(seen1)     # disjoint shards of one shuffle
    assert n0 == n1 == 3
    assert abs(loss0 - loss1) < 1e-12 and abs(loss0 - (local0 + local1) / (n0 + n1)) < 1e-9
