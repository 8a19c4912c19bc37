"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol include/mmbert_sm100.h
declares, the ctypes structures generated from the header match the C compiler's layout, the flat parameter
store / state_dict surface mirrors the reference, the synthetic batches have the reference's dtypes, and the
data-parallel bucket schedule works across two gloo ranks.  No compute entry point is called (no GPU here)."""
import ctypes
import os
import subprocess
import sys

import pytest
import torch

from msa_b200 import capi, synth
from msa_b200.params import NO_GRAD, TIED, BertShape, param_shapes, train_gflop_per_sample
from msa_b200.store import FlatStore

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _small():
    return BertShape(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, vocab_size=256,
                     max_position_embeddings=32)


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    assert len(capi.DECLARED_FUNCTIONS) >= 20
    for name in capi.DECLARED_FUNCTIONS:
        assert hasattr(L, name), name
    assert L.mmb_version() == 204
    assert capi.launch_count() >= 0


def test_ctypes_structs_match_the_c_compiler(tmp_path):
    names = ["mmb_gemm_args", "mmb_drln_fwd_args", "mmb_drln_bwd_args", "mmb_colsum_args", "mmb_attn_args",
             "mmb_pack_args", "mmb_embed_args", "mmb_ce_args", "mmb_heads_args", "mmb_adamw_args"]
    src = tmp_path / "sz.c"
    src.write_text('#include "%s"\n#include <stdio.h>\nint main(){%s return 0;}\n' % (
        capi.HEADER_PATH, "".join('printf("%%zu\\n", sizeof(%s));' % n for n in names)))
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    for n, sz in zip(names, sizes):
        assert ctypes.sizeof(capi._make_struct(n)) == sz, n


def test_compute_entry_point_without_gpu_fails_loudly():
    """No silent CPU path: on a box without a CUDA device the device check reports an error code."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rc = capi.lib().mmb_check_device()
    assert rc != capi.MMB_OK
    with pytest.raises(capi.MMBError):
        capi.check(rc, "mmb_check_device")


def test_flat_store_layout():
    shape = _small()
    st = FlatStore(shape, "mosi")
    shapes = param_shapes(shape, "mosi")
    assert set(st.offsets) == set(shapes)
    H = shape.hidden_size
    for l in range(shape.num_hidden_layers):       # fused QKV: the three weights / biases are adjacent
        p = f"bert.encoder.layer.{l}.attention.self."
        assert st.offsets[p + "key.weight"] == st.offsets[p + "query.weight"] + H * H
        assert st.offsets[p + "value.weight"] == st.offsets[p + "key.weight"] + H * H
        assert st.offsets[p + "key.bias"] == st.offsets[p + "query.bias"] + H
        assert st.offsets[p + "value.bias"] == st.offsets[p + "key.bias"] + H
    for n, off in st.offsets.items():
        assert off % 64 == 0
        if n in NO_GRAD:
            assert off >= st.trainable_end
        elif "bias" in n or "LayerNorm.weight" in n:      # train.py:77-91 no_decay rule
            assert st.decay_end <= off < st.trainable_end, n
        else:
            assert off < st.decay_end, n


def test_module_surface_matches_reference_names():
    from msa_b200.api import MMBertForPretraining
    shape = _small()
    m = MMBertForPretraining(shape)
    m.bert.set_joint_embeddings("mosi")
    names = [n for n, _ in m.named_parameters()]
    assert names == list(param_shapes(shape, "mosi"))                      # same names, same order
    sd = m.state_dict()
    assert set(sd) == set(param_shapes(shape, "mosi")) | set(TIED)
    for alias, canon in TIED.items():
        assert sd[alias].data_ptr() == sd[canon].data_ptr()
    assert m.num_labels == 7 and m.alpha == 1 and m.beta == 1
    m.set_alpha_beta(0.3, 0.6)
    assert (m.alpha, m.beta) == (0.3, 0.6)
    # materialising on CPU keeps values, makes parameters views of one buffer and survives a state_dict round trip
    before = {k: v.clone() for k, v in sd.items()}
    m._ensure_store(torch.device("cpu"))
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k]), k
    flat = m._store.flat
    p = m.bert.encoder.layer._modules["0"].attention.self.query.weight
    assert flat.data_ptr() <= p.data_ptr() < flat.data_ptr() + 4 * flat.numel()
    m.load_state_dict(before)
    assert m._store.is_current(m._named(), torch.device("cpu"))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference checkout not present")
def test_state_dict_keys_equal_the_live_reference():
    from transformers import BertConfig
    from oracle import ref_loader
    from msa_b200.api import MMBertForPretraining
    kw = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, vocab_size=256,
              max_position_embeddings=32)
    ref = ref_loader.build_model(BertConfig(**kw), "ur_funny")
    m = MMBertForPretraining(BertShape(**kw))
    m.bert.set_joint_embeddings("ur_funny")
    ref_sd, sd = ref.state_dict(), m.state_dict()
    assert list(ref_sd) == list(sd) or set(ref_sd) == set(sd)
    for k in ref_sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in m.named_parameters()]
    m.load_state_dict(ref_sd)     # checkpoints interchange (trainer.py:269 / sampling.py:344) ...
    for k, v in m.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k
    torch.manual_seed(5)
    m2 = MMBertForPretraining(BertShape(**kw))
    m2.bert.set_joint_embeddings("ur_funny")
    res = ref.load_state_dict(m2.state_dict(), strict=True)   # ... in both directions
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(ref.state_dict()["cls.predictions.decoder.weight"], m2.state_dict()["bert.embeddings.word_embeddings.weight"])


def test_synthetic_batch_has_reference_dtypes_and_layout():
    b = synth.make_batch(4, 12, 12, 20, 47, 74, seed=3)
    ids_t, vis, aud, ids_v, ids_s = b["input_ids"]
    assert ids_t.dtype == torch.int64 and vis.dtype == torch.float64 and aud.dtype == torch.float64
    m_t, (m_tv, m_v), (m_ts, m_s) = b["attention_mask"]
    assert m_t.dtype == torch.float64 and m_tv.dtype == torch.float64 and m_v.dtype == torch.float64
    assert m_ts.dtype == torch.int64 and m_s.dtype == torch.int64             # model_utils.py:132-136
    assert bool(m_tv.eq(1).all()) and bool(m_ts.eq(1).all())                  # the collate typo: never masked
    assert tuple(m_v.shape) == (4, 12, 47) and tuple(m_s.shape) == (4, 20, 74)
    lab_t, lab_v, lab_s = b["masked_labels"]
    assert tuple(lab_v.shape) == (4, 24) and tuple(lab_s.shape) == (4, 32)
    assert bool((lab_v[:, :12] == lab_v[:, 12:]).all())                        # aligned: cat((labels, labels))
    assert bool((lab_s[:, 12:] == -100).all())                                 # unaligned: frame half unlabelled
    assert all(int((l != -100).sum()) >= 1 for l in (lab_t, lab_v, lab_s))
    assert bool((ids_t[:, 0] == 101).all())
    n = (ids_t != 0).sum(1)
    assert bool((vis[0, int(n[0]) - 1:] == 0).all())                           # SEP row + padding rows are zero


def test_flop_accounting_matches_survey_table():
    assert abs(train_gflop_per_sample(BertShape(), synth.WORKLOADS["mosi_aligned_b64"]) - 166.02) < 0.01
    assert abs(train_gflop_per_sample(BertShape(), synth.WORKLOADS["mosei_unaligned_b64"]) - 819.35) < 0.01
    assert abs(train_gflop_per_sample(BertShape(), synth.WORKLOADS["ur_funny_b64"]) - 166.09) < 0.01


def test_executed_flop_accounting_and_live_row_fraction(monkeypatch):
    """bench.py's two padding-dependent figures on hand-made masks: the attention term counted as executed (forward
    4 q e H with q = the query rows of the 128-row tiles that start before e, backward 8 e^2 H; everything else dense) and
    the share of packed rows before their sequence's last unmasked key."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    shape, wl = BertShape(), synth.WORKLOADS["mosei_unaligned_b64"]
    B, T, L = 2, wl.T, 500

    def batch(text_len, frame_len):
        m_t = torch.zeros(B, T)
        m_t[:, :text_len] = 1
        m_f = torch.zeros(B, L, 3)
        m_f[:, :frame_len] = 1
        ones = torch.ones(B, T)
        return {"attention_mask": (m_t, (ones, m_f), (ones, m_f[..., 0].clone()))}

    full = batch(T, L)
    dense = train_gflop_per_sample(shape, wl)
    assert abs(bench.executed_train_gflop_per_sample(shape, wl, [full]) - dense) < 1e-6 * dense
    assert bench.live_row_fraction([full]) == 1.0
    b = batch(20, 100)          # text pass: e = 20 of 50; joint passes: e = 50 + 100 = 150 of 550 (a hole-free prefix)
    H, N = shape.hidden_size, shape.num_hidden_layers
    want = dense
    for S, e in ((T, 20), (T + L, 150), (T + L, 150)):
        q = min(S, -(-e // 128) * 128)
        want += N * H * (4.0 * q * e + 8.0 * e * e - 12.0 * S * S) / 1e9
    assert abs(bench.executed_train_gflop_per_sample(shape, wl, [b]) - want) < 1e-6 * dense
    monkeypatch.setenv("MMB_ATTN_FWD_QSKIP", "0")          # forward computes every query row again
    want0 = dense + sum(N * H * (4.0 * S * e + 8.0 * e * e - 12.0 * S * S) / 1e9 for S, e in ((T, 20), (T + L, 150), (T + L, 150)))
    assert abs(bench.executed_train_gflop_per_sample(shape, wl, [b]) - want0) < 1e-6 * dense
    assert abs(bench.live_row_fraction([b]) - (20 + 150 + 150) / (T + 2 * (T + L))) < 1e-12
    hole = batch(20, 100)       # an unmasked key at the very end: nothing is padding any more
    hole["attention_mask"][1][1][:, -1, 0] = 1
    assert abs(bench.live_row_fraction([hole]) - (20 + 550 + 150) / (T + 2 * (T + L))) < 1e-12


def test_bucket_schedule_partitions_the_gradient_buffer():
    from msa_b200.ddp import bucket_schedule, check_partition
    for layers in (1, 2, 12):
        shape = BertShape(num_hidden_layers=layers)
        st = FlatStore(shape, "mosei")
        sched = bucket_schedule(st, layers)
        assert check_partition(sched, st.trainable_end)
        assert [t for t, _ in sched] == ["heads"] + [("layer", l) for l in range(layers - 1, -1, -1)] + ["final"]


def test_root_level_drop_in_modules_import():
    """`from MMBertForPretraining import MMBertForPretraining`, `import config` ... resolve to this repo's classes
    (checked in a subprocess: inside pytest the same top-level names may be held by the reference)."""
    code = ("import config, MMBertEmbedding, MMBertForPretraining as M;"
            "assert config.TEXTDIM == 1024 and config.MOSIVISUALDIM == 47 and config.total_vocab_size == 30522;"
            "assert M.MMBertForPretraining.__module__ == 'msa_b200.api';"
            "assert MMBertEmbedding.JointEmbeddings(128, 0.5, 'mosi').Wv.weight.shape == (128, 47);"
            "print('ok')")
    out = subprocess.check_output([sys.executable, "-c", code], cwd=ROOT, text=True)
    assert out.strip().endswith("ok")


def _ddp_worker(rank, world, port, q, mode="overlap"):
    import torch.distributed as dist
    from msa_b200.ddp import GradReducer
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shape = _small()
    st = FlatStore(shape, "mosi")
    st.grad = torch.full((st.total,), float(rank + 1))
    st.grad[st.trainable_end:] = -7.0                      # never-touched parameters must not be communicated
    red = GradReducer(st, shape.num_hidden_layers, mode=mode)
    for _, ranges in red.sched:                            # same order on every rank (no-ops in deferred mode)
        red.reduce_bucket(ranges)
    red.finish()
    ok = bool((st.grad[:st.trainable_end] == sum(range(1, world + 1))).all()) and bool((st.grad[st.trainable_end:] == -7.0).all())
    # gradient accumulation: two micro-batches accumulate into the buffer the in-place SUM all-reduce works on; only the
    # backward that precedes optimizer.step() may reduce (no_sync on the first), else micro-batch 1 counts world times
    st.grad.zero_()
    g1, g2 = float(rank + 1), 10.0 * (rank + 1)
    with red.no_sync():
        st.grad[:st.trainable_end] += g1
        for _, ranges in red.sched:
            red.reduce_bucket(ranges)
        red.finish()
    ok = ok and bool((st.grad[:st.trainable_end] == g1).all())                       # nothing was communicated
    st.grad[:st.trainable_end] += g2
    for _, ranges in red.sched:
        red.reduce_bucket(ranges)
    red.finish()
    want = sum((r + 1) + 10.0 * (r + 1) for r in range(world))                       # sum over ranks of (g1 + g2)
    ok = ok and bool((st.grad[:st.trainable_end] == want).all())
    if mode == "deferred":
        # pipelined form: finish() only ISSUES the pieces; they tile the trainable range and wait_all() completes them
        red.chunks = 3
        st.grad[:st.trainable_end] = float(rank + 1)
        red.finish()
        pieces = [(a, b) for a, b, _ in red.inflight]
        ok = ok and len(pieces) == 3 and pieces[0][0] == 0 and pieces[-1][1] == st.trainable_end
        ok = ok and all(pieces[i][1] == pieces[i + 1][0] for i in range(2))
        red.wait_all()
        ok = ok and not red.inflight and bool((st.grad[:st.trainable_end] == sum(range(1, world + 1))).all())
        ok = ok and bool((st.grad[st.trainable_end:] == 0.0).all())                  # (zeroed above) still not communicated
    q.put((rank, ok, red.bytes_per_step))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["deferred", "overlap"])
def test_gradient_reducer_two_ranks_gloo(mode):
    """Both schedules of msa_b200.ddp.GradReducer (one all-reduce after the backward sweep / per-layer buckets under it)
    sum the trainable range exactly once per stepping backward and leave the never-touched tail alone."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + (7 if mode == "overlap" else 0)) % 2000
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q, mode)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res


def test_wgrad_split_k_fills_whole_waves():
    """engine._split_k: the (tiles x splits) work units of a wgrad GEMM should fill whole waves of the 74 CTA pairs."""
    from msa_b200.engine import _split_k
    for (m, n, k) in ((768, 3072, 16000), (3072, 768, 73600), (2304, 768, 16000), (768, 768, 73600), (30522, 768, 16000)):
        sk = _split_k(m, n, k)
        tiles = -(-m // 256) * -(-n // 256)
        units = tiles * sk
        eff = units / (-(-units // 74) * 74)
        assert eff >= 0.95, (m, n, k, sk, eff)
        assert (k + 63) // 64 // sk >= 8            # every K-slice keeps at least 8 k-blocks
    assert _split_k(768, 3072, 600) == 1             # too short to split
    assert _split_k(30522, 768, 16000) == 1          # already 4.86 waves of tiles


def test_attention_workspace_size_is_exported():
    from msa_b200 import capi
    L = capi.lib()
    import ctypes
    L.mmb_attn_bwd_workspace_bytes.restype = ctypes.c_size_t
    L.mmb_attn_bwd_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int]
    assert L.mmb_attn_bwd_workspace_bytes(1000, 12) == 2 * 12 * 1000 * 16     # Rq + Rk planes of 16-byte records
    assert L.mmb_attn_bwd_workspace_bytes(0, 12) == 0
