"""Host-side mirror of the reference's Python surface for the MMBert hot path.

Same class names, constructor arguments, attribute / parameter names, ``forward`` signature and return
structure as the reference:
  MMBertForPretraining  /root/reference/MMBertForPretraining.py:304-449
  MMBertModel           :13-285   (parameter container; its arithmetic runs inside the packed step)
  MMBertPreTrainingHeads :287-302
  JointEmbeddings, CPC  /root/reference/MMBertEmbedding.py:34-72, :7-32
Call sites that keep working unchanged: train.py:70-76 (construct, ``.num_labels``,
``.bert.set_joint_embeddings(dataset)``, ``.set_alpha_beta``, ``.cuda()``, ``named_parameters()``),
trainer.py:66-73,83 (``model(input_ids=..., ...)``, ``loss.mean().backward()``), :269 (``state_dict()``).

All arithmetic is done by libmmbert_sm100.so through msa_b200.engine.Plan; there is no torch fallback:
calling ``forward`` on a CPU model or without the built library raises.
"""
from collections import OrderedDict

import torch
from torch import nn

from . import capi
from .engine import Plan
from .params import NO_GRAD, TIED, param_shapes
from .store import FlatStore
from .synth import DATASET_DIMS


class _Node(nn.Module):
    """Structural module: holds parameters / children under the reference's names."""


def _attach(root, dotted, param):
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Node())
        mod = mod._modules[p]
    mod.register_parameter(parts[-1], param)


def _init_param(name, shape, std):
    """HF BERT ``_init_weights``: Linear / Embedding weights ~ N(0, initializer_range), biases 0, LayerNorm (1, 0);
    the word-embedding padding row is zero."""
    if "LayerNorm.weight" in name:
        return torch.ones(shape)
    if name.endswith("bias"):
        return torch.zeros(shape)
    w = torch.empty(shape).normal_(0.0, std)
    if name == "bert.embeddings.word_embeddings.weight":
        w[0].zero_()
    return w


def _torch_linear_init(weight_shape):
    """torch.nn.Linear default init (kaiming_uniform(a=sqrt(5)) weight, uniform bias) — what JointEmbeddings
    keeps in the reference because it is attached after ``init_weights``."""
    lin = nn.Linear(weight_shape[1], weight_shape[0])
    return lin.weight.detach().clone(), lin.bias.detach().clone()


class CPC(_Node):
    """Parameter container with the reference's constructor (MMBertEmbedding.py:7-19); the loss itself is part of
    the fused heads kernels (csrc/heads.cu)."""

    def __init__(self, x_size, y_size, n_layers=1, activation="Tanh"):
        super().__init__()
        self.x_size, self.y_size, self.layers = x_size, y_size, n_layers
        self.net = nn.Linear(in_features=y_size, out_features=x_size)


class JointEmbeddings(_Node):
    """Parameter container with the reference's constructor (MMBertEmbedding.py:34-55).  W_cv / W_cs are the
    reference's dead parameters: kept for state_dict compatibility, never touched by the path."""

    def __init__(self, hidden_size, dropout_prob, dataset):
        super().__init__()
        if dataset not in DATASET_DIMS:
            raise ValueError(f"unknown dataset {dataset!r}")
        self.VISUALDIM, self.SPEECHDIM = DATASET_DIMS[dataset]
        H = hidden_size
        self.W_cv = nn.Linear(self.VISUALDIM + H, H)
        self.W_cs = nn.Linear(self.SPEECHDIM + H, H)
        self.Wv = nn.Linear(self.VISUALDIM, H)
        self.Ws = nn.Linear(self.SPEECHDIM, H)
        self.LayerNorm = nn.LayerNorm(hidden_size)     # eps 1e-5 (torch default), unlike BERT's 1e-12
        self.dropout = nn.Dropout(dropout_prob)


class MMBertModel(_Node):
    """bert.* parameters (embeddings, encoder, pooler, jointEmbeddings)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.dataset = None

    def set_joint_embeddings(self, dataset):
        self.dataset = dataset
        self.jointEmbeddings = JointEmbeddings(self.config.hidden_size, 0.5, dataset)
        owner = getattr(self, "_owner", None)
        if owner is not None:
            owner()._joint_changed()


class MMBertPreTrainingHeads(_Node):
    """cls.* parameters (LM transform, tied decoder, seq_relationship, align)."""


class _StepFn(torch.autograd.Function):
    """Anchors the packed step in the autograd graph: forward runs the forward plan, backward the backward
    plan, which accumulates straight into the flat gradient buffer (= every Parameter's ``.grad`` view)."""

    @staticmethod
    def forward(ctx, anchor, model, plan):
        ctx.model, ctx.plan = model, plan
        Plan.run(plan.fwd)
        model.launches += len(plan.fwd)
        return plan.losses[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        model, plan = ctx.model, ctx.plan
        plan.gscale.copy_(grad_out.reshape(1))
        model._prepare_grads()
        hooks = model._backward_hooks_for(plan)
        Plan.run(plan.bwd, hooks)
        model.launches += len(plan.bwd)
        if model._post_backward is not None:
            model._post_backward()
        return None, None, None


class MMBertForPretraining(_Node):
    def __init__(self, config):
        super().__init__()
        import weakref
        self.config = config
        self.num_labels = 7
        self.alpha = 1
        self.beta = 1
        self.bert = MMBertModel(config)
        self.cls = MMBertPreTrainingHeads()
        self.bert._owner = weakref.ref(self)
        self._store = None
        self._plans = {}
        self._keep = None
        self._post_backward = None
        self._bwd_hooks = None
        self._reducer = None       # msa_b200.ddp.GradReducer once attached (train_epoch toggles its ``sync`` flag)
        self.launches = 0          # kernel-launching C calls issued so far (bench.py reports the per-step count)
        self.dense_mlm = True
        # False (default): the masked-LM cross entropy is fused into the tied-decoder GEMM and the [rows, vocab] logits are
        # never written — pred_t / pred_v / pred_s (outputs[7], [9], [11]; trainer.py never reads them) come back as None.
        # True: the reference's full output tuple, at the price of 61 KB of HBM per sequence position.
        self.materialize_logits = False
        # Forward-only (no_grad) calls on an eval() model replay the launch plan as one CUDA graph (engine.Plan.forward_graph).
        self.use_cuda_graph = False
        # "bf16": tcgen05 tensor-core path (training and inference).  "fp32": validation path, forward only, fp32
        # storage and CUDA-core arithmetic (engine_f32.PlanF32) for the reference's 1e-4 fp32 tolerance.
        self.precision = "bf16"
        H = config.hidden_size
        std = getattr(config, "initializer_range", 0.02)
        # every parameter except bert.jointEmbeddings.* (created by set_joint_embeddings, as in the reference)
        shapes = param_shapes(config, "mosi")
        for name, shape in shapes.items():
            if name.startswith("bert.jointEmbeddings."):
                continue
            _attach(self, name, nn.Parameter(_init_param(name, shape, std)))
        for alias, canon in TIED.items():
            _attach(self, alias, self.get_parameter(canon))

    # ------------------------------------------------------------------ reference API
    @classmethod
    def from_pretrained(cls, name_or_path, *args, **kwargs):
        """The reference calls ``MMBertForPretraining.from_pretrained('bert-base-uncased')`` (train.py:70); that
        needs the HF hub.  A local directory with config.json (+ optional pytorch_model.bin / model.pt) works."""
        import os
        from transformers import BertConfig
        cfg = BertConfig.from_pretrained(name_or_path)
        model = cls(cfg)
        for fn in ("pytorch_model.bin", "model.pt"):
            path = os.path.join(str(name_or_path), fn)
            if os.path.isfile(path):
                model.load_state_dict(torch.load(path, map_location="cpu"), strict=False)
                break
        return model

    def set_alpha_beta(self, alpha, beta):
        self.alpha = alpha
        self.beta = beta

    def get_input_embeddings(self):
        return self.bert.embeddings.word_embeddings

    # ------------------------------------------------------------------ storage
    def _joint_changed(self):
        self._store = None
        self._plans = {}

    def _named(self):
        return OrderedDict((n, p) for n, p in self.named_parameters())

    def _ensure_store(self, device):
        if self.bert.dataset is None:
            raise capi.MMBError("call model.bert.set_joint_embeddings(dataset) first (train.py:72)")
        params = self._named()
        if self._store is None:
            self._store = FlatStore(self.config, self.bert.dataset)
        if not self._store.is_current(params, device):
            self._store.materialize(params, device)
            self._plans = {}
            self._params = params
            self._trainable = [params[n] for n in self._store.offsets if n not in NO_GRAD]
            self._grad_views = [self._store.view(n, self._store.grad) for n in self._store.offsets if n not in NO_GRAD]
        return self._store

    def _signature(self):
        return sum(p._version for p in self._params.values())

    def _prepare_grads(self):
        """Makes every trainable Parameter's ``.grad`` the matching view of the flat gradient buffer, with
        PyTorch's accumulate semantics: ``None`` grads (after ``zero_grad(set_to_none=True)``) start from zero."""
        st = self._store
        none = [i for i, p in enumerate(self._trainable) if p.grad is None]
        if len(none) == len(self._trainable):
            st.grad[:st.trainable_end].zero_()
        else:
            for i in none:
                self._grad_views[i].zero_()
        for i, p in enumerate(self._trainable):
            g = self._grad_views[i]
            if p.grad is None:
                p.grad = g
            elif p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)
                p.grad = g

    def _backward_hooks_for(self, plan):
        return self._bwd_hooks(plan) if self._bwd_hooks is not None else None

    def _plan(self, B, T, Lv, La, device, needs_grad):
        if self.precision not in ("bf16", "fp32"):
            raise capi.MMBError(f"precision must be 'bf16' or 'fp32', not {self.precision!r}")
        fp32 = self.precision == "fp32"
        if not self.dense_mlm:
            # mmb_ce_bwd's dense = 0 leaves the unlabelled rows of dlogits untouched while the decoder dgrad / wgrad GEMMs
            # read every row; the fused default already restricts the element-wise work to the labelled rows
            raise capi.MMBError("dense_mlm = False is not supported (the decoder GEMMs read every dlogits row); the default "
                                "fused cross entropy already touches only the labelled rows")
        # everything that is baked into a plan when it is built is part of its key
        p_joint = float(self.bert.jointEmbeddings.dropout.p)
        key = (B, T, Lv, La, bool(needs_grad), bool(self.training), fp32, p_joint, bool(self.dense_mlm),
               bool(self.materialize_logits),
               float(self.config.hidden_dropout_prob), float(self.config.attention_probs_dropout_prob))
        plan = self._plans.get(key)
        if plan is None and fp32:
            if len(self._plans) >= 4:
                self._plans.clear()
            p_any = max(float(self.config.hidden_dropout_prob), float(self.config.attention_probs_dropout_prob), p_joint)
            if self.training and p_any > 0:
                raise capi.MMBError("the fp32 validation path is dropout-free: call .eval() or set the dropout "
                                    "probabilities to 0 (the reference parity recipe)")
            from .engine_f32 import PlanF32
            plan = PlanF32(self.config, self.bert.dataset, self._store, B, T, Lv, La, device)
            self._plans[key] = plan
        if plan is None:
            if len(self._plans) >= 4:          # bound the activation memory held by stale shapes
                self._plans.clear()
            plan = Plan(self.config, self.bert.dataset, self._store, B, T, Lv, La, bool(needs_grad), device,
                        p_joint=p_joint, dense_mlm=self.dense_mlm, dropout=bool(self.training),
                        materialize_logits=self.materialize_logits)
            self._plans[key] = plan
            plan._frame_sig = None
        return plan

    # ------------------------------------------------------------------ forward
    def forward(self, input_ids, token_type_ids, attention_mask, masked_labels, ap_label, sentiment):
        """Same contract as MMBertForPretraining.forward (MMBertForPretraining.py:392-449): returns
        ``((joint_loss, None, None, None, ap_loss, label_loss, nce, pred_t, rel_t, pred_v, align_v, pred_s, align_s),
        logits)``.  ``joint_loss`` is differentiable (``.backward()`` fills every parameter's ``.grad``); the other
        outputs are plain tensors.  ``pred_*`` are bf16 views of the decoder output of the packed batch when
        ``self.materialize_logits`` is set, else None (fused cross entropy, the default)."""
        dev = self.bert.embeddings.word_embeddings.weight.device
        if dev.type != "cuda":
            raise capi.MMBError("MMBertForPretraining.forward needs the model on a CUDA (sm_100) device: "
                                "there is no CPU path (call .cuda())")
        capi.check(capi.lib().mmb_check_device(), "mmb_check_device")
        store = self._ensure_store(dev)
        ids_t, vis, aud = input_ids[0], input_ids[1], input_ids[2]
        B, T = ids_t.shape
        Lv, La = vis.shape[1], aud.shape[1]
        fp32 = self.precision == "fp32"
        needs_grad = torch.is_grad_enabled() and not fp32
        classify = self.num_labels not in (1, 7)
        if classify:
            # MMBertForPretraining.py:437-442: CrossEntropyLoss()(logits [B, 1], sentiment).  classifier1_2 is always
            # Linear(H, 1) (:311-314 — num_labels is 7 when it is built), so this is a ONE-class problem: torch accepts
            # integer class indices only here (a float target must have the logits' shape) and every index must be 0.
            if sentiment.is_floating_point() or sentiment.dim() != 1:
                raise capi.MMBError("num_labels not in (1, 7) selects the reference's CrossEntropyLoss branch "
                                    "(MMBertForPretraining.py:437-442): sentiment must be an integer class-index tensor "
                                    "of shape [B] (torch raises for a floating-point [B] target against [B, 1] logits)")
        plan = self._plan(B, T, Lv, La, dev, needs_grad)
        plan.set_loss_weights(self.alpha, self.beta, self.num_labels)
        graph = self.use_cuda_graph and not needs_grad and not self.training and not fp32
        if not graph:
            self._keep = plan.bind_inputs(input_ids, token_type_ids, attention_mask, masked_labels, ap_label, sentiment)
        sig = self._signature()
        store.refresh_bf16(sig)
        if plan._frame_sig != sig:
            plan.refresh_frame_weights()
            plan._frame_sig = sig
        if self.training:
            plan.set_seed(int(torch.randint(0, 2 ** 62, (1,)).item()))
        if needs_grad:
            anchor = self._params["classifier1_2.bias"]
            joint = _StepFn.apply(anchor, self, plan)
        elif graph:
            self.launches += plan.forward_graph(input_ids, token_type_ids, attention_mask, masked_labels, ap_label, sentiment)
            joint = None
        else:                                   # no_grad, or the forward-only fp32 validation path
            Plan.run(plan.fwd)
            self.launches += len(plan.fwd)
            joint = None
        # The plan's buffers are overwritten by the next step: every small result is handed out as a fresh copy (one
        # device copy), so callers may accumulate them across batches as trainer.py:165-180 does.  The pred_* outputs
        # stay views of the packed decoder output (up to 4.5 GB): valid until the next forward.
        small = plan.small_out.clone()
        losses = small[:8]
        if joint is None:
            joint = losses[0]
        V = self.config.vocab_size
        b1, b2 = B * T, B * T + B * (T + Lv)
        if plan.logits is None:                 # fused cross entropy (materialize_logits = False)
            pred_t = pred_v = pred_s = None
        else:
            pred_t = plan.logits[:b1, :V].view(B, T, V)
            pred_v = plan.logits[b1:b2, :V].view(B, T + Lv, V)
            pred_s = plan.logits[b2:, :V].view(B, T + La, V)
        rel = small[8 + B:8 + 3 * B].view(B, 2)
        align = small[8 + 3 * B:].view(2 * B, 2)
        outputs = (joint, None, None, None, losses[2], losses[3], losses[4], pred_t, rel, pred_v,
                   align[:B], pred_s, align[B:])
        logits = small[8:8 + B]
        if classify:
            return outputs, logits.long()       # torch.argmax(sigmoid(logits [B, 1]), dim=1): all zeros
        return outputs, logits.view(B, 1)
