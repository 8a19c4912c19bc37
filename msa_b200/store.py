"""Flat parameter / gradient storage for MMBertForPretraining.

All parameters live in ONE fp32 device buffer (and all gradients in one more, plus a bf16 mirror that the
tcgen05 GEMMs read), laid out so that
  * the per-layer query/key/value weights (and biases) are adjacent -> the fused QKV GEMM and its wgrad
    read/write them as one [3H, H] matrix while ``named_parameters()`` still exposes the reference's three;
  * the buffer is split [weight-decay | no-decay | untouched] following train.py:77-91's substring rule
    ('bias', 'LayerNorm.bias', 'LayerNorm.weight' -> no decay), so the fused AdamW is two launches and the
    data-parallel all-reduce covers one contiguous range;
  * parameters the path never touches (W_cv, W_cs, cls.seq_relationship) sit at the end, are
    never all-reduced or stepped, and keep ``.grad is None`` exactly as in the reference.
The nn.Parameters of the module tree are views into these buffers (``p.data`` / ``p.grad``), so stock
optimizers, ``state_dict()`` and ``load_state_dict()`` keep working.
"""
from collections import OrderedDict

import torch

from .params import NO_GRAD, param_shapes

ALIGN = 64  # elements; keeps every tensor 256-byte (fp32) / 128-byte (bf16) aligned for TMA and vector access


def _round_up(n, a=ALIGN):
    return (n + a - 1) // a * a


def _is_no_decay(name):
    return "bias" in name or "LayerNorm.weight" in name


class FlatStore:
    def __init__(self, cfg, dataset):
        shapes = param_shapes(cfg, dataset)
        self.shapes = shapes
        decay = [n for n in shapes if n not in NO_GRAD and not _is_no_decay(n)]
        nodecay = [n for n in shapes if n not in NO_GRAD and _is_no_decay(n)]
        frozen = [n for n in shapes if n in NO_GRAD]
        self.offsets = OrderedDict()
        off = 0
        for group in (decay, nodecay, frozen):
            for n in group:           # param_shapes order keeps query/key/value adjacent inside a group
                self.offsets[n] = off
                numel = 1
                for d in shapes[n]:
                    numel *= d
                off += _round_up(numel)
            if group is decay:
                self.decay_end = off
            elif group is nodecay:
                self.trainable_end = off
        self.total = off
        # q/k/v adjacency requires numel(H*H) and numel(H) to be multiples of ALIGN
        H = cfg.hidden_size
        assert H % ALIGN == 0, "hidden size must be a multiple of 64"
        self.flat = None
        self.grad = None
        self.bf16 = None
        self._bf16_version = -1

    def materialize(self, params, device):
        """(Re)builds the flat buffers on ``device`` from the current values of ``params`` (name -> Parameter)
        and re-points every Parameter's storage into them."""
        flat = torch.zeros(self.total, device=device, dtype=torch.float32)
        for n, off in self.offsets.items():
            p = params[n]
            flat[off:off + p.numel()].copy_(p.detach().reshape(-1).to(device=device, dtype=torch.float32))
        self.flat = flat
        self.grad = torch.zeros(self.total, device=device, dtype=torch.float32)
        self.bf16 = torch.empty(self.total, device=device, dtype=torch.bfloat16)
        self._bf16_version = -1
        for n, off in self.offsets.items():
            p = params[n]
            p.data = flat[off:off + p.numel()].view(self.shapes[n])
            p.grad = None

    def is_current(self, params, device):
        if self.flat is None or self.flat.device != device:
            return False
        for n in ("bert.embeddings.word_embeddings.weight", "classifier1_2.bias", "cpc_za.net.bias"):
            if params[n].data_ptr() != self.flat.data_ptr() + 4 * self.offsets[n]:
                return False
        return True

    def view(self, name, buf=None):
        buf = self.flat if buf is None else buf
        off = self.offsets[name]
        numel = 1
        for d in self.shapes[name]:
            numel *= d
        return buf[off:off + numel].view(self.shapes[name])

    def span(self, first, last, buf=None):
        """1-D view covering parameters ``first`` .. ``last`` (adjacent in the buffer, padding included)."""
        buf = self.flat if buf is None else buf
        a = self.offsets[first]
        numel = 1
        for d in self.shapes[last]:
            numel *= d
        return buf[a:self.offsets[last] + numel]

    def refresh_bf16(self, signature):
        """Re-casts the bf16 GEMM operand mirror if ``signature`` (the sum of the Parameters' version counters:
        every in-place update by an optimizer, ``load_state_dict`` ... bumps one) changed since the last cast.
        The fused AdamW writes the mirror itself and calls ``mark_bf16_fresh``."""
        from . import capi
        if signature != self._bf16_version:
            capi.cast_bf16(self.flat, self.bf16)
            self._bf16_version = signature
            return True
        return False

    def mark_bf16_fresh(self, signature):
        self._bf16_version = signature
