"""Data-parallel gradient exchange for the MMBert step: one process per GPU, one SUM
all-reduce of the gradient per step over NCCL (NVLink 5 / NVSwitch), bucketed and overlapped with backward.

The reference has no distributed code.  Semantics are PyTorch-DDP semantics: every rank runs the packed step
on its own micro-batch (per-rank CE means, per-rank in-batch CPC negatives) and the parameter gradients are
averaged; the 1/world factor is folded into the fused AdamW (``FusedAdamW.grad_scale``).

Two schedules (``mode``):

* ``"deferred"`` (default): ONE all-reduce of the whole flat gradient buffer right after the backward sweep.  Every
  compute kernel of this path is persistent — one CTA (or CTA pair) per SM with the whole register file and ~220 KB of
  shared memory, its tiles divided statically over the grid — so an NCCL kernel running under the backward cannot
  share an SM with them: it either waits for a kernel boundary or takes SMs away from the next persistent grid, whose
  late CTAs then stretch that kernel.  Measured at 2 GPUs (MOSEI shape): overlapped buckets cost +2.5 ms per 59.5 ms
  step (with or without SMs reserved for NCCL), while 460 MB over NVLink 5 with every SM free for NCCL is well under
  that.  ``compress="bf16"`` sends a bf16 copy (half the bytes; gradients are summed in bf16 by NCCL).
* ``"overlap"``: bucketed all-reduces issued from inside the backward launch loop.  The three reference passes are
  packed into ONE backward sweep, so layer l's weight gradients are final the moment its last wgrad GEMM is enqueued —
  that launch index triggers the bucket's all-reduce on NCCL's stream while the compute stream continues with layer
  l-1.  The wgrad epilogues accumulate straight into the flat gradient buffer (msa_b200.store.FlatStore), so a bucket
  is a contiguous slice of that buffer: nothing is copied or re-packed before the collective.

Params without gradient (W_cv, W_cs, cls.seq_relationship) live outside [0, trainable_end) and are never sent.
"""
import contextlib
import os

import torch
import torch.distributed as dist


def bucket_schedule(store, num_layers):
    """Returns [(trigger, [(start, end), ...]), ...] in backward order.  trigger is 'heads', ('layer', l) or
    'final'; the ranges are element offsets into the flat gradient buffer and partition [0, trainable_end)."""
    off = store.offsets
    first_layer = off["bert.encoder.layer.0.attention.self.query.weight"]
    pooler = off["bert.pooler.dense.weight"]
    transform = off["cls.predictions.transform.dense.weight"]
    sched = [("heads", [(transform, store.decay_end)])]
    for l in range(num_layers - 1, -1, -1):
        a = off[f"bert.encoder.layer.{l}.attention.self.query.weight"]
        b = off[f"bert.encoder.layer.{l + 1}.attention.self.query.weight"] if l + 1 < num_layers else pooler
        sched.append((("layer", l), [(a, b)]))
    # embeddings (tied decoder gradient + scatter-add finish last), pooler + frame projections, all biases / LayerNorms
    sched.append(("final", [(0, first_layer), (pooler, transform), (store.decay_end, store.trainable_end)]))
    return sched


def check_partition(sched, trainable_end):
    """True when the scheduled ranges tile [0, trainable_end) exactly once."""
    ranges = sorted(r for _, rs in sched for r in rs)
    pos = 0
    for a, b in ranges:
        if a != pos or b <= a:
            return False
        pos = b
    return pos == trainable_end


class GradReducer:
    """Issues the bucketed all-reduces.  ``attach(model)`` wires it into MMBertForPretraining's backward."""

    def __init__(self, store, num_layers, process_group=None, mode=None, compress=None):
        self.store = store
        self.group = process_group
        self.mode = mode or os.environ.get("MMB_DP_MODE", "deferred")
        self.compress = compress if compress is not None else (os.environ.get("MMB_DP_COMPRESS") or None)
        if self.mode not in ("deferred", "overlap") or self.compress not in (None, "bf16"):
            raise ValueError(f"GradReducer: mode={self.mode!r} compress={self.compress!r}")
        self._cbuf = None
        # deferred mode, optional: the all-reduce goes out in ``chunks`` pieces that nobody waits for at the end of the
        # backward; the fused AdamW then waits for piece k and updates that slice while pieces k+1.. are still in flight
        # (pipeline_optimizer).  Until optimizer.step() — or wait_all() — the gradient buffer is NOT reduced.
        self.chunks = 1
        self.inflight = []
        self.sched = bucket_schedule(store, num_layers)
        assert check_partition(self.sched, store.trainable_end)
        self.pending = []
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.bytes_per_step = 4 * store.trainable_end
        # False = accumulate locally (DistributedDataParallel.no_sync semantics).  The all-reduce is an in-place SUM over
        # the flat gradient buffer that backward also accumulates into, so with gradient accumulation ONLY the backward
        # that precedes optimizer.step() may reduce: reducing an earlier micro-batch and then again with the next one
        # would weight it world_size times (AR(AR(g1) + g2) = world * AR(g1) + AR(g2)).
        self.sync = True

    @contextlib.contextmanager
    def no_sync(self):
        old, self.sync = self.sync, False
        try:
            yield
        finally:
            self.sync = old

    def reduce_bucket(self, ranges):
        if self.world == 1 or not self.sync or self.mode != "overlap":
            return
        for a, b in ranges:
            self.pending.append(dist.all_reduce(self.store.grad[a:b], op=dist.ReduceOp.SUM, group=self.group,
                                                async_op=True))

    def pipeline_optimizer(self, optimizer, chunks=4):
        """Lets ``optimizer`` (msa_b200.optim.FusedAdamW) consume the deferred all-reduce piece by piece: its step() waits
        for one piece at a time, so the parameter update runs under the rest of the exchange.  Anything else that reads
        gradients between backward and step must call ``wait_all()`` first."""
        if self.mode != "deferred" or self.compress is not None:
            raise ValueError("pipeline_optimizer needs mode='deferred' without compression")
        self.chunks = max(1, int(chunks))
        optimizer._reducer = self
        return self

    def wait_all(self):
        for _, _, w in self.inflight:
            w.wait()
        self.inflight = []

    def finish(self):
        """End of the backward sweep.  deferred: the one all-reduce of the step; both modes: the current stream then
        waits for every outstanding collective (no host synchronisation on NCCL)."""
        if self.mode == "deferred" and self.world > 1 and self.sync and self.chunks > 1:
            n = self.store.trainable_end
            step = ((n + self.chunks - 1) // self.chunks + 63) // 64 * 64      # 256-byte aligned slices
            self.wait_all()
            for a in range(0, n, step):
                b = min(n, a + step)
                self.inflight.append((a, b, dist.all_reduce(self.store.grad[a:b], op=dist.ReduceOp.SUM, group=self.group,
                                                            async_op=True)))
            return
        if self.mode == "deferred" and self.world > 1 and self.sync:
            g = self.store.grad[:self.store.trainable_end]
            if self.compress == "bf16":
                from . import capi
                if self._cbuf is None:
                    self._cbuf = torch.empty(g.numel(), device=g.device, dtype=torch.bfloat16)
                capi.cast_bf16(g, self._cbuf)
                dist.all_reduce(self._cbuf, op=dist.ReduceOp.SUM, group=self.group, async_op=True).wait()
                g.copy_(self._cbuf)
            else:
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group, async_op=True).wait()
        for w in self.pending:
            w.wait()
        self.pending = []

    def hooks_for(self, plan):
        """launch index in plan.bwd -> callable.  Triggers: the heads_bwd launch, each layer's last wgrad launch."""
        if self.mode == "deferred":
            return None
        hooks = {}
        lib_heads = plan._fn("heads_bwd")
        gemm = plan._fn("gemm")
        heads_idx = next(i for i, (fn, _) in enumerate(plan.bwd) if fn is lib_heads)
        by_trigger = dict((t, r) for t, r in self.sched)
        hooks[heads_idx] = lambda r=by_trigger["heads"]: self.reduce_bucket(r)
        # each layer's backward ends with the fused-QKV wgrad GEMM: C == grad view of that layer's query weight
        for l in range(plan.N):
            gq = self.store.view(f"bert.encoder.layer.{l}.attention.self.query.weight", self.store.grad).data_ptr()
            idx = next(i for i, (fn, a) in enumerate(plan.bwd) if fn is gemm and a.C == gq)
            hooks[idx] = lambda r=by_trigger[("layer", l)]: self.reduce_bucket(r)
        hooks[len(plan.bwd) - 1] = lambda r=by_trigger["final"]: self.reduce_bucket(r)
        return hooks

    def attach(self, model):
        cache = {}

        def hooks(plan):
            if id(plan) not in cache:
                cache[id(plan)] = self.hooks_for(plan)
            return cache[id(plan)]

        model._bwd_hooks = hooks
        model._post_backward = self.finish
        model._reducer = self
        return self


def broadcast_parameters(model, src=0, process_group=None):
    """Rank ``src``'s flat parameter buffer to every rank (one collective)."""
    if dist.is_initialized() and dist.get_world_size(process_group) > 1:
        dist.broadcast(model._store.flat, src=src, group=process_group)
        model._store._bf16_version = -1
