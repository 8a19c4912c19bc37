"""Fused AdamW over the flat parameter buffer, with the semantics of the optimizer the reference builds
(train.py:76-92: ``transformers.AdamW(grouped_parameters, lr)`` — transformers <= 4.x — i.e. betas (0.9, 0.999),
eps 1e-6, correct_bias=True, weight decay 0.01 on everything except names containing 'bias' /
'LayerNorm.weight', applied after the Adam update; parameters whose grad is None are skipped).

It is a ``torch.optim.Optimizer`` so LR schedulers (train.py:93-97: get_linear_schedule_with_warmup) drive
``param_groups[i]['lr']`` as usual; ``step()`` is two launches of mmb_adamw (decay / no-decay range), each of
which also refreshes the bf16 GEMM operand mirror in the same pass.
"""
import torch

from . import capi


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr=5e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01, correct_bias=True):
        if model._store is None or model._store.flat is None:
            dev = next(model.parameters()).device
            model._ensure_store(dev)
        self.model = model
        st = model._store
        names = list(st.offsets)
        decay = [model._params[n] for n in names if st.offsets[n] < st.decay_end]
        nodecay = [model._params[n] for n in names if st.decay_end <= st.offsets[n] < st.trainable_end]
        groups = [dict(params=decay, weight_decay=weight_decay), dict(params=nodecay, weight_decay=0.0)]
        super().__init__(groups, dict(lr=lr, betas=betas, eps=eps, correct_bias=correct_bias))
        self._ranges = [(0, st.decay_end), (st.decay_end, st.trainable_end)]
        self._m = torch.zeros(st.trainable_end, device=st.flat.device, dtype=torch.float32)
        self._v = torch.zeros_like(self._m)
        self._step = 0
        self._flat_ptr = st.flat.data_ptr()     # the buffers the moments belong to (checked on every step)
        self.grad_scale = 1.0      # 1/world_size after a SUM all-reduce
        self._reducer = None       # msa_b200.ddp.GradReducer.pipeline_optimizer: pieces of the all-reduce still in flight

    @torch.no_grad()
    def step(self, closure=None):
        st = self.model._store
        if st is None or st.flat is None or st.flat.data_ptr() != self._flat_ptr or st.flat.device != self._m.device:
            raise capi.MMBError("model storage was re-materialised after the optimizer was built (build FusedAdamW "
                                "after .cuda() / set_joint_embeddings)")
        self._step += 1
        # slices to update, in order: the whole buffer, or the pieces of a pipelined all-reduce (each waited for right
        # before its update, so the update of piece k runs while pieces k+1.. are still being exchanged)
        red = self._reducer
        pieces = [(0, st.trainable_end, None)]
        if red is not None and red.inflight:
            pieces, red.inflight = red.inflight, []
        launches = 0
        for lo, hi, work in pieces:
            if work is not None:
                work.wait()
            for (a, b), group in zip(self._ranges, self.param_groups):
                a, b = max(a, lo), min(b, hi)
                if a >= b:
                    continue
                args = capi.fill(capi.AdamwArgs(), p=st.flat[a:b], g=st.grad[a:b], m=self._m[a:b], v=self._v[a:b],
                                 p_bf16=st.bf16[a:b], n=b - a, lr=group["lr"], beta1=group["betas"][0],
                                 beta2=group["betas"][1], eps=group["eps"], weight_decay=group["weight_decay"],
                                 grad_scale=self.grad_scale, step=self._step, correct_bias=int(group["correct_bias"]))
                capi.call("adamw", args)
                launches += 1
        self.model.launches += launches
        # the kernel wrote parameters through raw pointers: version counters did not move, the mirror is fresh.
        # the frame-projection transposes are rebuilt by the next forward:
        for plan in self.model._plans.values():
            plan._frame_sig = None
        return None

    def state_dict(self):
        """torch's param_groups plus the fused state: step count and the two flat moment buffers (the per-parameter
        ``state`` of a stock optimizer is empty here — the moments live in two flat tensors)."""
        sd = super().state_dict()
        sd["fused"] = {"step": self._step, "exp_avg": self._m.clone(), "exp_avg_sq": self._v.clone(), "grad_scale": self.grad_scale}
        return sd

    def load_state_dict(self, state_dict):
        fused = state_dict.get("fused")
        super().load_state_dict({k: v for k, v in state_dict.items() if k != "fused"})
        if fused is not None:
            if fused["exp_avg"].numel() != self._m.numel():
                raise capi.MMBError("optimizer state belongs to a model of another shape")
            self._step = int(fused["step"])
            self._m.copy_(fused["exp_avg"])
            self._v.copy_(fused["exp_avg_sq"])

    def zero_grad(self, set_to_none=False):
        """Zeroes the flat gradient buffer with one memset and keeps the ``.grad`` views attached."""
        st = self.model._store
        if self._reducer is not None:
            self._reducer.wait_all()        # pieces of a pipelined all-reduce nobody consumed
        st.grad[:st.trainable_end].zero_()
        if set_to_none:
            for p in self.model._trainable:
                p.grad = None
