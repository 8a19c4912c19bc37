"""Launch plan for one packed MMBert step.

A ``Plan`` is built once per (config, batch shape, mode): it owns every activation / scratch buffer of the
step and the pre-filled C argument structures of every kernel launch, so a forward or backward is a tight
loop of ``fn(args, stream)`` calls over the C ABI with no allocation and no host<->device synchronisation.
The call sequence restates MMBertForPretraining.forward (MMBertForPretraining.py:392-449) on the packed
3-pass batch (see csrc/embed.cu for the row order) and its autograd backward.

torch is used for device memory (``torch.empty``) and the stream handle only.
"""
import ctypes
import math
import os

import torch

from . import capi
from .synth import DATASET_DIMS

BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32

# dropout RNG streams (per launch site; the layer index is added in bits 8+)
ST_ATTN, ST_OUT1, ST_OUT2 = 1, 2, 3


def _split_k(M_out, N_out, K, clusters=74):
    """Split-K factor for a wgrad GEMM (small M_out x N_out output, K = token count) on the CTA-pair kernel: 256 x 256
    output tiles, one work unit = (tile, K-slice), `clusters` CTA pairs walking the units round-robin.  Picks the factor
    whose unit count fills whole waves best (e.g. 36 tiles x 2 = 72 units = one wave of 74 pairs at 97 %, where the old
    "2 x SMs / tiles" rule gave 36 x 5 = 180 units = 2.43 waves at 81 %); ties go to the smaller factor (fewer fp32
    reductions).  Every slice keeps at least 8 k-blocks of 64."""
    tiles = ((M_out + 255) // 256) * ((N_out + 255) // 256)
    kb = (K + 63) // 64
    best, best_eff = 1, 0.0
    for sk in range(1, max(1, min(32, kb // 8)) + 1):
        units = tiles * sk
        eff = units / (((units + clusters - 1) // clusters) * clusters)
        if eff >= 0.95:
            return sk                      # the smallest factor that fills its waves
        if eff > best_eff:
            best, best_eff = sk, eff
    return best


class Plan:
    def __init__(self, cfg, dataset, store, B, T, Lv, La, training, device, p_joint=0.5, dense_mlm=True, dropout=None,
                 materialize_logits=False):
        # ``training``: keep the activations and build the backward plan.  ``dropout`` (default: same as training):
        # apply the dropout probabilities — a forward with autograd enabled on an eval() model keeps the
        # activations but drops nothing, exactly what the reference's modules would do.
        dropout = training if dropout is None else dropout
        self.cfg, self.store, self.training, self.device = cfg, store, training, device
        self.B, self.T, self.Lv, self.La = B, T, Lv, La
        self.Dv, self.Da = DATASET_DIMS[dataset]
        H, I, V, N = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.num_hidden_layers
        nh = cfg.num_attention_heads
        if H != nh * 64:
            raise capi.MMBError(f"head dim must be 64 (hidden {H}, heads {nh})")
        self.H, self.I, self.V, self.N, self.nh = H, I, V, N, nh
        self.Vp = (V + 7) // 8 * 8
        self.M = M = B * (3 * T + Lv + La)
        self.nfr = nfr = B * (Lv + La)
        self.max_S = T + max(Lv, La)
        self.p_hidden = float(cfg.hidden_dropout_prob) if dropout else 0.0
        self.p_attn = float(cfg.attention_probs_dropout_prob) if dropout else 0.0
        self.p_joint = float(p_joint) if dropout else 0.0
        self.dense_mlm = dense_mlm
        # False (default): the tied-decoder GEMM is fused with the cross entropy (MMB_EPI_CE_STATS) and the [M, V] logits are
        # never written; True: the reference's pred_t / pred_v / pred_s are produced (parity tests, callers that read them)
        self.materialize_logits = bool(materialize_logits)
        self.alpha, self.beta, self.num_labels = 1.0, 1.0, 7
        dev = device

        def buf(*shape, dtype=BF16):
            return torch.empty(*shape, device=dev, dtype=dtype)

        # ---- packed metadata
        self.keybias, self.cu = buf(M, dtype=F32), buf(3 * B + 1, dtype=I32)
        self.label_count = buf(4, dtype=I32)
        self.row_label = buf(M, dtype=I32)
        self.kv_end = buf(3 * B, dtype=I32)     # per sequence: 1 + last unmasked key (attention skips the masked tail)
        # work lists of the persistent attention kernels (longest items first), rebuilt on the device for every batch
        # and shared by all layers; MMB_ATTN_SCHED=0 leaves the items in index order (A/B runs)
        self.attn_work = None
        if os.environ.get("MMB_ATTN_SCHED", "1") != "0" and 3 * B <= 8192:
            self.attn_work = capi.attn_schedule_buffer(3 * B, self.nh, self.max_S, dev)
        # ---- activations (kept for backward when training; one reused set in eval)
        nsets = N if training else 1
        self.layers = []
        # Forward attention skips the query tiles that lie entirely behind kv_end (padding rows nothing observable reads)
        # once the schedule has verified that no row behind kv_end carries a label: plans with the fused cross entropy only,
        # training or not — with materialised logits pred_t / pred_v / pred_s expose those positions.
        # MMB_ATTN_FWD_QSKIP=0 for A/B runs.
        self.attn_fwd_skip = (not self.materialize_logits and self.attn_work is not None and
                              os.environ.get("MMB_ATTN_QSKIP", "1") != "0" and os.environ.get("MMB_ATTN_FWD_QSKIP", "1") != "0")

        # The row kernels (LayerNorm forward / backward, column sums, the attention backward's preparation) take the
        # schedule's row list under the same premise and leave padding rows alone (include/mmbert_sm100.h:
        # mmb_attn_schedule_args.row_list).  MMB_ROW_SKIP=0 for A/B runs.
        self.row_list = None
        if self.attn_fwd_skip and os.environ.get("MMB_ROW_SKIP", "1") != "0":
            self.row_list = capi.row_list_buffer(M, dev)

        def zbuf(*shape, dtype=BF16):      # rows a skipped tile never writes are still GEMM operands: finite from the start
            return torch.zeros(*shape, device=dev, dtype=dtype) if self.attn_fwd_skip else buf(*shape, dtype=dtype)

        self.x = [zbuf(M, H) for _ in range(N + 1)] if training else [zbuf(M, H), zbuf(M, H)]
        # fp32 copies of the residual stream (see csrc/ln.cu: precision note); the last layer's is never read
        self.x32 = [zbuf(M, H, dtype=F32) for _ in range(N if training else 2)]
        for _ in range(nsets):
            self.layers.append(dict(
                qkv=zbuf(M, 3 * H), ctx=zbuf(M, H), lse=zbuf(nh, M, dtype=F32), y1=buf(M, H), a=zbuf(M, H),
                a32=zbuf(M, H, dtype=F32),
                m1=buf(M, dtype=F32), r1=buf(M, dtype=F32), u=zbuf(M, I), hg=zbuf(M, I), y2=buf(M, H),
                m2=buf(M, dtype=F32), r2=buf(M, dtype=F32)))
        self.e_m1, self.e_r1, self.e_m2, self.e_r2 = (buf(M, dtype=F32) for _ in range(4))
        self.pframe = zbuf(max(nfr, 1), H)
        # training: bf16 copies of the frames + padded scratch, so that the projection wgrad runs on the tensor cores
        ldf = [(d + 7) // 8 * 8 for d in (self.Dv, self.Da)]
        self.frames_bf16 = [zbuf(max(B * L, 1), l8) for L, l8 in zip((Lv, La), ldf)] if training else [None, None]
        self.gw_pad = [buf(H, l8, dtype=F32) for l8 in ldf] if training else [None, None]
        self.t_u, self.t_g, self.t_ln = buf(M, H), buf(M, H), zbuf(M, H)
        self.t_m, self.t_r = buf(M, dtype=F32), buf(M, dtype=F32)
        self.logits = buf(M, self.Vp) if self.materialize_logits else None
        self.ce_stats = None
        if not self.materialize_logits:
            # per-row (max, sum exp) records of the fused CE + label logit + row-written flag; zeroed once (the flag plane persists)
            self.ce_stats = torch.zeros(capi.ce_stats_floats(M, V), device=dev, dtype=F32)
        self.row_lse = buf(M, dtype=F32)
        self.ce_sum = buf(4, dtype=F32)
        self.heads_ws = torch.empty(capi.heads_workspace_bytes(B, H) // 4 + 16, device=dev, dtype=F32)
        # every small user-visible result lives in ONE buffer, so that forward() hands out fresh copies with a single
        # device-to-device copy (the plan's buffers are overwritten by the next step)
        self.small_out = buf(8 + 7 * B, dtype=F32)
        self.losses, self.logits_out = self.small_out[:8], self.small_out[8:8 + B]
        self.rel_out, self.align_out = self.small_out[8 + B:8 + 3 * B].view(B, 2), self.small_out[8 + 3 * B:].view(2 * B, 2)
        self.wT = [buf(self.Dv, H, dtype=F32), buf(self.Da, H, dtype=F32)]
        self.gscale = torch.ones(1, device=dev, dtype=F32)
        if training:
            # fused CE: only the labelled rows are ever written (logits by the GEMM epilogue, then dlogits in place by
            # mmb_ce_sparse_bwd); every other row stays zero from here on
            self.dlogits = buf(M, self.Vp) if self.materialize_logits else torch.zeros(M, self.Vp, device=dev, dtype=BF16)
            self.GA, self.GC = buf(M, H), buf(M, H)                          # bf16: dgrad outputs / dense-output grads
            self.GB, self.GD = buf(M, H, dtype=F32), buf(M, H, dtype=F32)    # fp32: residual-stream gradients
            self.GT = buf(M, H)                                              # bf16 scratch (LM-head d_tln, dCtx)
            self.dqkv, self.du = buf(M, 3 * H), buf(M, I)
            self.attn_ws = capi.attn_bwd_workspace(M, nh, dev)
            self.dpre = buf(max(nfr, 1), H)
        self._seeded = []      # arg structs carrying a dropout seed
        self._build_forward()
        if training:
            self._build_backward()

    # ------------------------------------------------------------------ helpers
    def _w(self, name, buf=None):
        return self.store.view(name, self.store.bf16 if buf is None else buf)

    def _p(self, name):
        return self.store.view(name)

    def _g(self, name):
        return self.store.view(name, self.store.grad)

    def _fn(self, name):
        return getattr(capi.lib(), "mmb_" + name)

    def _gemm(self, seq, A, B, C, M, N, K, **kw):
        seq.append((self._fn("gemm"), capi.gemm_args(A, B, C, M, N, K, **kw)))

    # ------------------------------------------------------------------ forward plan
    def _build_forward(self):
        c, H, I, M, N = self.cfg, self.H, self.I, self.M, self.N
        st = self.store
        f = []
        self.pack_args = capi.fill(capi.PackArgs(), keybias=self.keybias, cu_seqlens=self.cu,
                                   label_count=self.label_count, kv_end=self.kv_end, B=self.B, T=self.T, L=[self.Lv, self.La],
                                   frame_dim=[self.Dv, self.Da], row_label=self.row_label, vocab=self.V)
        f.append((self._fn("pack_prepare"), self.pack_args))
        if self.attn_work is not None:
            # row labels: lets the backward skip the query rows behind kv_end once it is verified that none is labelled
            # (their upstream gradient is exactly zero; include/mmbert_sm100.h).  MMB_ATTN_QSKIP=0 for A/B runs.
            qskip = (self.training or self.attn_fwd_skip) and os.environ.get("MMB_ATTN_QSKIP", "1") != "0"
            self.sched_args = capi.attn_schedule_args(self.cu, self.kv_end, self.attn_work, self.nh, self.max_S,
                                                      row_label=self.row_label if qskip else None,
                                                      row_list=self.row_list if qskip else None)
            if not qskip:
                self.row_list = None
            f.append((self._fn("attn_schedule"), self.sched_args))
        # per-row live flags behind the list: the GELU / multiply GEMM epilogues skip all-padding 32-row slices
        self.row_live = None
        if self.row_list is not None and os.environ.get("MMB_GEMM_ROW_SKIP", "1") != "0":
            self.row_live = self.row_list[4 + M:]
        je = "bert.jointEmbeddings."
        self.embed_args = capi.fill(
            capi.EmbedArgs(), frame_dim=[self.Dv, self.Da],
            word=self._p("bert.embeddings.word_embeddings.weight"),
            pos=self._p("bert.embeddings.position_embeddings.weight"),
            type=self._p("bert.embeddings.token_type_embeddings.weight"),
            ln1_g=self._p("bert.embeddings.LayerNorm.weight"), ln1_b=self._p("bert.embeddings.LayerNorm.bias"),
            ln2_g=self._p(je + "LayerNorm.weight"), ln2_b=self._p(je + "LayerNorm.bias"),
            wT=self.wT, wb=[self._p(je + "Wv.bias"), self._p(je + "Ws.bias")],
            eps1=c.layer_norm_eps, eps2=1e-5, p_drop1=self.p_hidden, p_drop2=self.p_joint, seed=0,
            x0=self.x[0], x0_f32=self.x32[0], mean1=self.e_m1, rstd1=self.e_r1, mean2=self.e_m2, rstd2=self.e_r2, pframe=self.pframe,
            B=self.B, T=self.T, L=[self.Lv, self.La], H=H, V=self.V, max_pos=c.max_position_embeddings,
            err_count=self.label_count[3:], row_live=self.row_live)
        if self.training:
            capi.fill(self.embed_args, dpre=self.dpre, frames_bf16=self.frames_bf16, gw_pad=self.gw_pad,
                      g_word=self._g("bert.embeddings.word_embeddings.weight"),
                      g_pos=self._g("bert.embeddings.position_embeddings.weight"),
                      g_type=self._g("bert.embeddings.token_type_embeddings.weight"),
                      g_ln1_g=self._g("bert.embeddings.LayerNorm.weight"), g_ln1_b=self._g("bert.embeddings.LayerNorm.bias"),
                      g_ln2_g=self._g(je + "LayerNorm.weight"), g_ln2_b=self._g(je + "LayerNorm.bias"),
                      g_w=[self._g(je + "Wv.weight"), self._g(je + "Ws.weight")],
                      g_wb=[self._g(je + "Wv.bias"), self._g(je + "Ws.bias")])
        self._seeded.append(self.embed_args)
        f.append((self._fn("embed_fwd"), self.embed_args))
        for l in range(N):
            L = self.layers[l if self.training else 0]
            xin = self.x[l] if self.training else self.x[l % 2]
            xout = self.x[l + 1] if self.training else self.x[(l + 1) % 2]
            xin32 = self.x32[l] if self.training else self.x32[l % 2]
            xout32 = None if l == N - 1 else (self.x32[l + 1] if self.training else self.x32[(l + 1) % 2])
            pre = f"bert.encoder.layer.{l}."
            wqkv = st.span(pre + "attention.self.query.weight", pre + "attention.self.value.weight", st.bf16).view(3 * H, H)
            bqkv = st.span(pre + "attention.self.query.bias", pre + "attention.self.value.bias")
            self._gemm(f, xin, wqkv, L["qkv"], M, 3 * H, H, bias=bqkv, row_live=self.row_live)
            a = capi.attn_args(L["qkv"], L["ctx"], L["lse"], self.keybias, self.cu, H, self.nh, self.max_S,
                               p_drop=self.p_attn, rng_stream=(l << 8) | ST_ATTN, kv_end=self.kv_end, work=self.attn_work,
                               flags=8 if self.attn_fwd_skip else 0, row_list=self.row_list)
            L["attn_args"] = a
            self._seeded.append(a)
            f.append((self._fn("attn_fwd"), a))
            self._gemm(f, L["ctx"], self._w(pre + "attention.output.dense.weight"), L["y1"], M, H, H,
                       bias=self._p(pre + "attention.output.dense.bias"), row_live=self.row_live)
            a = capi.drln_fwd_args(L["y1"], xin32, self._p(pre + "attention.output.LayerNorm.weight"),
                                   self._p(pre + "attention.output.LayerNorm.bias"), L["a"], L["m1"], L["r1"],
                                   c.layer_norm_eps, p_drop=self.p_hidden, rng_stream=(l << 8) | ST_OUT1,
                                   out_f32=L["a32"], row_list=self.row_list)
            self._seeded.append(a)
            f.append((self._fn("dropout_residual_ln_fwd"), a))
            # training: the epilogue also saves gelu'(u) (in the "u" buffer) so that the backward is a plain multiply
            self._gemm(f, L["a"], self._w(pre + "intermediate.dense.weight"), L["hg"], M, I, H,
                       epilogue=capi.EPI_GELU_GRAD_BF16 if self.training else capi.EPI_GELU_BF16,
                       aux=L["u"] if self.training else None, bias=self._p(pre + "intermediate.dense.bias"),
                       row_live=self.row_live)
            self._gemm(f, L["hg"], self._w(pre + "output.dense.weight"), L["y2"], M, H, I,
                       bias=self._p(pre + "output.dense.bias"), row_live=self.row_live)
            a = capi.drln_fwd_args(L["y2"], L["a32"], self._p(pre + "output.LayerNorm.weight"),
                                   self._p(pre + "output.LayerNorm.bias"), xout, L["m2"], L["r2"],
                                   c.layer_norm_eps, p_drop=self.p_hidden, rng_stream=(l << 8) | ST_OUT2,
                                   out_f32=xout32, row_list=self.row_list)
            self._seeded.append(a)
            f.append((self._fn("dropout_residual_ln_fwd"), a))
        self.seq_out = self.x[N] if self.training else self.x[N % 2]
        # LM head: decoder(LayerNorm(gelu(dense(seq)))) on ALL positions (MMBertForPretraining.py:293)
        tp = "cls.predictions.transform."
        self._gemm(f, self.seq_out, self._w(tp + "dense.weight"), self.t_g, M, H, H, epilogue=capi.EPI_GELU_BF16,
                   aux=self.t_u if self.training else None, bias=self._p(tp + "dense.bias"), row_live=self.row_live)
        f.append((self._fn("dropout_residual_ln_fwd"),
                  capi.drln_fwd_args(self.t_g, None, self._p(tp + "LayerNorm.weight"), self._p(tp + "LayerNorm.bias"),
                                     self.t_ln, self.t_m, self.t_r, c.layer_norm_eps, row_list=self.row_list)))
        word_bf = self._w("bert.embeddings.word_embeddings.weight")
        if self.materialize_logits:
            self._gemm(f, self.t_ln, word_bf, self.logits, M, self.V, H, bias=self._p("cls.predictions.bias"))
            self.ce_args = capi.fill(capi.CeArgs(), logits=self.logits, label_count=self.label_count, row_lse=self.row_lse,
                                     loss_sum=self.ce_sum, gscale=self.gscale, coef=self.alpha / 3.0, V=self.V,
                                     ldl=self.Vp, B=self.B, T=self.T, L=[self.Lv, self.La], dense=1 if self.dense_mlm else 0)
            if self.training:
                capi.fill(self.ce_args, dlogits=self.dlogits)
            f.append((self._fn("ce_fwd"), self.ce_args))
        else:
            # fused CE (MMBertForPretraining.py:293 + :381-384): the decoder GEMM keeps online softmax statistics of the
            # labelled rows and stores only those rows' logits (training: into the dlogits buffer)
            C = self.dlogits if self.training else None
            self._gemm(f, self.t_ln, word_bf, C, M, self.V, H, bias=self._p("cls.predictions.bias"),
                       epilogue=capi.EPI_CE_STATS, aux=self.row_label, aux2=self.ce_stats, ldc=self.Vp, ldaux=0)
            gb0 = self.store.offsets["cls.predictions.bias"]
            self.ce_args = capi.fill(capi.CeSparseArgs(), row_label=self.row_label, stats=self.ce_stats,
                                     label_count=self.label_count, row_lse=self.row_lse, loss_sum=self.ce_sum,
                                     gscale=self.gscale, coef=self.alpha / 3.0, V=self.V, ldl=self.Vp, B=self.B, T=self.T,
                                     L=[self.Lv, self.La])
            if self.training:
                capi.fill(self.ce_args, dlogits=self.dlogits, dbias=self.store.grad[gb0:gb0 + self.V])
            f.append((self._fn("ce_sparse_fwd"), self.ce_args))
        hp = dict(
            seq_out=self.seq_out, cu_seqlens=self.cu, workspace=self.heads_ws,
            w_pooler=self._p("bert.pooler.dense.weight"), b_pooler=self._p("bert.pooler.dense.bias"),
            w_seqrel=self._p("cls.seq_relationship.weight"), b_seqrel=self._p("cls.seq_relationship.bias"),
            w_align=self._p("cls.align.weight"), b_align=self._p("cls.align.bias"),
            w_attn=self._p("attn.weight"), b_attn=self._p("attn.bias"),
            w_c11=self._p("classifier1_1.weight"), b_c11=self._p("classifier1_1.bias"),
            w_c12=self._p("classifier1_2.weight"), b_c12=self._p("classifier1_2.bias"),
            w_v=[self._p(n + ".weight") for n in ("vt", "vv", "vs")], b_v=[self._p(n + ".bias") for n in ("vt", "vv", "vs")],
            w_cpc=[self._p(n + ".net.weight") for n in ("cpc_zt", "cpc_zv", "cpc_za")],
            b_cpc=[self._p(n + ".net.bias") for n in ("cpc_zt", "cpc_zv", "cpc_za")],
            ce_loss_sum=self.ce_sum, label_count=self.label_count, losses=self.losses, logits_out=self.logits_out,
            rel_out=self.rel_out, align_out=self.align_out, gscale=self.gscale, alpha=self.alpha, beta=self.beta,
            B=self.B, H=H, num_labels=self.num_labels)
        if self.training:
            hp.update(
                dseq_out=self.GA,
                g_w_pooler=self._g("bert.pooler.dense.weight"), g_b_pooler=self._g("bert.pooler.dense.bias"),
                g_w_align=self._g("cls.align.weight"), g_b_align=self._g("cls.align.bias"),
                g_w_attn=self._g("attn.weight"), g_b_attn=self._g("attn.bias"),
                g_w_c11=self._g("classifier1_1.weight"), g_b_c11=self._g("classifier1_1.bias"),
                g_w_c12=self._g("classifier1_2.weight"), g_b_c12=self._g("classifier1_2.bias"),
                g_w_v=[self._g(n + ".weight") for n in ("vt", "vv", "vs")],
                g_b_v=[self._g(n + ".bias") for n in ("vt", "vv", "vs")],
                g_w_cpc=[self._g(n + ".net.weight") for n in ("cpc_zt", "cpc_zv", "cpc_za")],
                g_b_cpc=[self._g(n + ".net.bias") for n in ("cpc_zt", "cpc_zv", "cpc_za")])
        self.heads_args = capi.fill(capi.HeadsArgs(), **hp)
        f.append((self._fn("heads_fwd"), self.heads_args))
        self.fwd = f

    # ------------------------------------------------------------------ backward plan
    def _build_backward(self):
        c, H, I, M, N, V = self.cfg, self.H, self.I, self.M, self.N, self.V
        st = self.store
        b = []
        MN, K_ = capi.MAJOR_MN, capi.MAJOR_K
        ATOM = capi.EPI_ATOMIC_ADD_F32
        word_bf = self._w("bert.embeddings.word_embeddings.weight")
        fuse_colsum = os.environ.get("MMB_GEMM_COLSUM", "1") != "0"      # A/B runs: 0 = separate mmb_colsum_bf16 launches
        b.append((self._fn("ce_bwd" if self.materialize_logits else "ce_sparse_bwd"), self.ce_args))
        # tied decoder: d_tln = dlogits · Wword ; g_word += dlogits^T · t_ln ; g_dec_bias += colsum(dlogits)
        self._gemm(b, self.dlogits, word_bf, self.GT, M, H, V, b_major=MN)
        self._gemm(b, self.dlogits, self.t_ln, self._g("bert.embeddings.word_embeddings.weight"), V, H, M,
                   a_major=MN, b_major=MN, epilogue=ATOM, split_k=_split_k(V, H, M))
        if self.materialize_logits:
            gbias = st.grad[st.offsets["cls.predictions.bias"]:st.offsets["cls.predictions.bias"] + self.Vp]
            b.append((self._fn("colsum_bf16"), capi.colsum_args(self.dlogits, gbias)))
        else:
            b.append((self._fn("colsum_rows_bf16"), self.ce_args))
        tp = "cls.predictions.transform."
        b.append((self._fn("dropout_residual_ln_bwd"),
                  capi.fill(capi.drln_bwd_args(self.GT, None, self.t_g, None, self.t_m, self.t_r,
                                               self._p(tp + "LayerNorm.weight"), self.GC, None,
                                               self._g(tp + "LayerNorm.weight"), self._g(tp + "LayerNorm.bias"),
                                               self._g(tp + "dense.bias"), row_list=self.row_list), gelu_aux=self.t_u)))
        self._gemm(b, self.GC, self._w(tp + "dense.weight"), self.GA, M, H, H, b_major=MN, row_live=self.row_live)
        self._gemm(b, self.GC, self.seq_out, self._g(tp + "dense.weight"), H, H, M, a_major=MN, b_major=MN,
                   epilogue=ATOM, split_k=_split_k(H, H, M))
        b.append((self._fn("heads_bwd"), self.heads_args))   # adds the [CLS]-row gradients into GA
        for l in range(N - 1, -1, -1):
            L = self.layers[l]
            pre = f"bert.encoder.layer.{l}."
            g2 = None if l == N - 1 else self.GB
            # padding rows of GC / GD / GB / du are zero-filled by their first writer of the step (the top layer's launches);
            # only these kernels write those buffers, so every later launch of the sweep leaves the rows alone
            zeroed = 1 if (l < N - 1 and self.row_list is not None) else 0
            a = capi.drln_bwd_args(self.GA, g2, L["y2"], L["a32"], L["m2"], L["r2"], self._p(pre + "output.LayerNorm.weight"),
                                   self.GC, self.GD, self._g(pre + "output.LayerNorm.weight"),
                                   self._g(pre + "output.LayerNorm.bias"), self._g(pre + "output.dense.bias"),
                                   p_drop=self.p_hidden, rng_stream=(l << 8) | ST_OUT2, row_list=self.row_list,
                                   dead_rows_zeroed=zeroed)
            self._seeded.append(a)
            b.append((self._fn("dropout_residual_ln_bwd"), a))
            # FFN2: du = (dY2 · W2) ∘ gelu'(u) ; gW2 += dY2^T · hg     (L["u"] holds gelu'(u), see the forward)
            # (the epilogue also takes the column sums of du = the FFN1 bias gradient, from its staging tiles)
            self._gemm(b, self.GC, self._w(pre + "output.dense.weight"), self.du, M, I, H, b_major=MN,
                       epilogue=capi.EPI_MUL_AUX_BF16, aux=L["u"],
                       colsum=self._g(pre + "intermediate.dense.bias") if fuse_colsum else None, row_live=self.row_live,
                       dead_rows_zeroed=zeroed if self.row_live is not None else 0)
            self._gemm(b, self.GC, L["hg"], self._g(pre + "output.dense.weight"), H, I, M, a_major=MN, b_major=MN,
                       epilogue=ATOM, split_k=_split_k(H, I, M))
            if not fuse_colsum:
                b.append((self._fn("colsum_bf16"), capi.colsum_args(self.du, self._g(pre + "intermediate.dense.bias"))))
            # FFN1: dA = du · W1 ; gW1 += du^T · a
            # (GA's padding rows are read by nobody: the LayerNorm / embedding backward skip them.  dCtx below is different:
            # the attention backward reads the padding rows that share a 128-row tile with live ones, so it is written in full)
            self._gemm(b, self.du, self._w(pre + "intermediate.dense.weight"), self.GA, M, H, I, b_major=MN,
                       row_live=self.row_live)
            self._gemm(b, self.du, L["a"], self._g(pre + "intermediate.dense.weight"), I, H, M, a_major=MN, b_major=MN,
                       epilogue=ATOM, split_k=_split_k(I, H, M))
            a = capi.drln_bwd_args(self.GA, self.GD, L["y1"], self.x32[l], L["m1"], L["r1"],
                                   self._p(pre + "attention.output.LayerNorm.weight"), self.GC, self.GB,
                                   self._g(pre + "attention.output.LayerNorm.weight"),
                                   self._g(pre + "attention.output.LayerNorm.bias"),
                                   self._g(pre + "attention.output.dense.bias"),
                                   p_drop=self.p_hidden, rng_stream=(l << 8) | ST_OUT1, row_list=self.row_list,
                                   dead_rows_zeroed=zeroed)
            self._seeded.append(a)
            b.append((self._fn("dropout_residual_ln_bwd"), a))
            # attention output projection: dCtx = dY1 · Wo ; gWo += dY1^T · ctx
            self._gemm(b, self.GC, self._w(pre + "attention.output.dense.weight"), self.GT, M, H, H, b_major=MN)
            self._gemm(b, self.GC, L["ctx"], self._g(pre + "attention.output.dense.weight"), H, H, M, a_major=MN,
                       b_major=MN, epilogue=ATOM, split_k=_split_k(H, H, M))
            capi.fill(L["attn_args"], dctx=self.GT, dqkv=self.dqkv, bwd_ws=self.attn_ws)
            if zeroed:          # dqkv: the tiles behind kv_end were zero-filled by the top layer's launch
                L["attn_args"].flags |= 16
            b.append((self._fn("attn_bwd"), L["attn_args"]))
            wqkv = st.span(pre + "attention.self.query.weight", pre + "attention.self.value.weight", st.bf16).view(3 * H, H)
            gwqkv = st.span(pre + "attention.self.query.weight", pre + "attention.self.value.weight", st.grad).view(3 * H, H)
            gbqkv = st.span(pre + "attention.self.query.bias", pre + "attention.self.value.bias", st.grad)
            b.append((self._fn("colsum_bf16"), capi.colsum_args(self.dqkv, gbqkv, row_list=self.row_list)))
            self._gemm(b, self.dqkv, wqkv, self.GA, M, H, 3 * H, b_major=MN, row_live=self.row_live)
            self._gemm(b, self.dqkv, self.x[l], gwqkv, 3 * H, H, M, a_major=MN, b_major=MN, epilogue=ATOM,
                       split_k=_split_k(3 * H, H, M))
        capi.fill(self.embed_args, dx0=self.GA, dx0b=self.GB)
        b.append((self._fn("embed_bwd"), self.embed_args))
        self.bwd = b

    # ------------------------------------------------------------------ per-step binding
    def convert_inputs(self, input_ids, token_type_ids, attention_mask, masked_labels, ap_label, sentiment):
        """The step's 16 input tensors on the plan's device, contiguous, in the dtypes the kernels read (dtypes of the
        masks and frames are kept as the reference's collate produces them and passed as dtype codes), shape-checked.
        Order: ids x3, token types, frames x2, text masks x3, frame masks x2, labels x3, ap labels x2, sentiment."""
        ids_t, vis, aud, ids_v, ids_s = input_ids
        m_t, (m_tv, m_v), (m_ts, m_s) = attention_mask
        lab_t, lab_v, lab_s = masked_labels

        def dev(t, dtype=None):
            if t.device != self.device or (dtype is not None and t.dtype != dtype) or not t.is_contiguous():
                t = t.to(device=self.device, dtype=dtype).contiguous()
            return t

        ids = [dev(ids_t, torch.int64), dev(ids_v, torch.int64), dev(ids_s, torch.int64)]
        tt = dev(token_type_ids[0], torch.int64)
        frames = [dev(vis), dev(aud)]
        mt = [dev(m_t), dev(m_tv), dev(m_ts)]
        mf = [dev(m_v), dev(m_s)]
        labs = [dev(lab_t, torch.int64), dev(lab_v, torch.int64), dev(lab_s, torch.int64)]
        ap = [dev(ap_label[0], torch.int64), dev(ap_label[1], torch.int64)]
        sent = dev(sentiment.reshape(-1), torch.float32)
        B, T = self.B, self.T
        if tuple(ids[0].shape) != (B, T) or tuple(frames[0].shape) != (B, self.Lv, self.Dv) or \
                tuple(frames[1].shape) != (B, self.La, self.Da):
            raise capi.MMBError("input shapes do not match the plan")
        if tuple(labs[1].shape) != (B, T + self.Lv) or tuple(labs[2].shape) != (B, T + self.La):
            raise capi.MMBError("masked_labels of the joint passes must have shape [B, T+L] "
                                "(the reference cats the text labels onto the frame half, trainer.py:50,53)")
        for t, L, D in zip(mf, (self.Lv, self.La), (self.Dv, self.Da)):
            # [B,L,D] as the reference's collate builds them (only feature 0 is read), or already [B,L] (:74-77)
            if tuple(t.shape) not in ((B, L, D), (B, L)):
                raise capi.MMBError(f"frame attention mask must be [B, L, D] or [B, L], got {tuple(t.shape)}")
        return ids + [tt] + frames + mt + mf + labs + ap + [sent]

    def bind_tensors(self, ts):
        """Points the input-consuming launches at the 16 tensors of ``convert_inputs``."""
        ids, tt, frames, mt, mf, labs, ap, sent = ts[0:3], ts[3], ts[4:6], ts[6:9], ts[9:11], ts[11:14], ts[14:16], ts[16]
        capi.fill(self.pack_args, mask_text=mt, mask_text_dtype=[capi.dtype_code(t) for t in mt],
                  mask_frame=mf, mask_frame_dtype=[capi.dtype_code(t) for t in mf],
                  mask_frame_stride=[0 if t.dim() == 3 else 1 for t in mf], labels=labs)
        capi.fill(self.embed_args, ids=ids, token_type=tt, frames=frames,
                  frames_dtype=[capi.dtype_code(t) for t in frames])
        if self.materialize_logits:
            capi.fill(self.ce_args, labels=labs)
        capi.fill(self.heads_args, ap_label=ap, sentiment=sent)

    def bind_inputs(self, input_ids, token_type_ids, attention_mask, masked_labels, ap_label, sentiment):
        """Points the input-consuming launches at this step's tensors (device, contiguous; dtypes as the
        reference's collate produces them).  Returns the list of tensors that must stay alive."""
        keep = self.convert_inputs(input_ids, token_type_ids, attention_mask, masked_labels, ap_label, sentiment)
        self.bind_tensors(keep)
        return keep

    # ------------------------------------------------------------------ CUDA-graph replay (forward-only, dropout-free plans)
    def forward_graph(self, input_ids, token_type_ids, attention_mask, masked_labels, ap_label, sentiment):
        """Replays the whole forward plan as ONE CUDA graph (the C5 inference sweep at small batches is launch-bound:
        ~190 launches for a 12-layer forward).  The graph reads fixed input buffers owned by the plan; each call copies
        the step's tensors into them and replays.  Everything baked into the kernel arguments at capture time (loss
        weights, input dtypes / layouts) is part of the capture key; a change re-captures."""
        if self.training or max(self.p_hidden, self.p_attn, self.p_joint) > 0:
            raise capi.MMBError("CUDA-graph replay is for forward-only, dropout-free plans (dropout seeds are kernel arguments)")
        conv = self.convert_inputs(input_ids, token_type_ids, attention_mask, masked_labels, ap_label, sentiment)
        key = (tuple((t.dtype, tuple(t.shape)) for t in conv), self.alpha, self.beta, self.num_labels)
        if getattr(self, "_graph_key", None) != key:
            self._graph_static = [torch.empty_like(t) for t in conv]
            self.bind_tensors(self._graph_static)
            for s, t in zip(self._graph_static, conv):
                s.copy_(t)
            Plan.run(self.fwd)                      # warm-up outside capture: function attributes, TMA descriptor cache
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                Plan.run(self.fwd)
            self._graph, self._graph_key = g, key
        else:
            for s, t in zip(self._graph_static, conv):
                s.copy_(t, non_blocking=True)
        self._graph.replay()
        return len(self.fwd)

    def set_seed(self, seed):
        for a in self._seeded:
            a.seed = seed

    def set_loss_weights(self, alpha, beta, num_labels):
        self.alpha, self.beta, self.num_labels = float(alpha), float(beta), int(num_labels)
        self.ce_args.coef = self.alpha / 3.0
        self.heads_args.alpha, self.heads_args.beta, self.heads_args.num_labels = self.alpha, self.beta, self.num_labels

    def refresh_frame_weights(self):
        je = "bert.jointEmbeddings."
        H = self.H
        capi.transpose_f32(self._p(je + "Wv.weight"), self.wT[0], H, self.Dv)
        capi.transpose_f32(self._p(je + "Ws.weight"), self.wT[1], H, self.Da)

    @staticmethod
    def run(seq, hooks=None):
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        lib = capi.lib()
        for i, (fn, args) in enumerate(seq):
            rc = fn(ctypes.byref(args), stream)
            if rc != 0:
                raise capi.MMBError(f"launch {i} failed ({rc}): {lib.mmb_last_error().decode()}")
            if hooks is not None and i in hooks:
                hooks[i]()
        return len(seq)
