"""fp32 validation path: the launch plan of ONE forward of the packed MMBert step with fp32 storage and fp32 CUDA-core
arithmetic (include/mmbert_sm100.h: "fp32 validation path").

BASELINE.json asks for an fp32 path within 1e-4 relative of the reference on logits and loss; the tensor-core path
(engine.Plan) is bf16 by construction.  This plan runs the same call sequence — MMBertForPretraining.forward
(MMBertForPretraining.py:392-449) on the packed 3-pass batch — through mmb_embed_fwd (exact_frames), mmb_linear_f32,
mmb_attn_f32_fwd, mmb_dropout_residual_ln_fwd (y_f32), mmb_ce_fwd (logits_f32) and mmb_heads_fwd (seq_out_f32).
Forward only, dropout-free (the reference parity recipe: eval() or p = 0), correctness-first; not the benchmarked path.
"""
import torch

from . import capi
from .engine import Plan
from .synth import DATASET_DIMS

F32, I32, BF16 = torch.float32, torch.int32, torch.bfloat16


class PlanF32:
    training = False

    def __init__(self, cfg, dataset, store, B, T, Lv, La, device):
        self.cfg, self.store, self.device = cfg, store, device
        self.B, self.T, self.Lv, self.La = B, T, Lv, La
        self.Dv, self.Da = DATASET_DIMS[dataset]
        H, I, V, N = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.num_hidden_layers
        nh = cfg.num_attention_heads
        if H != nh * 64:
            raise capi.MMBError(f"head dim must be 64 (hidden {H}, heads {nh})")
        self.H, self.I, self.V, self.N, self.nh = H, I, V, N, nh
        self.M = M = B * (3 * T + Lv + La)
        nfr = B * (Lv + La)
        self.max_S = T + max(Lv, La)
        self.alpha, self.beta, self.num_labels = 1.0, 1.0, 7
        self.materialize_logits = True                    # the validation path always produces pred_t / pred_v / pred_s

        def buf(*shape, dtype=F32):
            return torch.empty(*shape, device=device, dtype=dtype)

        self.keybias, self.cu = buf(M), buf(3 * B + 1, dtype=I32)
        self.label_count, self.kv_end = buf(4, dtype=I32), buf(3 * B, dtype=I32)
        self.x = [buf(M, H), buf(M, H)]
        self.x_bf16 = buf(M, H, dtype=BF16)               # the embedding kernel also writes its bf16 copy (unused here)
        self.e_stats = [buf(M) for _ in range(4)]
        self.pframe = buf(max(nfr, 1), H, dtype=BF16)
        self.qkv, self.ctx, self.y, self.a = buf(M, 3 * H), buf(M, H), buf(M, H), buf(M, H)
        self.hg = buf(M, I)
        self.t_g, self.t_ln = buf(M, H), buf(M, H)
        self.logits = buf(M, V)
        self.row_lse, self.ce_sum = buf(M), buf(4)
        self.heads_ws = torch.empty(capi.heads_workspace_bytes(B, H) // 4 + 16, device=device, dtype=F32)
        self.small_out = buf(8 + 7 * B)                   # see engine.Plan: one copy hands out all small results
        self.losses, self.logits_out = self.small_out[:8], self.small_out[8:8 + B]
        self.rel_out, self.align_out = self.small_out[8 + B:8 + 3 * B].view(B, 2), self.small_out[8 + 3 * B:].view(2 * B, 2)
        self.wT = [buf(self.Dv, H), buf(self.Da, H)]
        self._frame_sig = None
        self._build()

    _p = Plan._p
    _fn = Plan._fn
    refresh_frame_weights = Plan.refresh_frame_weights
    bind_inputs = Plan.bind_inputs
    convert_inputs = Plan.convert_inputs
    bind_tensors = Plan.bind_tensors
    set_loss_weights = Plan.set_loss_weights

    def set_seed(self, seed):         # dropout-free path
        pass
    run = staticmethod(Plan.run)

    def _linear(self, seq, X, W, bias, Y, M, N, K, act=capi.ACT_NONE):
        seq.append((self._fn("linear_f32"), capi.fill(capi.LinearF32Args(), X=X, W=W, bias=bias, Y=Y, ldx=K, ldw=K, ldy=N,
                                                      M=M, N=N, K=K, act=act)))

    def _ln(self, seq, y, res, name, out):
        c = self.cfg
        seq.append((self._fn("dropout_residual_ln_fwd"),
                    capi.fill(capi.DrlnFwdArgs(), y=y, res=res, gamma=self._p(name + ".weight"), beta=self._p(name + ".bias"),
                              out=None, out_f32=out, mean=None, rstd=None, M=self.M, H=self.H, eps=c.layer_norm_eps,
                              p_drop=0.0, seed=0, rng_stream=0, y_f32=1)))

    def _build(self):
        c, H, I, M, N, V = self.cfg, self.H, self.I, self.M, self.N, self.V
        st = self.store
        f = []
        self.pack_args = capi.fill(capi.PackArgs(), keybias=self.keybias, cu_seqlens=self.cu, label_count=self.label_count,
                                   kv_end=self.kv_end, B=self.B, T=self.T, L=[self.Lv, self.La], frame_dim=[self.Dv, self.Da])
        f.append((self._fn("pack_prepare"), self.pack_args))
        je = "bert.jointEmbeddings."
        m1, r1, m2, r2 = self.e_stats
        self.embed_args = capi.fill(
            capi.EmbedArgs(), frame_dim=[self.Dv, self.Da],
            word=self._p("bert.embeddings.word_embeddings.weight"), pos=self._p("bert.embeddings.position_embeddings.weight"),
            type=self._p("bert.embeddings.token_type_embeddings.weight"),
            ln1_g=self._p("bert.embeddings.LayerNorm.weight"), ln1_b=self._p("bert.embeddings.LayerNorm.bias"),
            ln2_g=self._p(je + "LayerNorm.weight"), ln2_b=self._p(je + "LayerNorm.bias"),
            wT=self.wT, wb=[self._p(je + "Wv.bias"), self._p(je + "Ws.bias")],
            eps1=c.layer_norm_eps, eps2=1e-5, p_drop1=0.0, p_drop2=0.0, seed=0,
            x0=self.x_bf16, x0_f32=self.x[0], mean1=m1, rstd1=r1, mean2=m2, rstd2=r2, pframe=self.pframe,
            B=self.B, T=self.T, L=[self.Lv, self.La], H=H, V=V, max_pos=c.max_position_embeddings, exact_frames=1)
        f.append((self._fn("embed_fwd"), self.embed_args))
        for l in range(N):
            xin, xout = self.x[l % 2], self.x[(l + 1) % 2]
            pre = f"bert.encoder.layer.{l}."
            wqkv = st.span(pre + "attention.self.query.weight", pre + "attention.self.value.weight").view(3 * H, H)
            bqkv = st.span(pre + "attention.self.query.bias", pre + "attention.self.value.bias")
            self._linear(f, xin, wqkv, bqkv, self.qkv, M, 3 * H, H)
            f.append((self._fn("attn_f32_fwd"),
                      capi.fill(capi.AttnF32Args(), qkv=self.qkv, ctx=self.ctx, keybias=self.keybias, cu_seqlens=self.cu, H=H,
                                nheads=self.nh, nseq=3 * self.B, max_seqlen=self.max_S, total_rows=M)))
            self._linear(f, self.ctx, self._p(pre + "attention.output.dense.weight"), self._p(pre + "attention.output.dense.bias"),
                         self.y, M, H, H)
            self._ln(f, self.y, xin, pre + "attention.output.LayerNorm", self.a)
            self._linear(f, self.a, self._p(pre + "intermediate.dense.weight"), self._p(pre + "intermediate.dense.bias"),
                         self.hg, M, I, H, act=capi.ACT_GELU)
            self._linear(f, self.hg, self._p(pre + "output.dense.weight"), self._p(pre + "output.dense.bias"), self.y, M, H, I)
            self._ln(f, self.y, self.a, pre + "output.LayerNorm", xout)
        self.seq_out = self.x[N % 2]
        tp = "cls.predictions.transform."
        self._linear(f, self.seq_out, self._p(tp + "dense.weight"), self._p(tp + "dense.bias"), self.t_g, M, H, H, act=capi.ACT_GELU)
        self._ln(f, self.t_g, None, tp + "LayerNorm", self.t_ln)
        self._linear(f, self.t_ln, self._p("bert.embeddings.word_embeddings.weight"), self._p("cls.predictions.bias"),
                     self.logits, M, V, H)
        self.ce_args = capi.fill(capi.CeArgs(), logits=self.logits, label_count=self.label_count, row_lse=self.row_lse,
                                 loss_sum=self.ce_sum, gscale=None, coef=self.alpha / 3.0, V=V, ldl=V, B=self.B, T=self.T,
                                 L=[self.Lv, self.La], dense=1, logits_f32=1)
        f.append((self._fn("ce_fwd"), self.ce_args))
        self.heads_args = capi.fill(
            capi.HeadsArgs(), seq_out=self.seq_out, cu_seqlens=self.cu, workspace=self.heads_ws,
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
            rel_out=self.rel_out, align_out=self.align_out, gscale=None, alpha=self.alpha, beta=self.beta,
            B=self.B, H=H, num_labels=self.num_labels, seq_out_f32=1)
        f.append((self._fn("heads_fwd"), self.heads_args))
        self.fwd = f
