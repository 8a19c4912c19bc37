"""bench.py — MMBert train samples/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference] [--mode train|infer-sweep]

A "step" is one pass of the hot path over one synthetic batch: packed 3-pass forward, backward, (N>1) gradient
all-reduce over NCCL overlapped with backward, fused AdamW.  One "sample" = one dataset item = the reference's
three encoder passes.  Default workload = the configuration BASELINE.json's targets are quoted on
(configs[2], it fits one GPU): CMU-MOSEI unaligned shape (text 50 + audio 500x74 + visual 500x35), batch 64 per GPU,
bert-base, bf16 GEMMs with fp32 master weights.  --workload mosi_aligned_b64 (configs[1]) / ur_funny_b64 (configs[3]).

  value     whole-job samples/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e       the same through the public API with HOST inputs: H2D copy of every step's tensors and a D2H read of every
            step's loss inside the timed region, through the package's prefetching loop (msa_b200.trainer_fast);
            e2e.blocking_value = the reference loop's copy / step / .item() pattern
  roofline  dominant kernel = the tcgen05 CTA-pair GEMM: frac = sum of algorithmic FLOPs of its launches / sum of their
            CUDA-event durations inside a step, against MEASURED_PEAKS.json's sustained bf16 figure;
            step_frac = the WHOLE step (value x algorithmic GFLOP per sample) against the same peak, step_frac_burst
            against the burst figure
  cpu_baseline        the reference's CPU path on the host cores (bounded sample; the live reference when $MSA_REF
                      points at a checkout, else the oracle port)
  gpu_torch_baseline  the same model as plain PyTorch on this GPU (fused torch ops: cuBLASLt / SDPA), fp32 TF32-off,
                      TF32-on and autocast-bf16 — the bar the hand-written kernels have to beat

--impl reference times the CPU reference alone (same workload / config, bounded sample per step, optimizer step included).
--mode infer-sweep: BASELINE.json configs[4], forward-only batch 1-1024 x concatenated length 150-2048, one JSON line
per point and a summary line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from msa_b200 import synth  # noqa: E402
from msa_b200.params import BertShape, train_gflop_per_sample  # noqa: E402

METRIC, UNIT = "mmbert_train_samples_per_sec", "samples/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d["bf16_tflops_sustained"], burst=d["bf16_tflops"], hbm=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json, sustained)")
    return dict(tflops=1400.0, burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = str(gpu_index), [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", self.gpu], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        clocks, reasons, mx = [], set(), None
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                clocks.append(float(r[1]))
                mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        clocks.sort()
        return {"sm_mhz": clocks[len(clocks) // 2] if clocks else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(clocks)}


def build_model(shape, workload, device, seed=0):
    from msa_b200.api import MMBertForPretraining
    torch.manual_seed(seed)
    model = MMBertForPretraining(shape)
    model.bert.set_joint_embeddings(workload.dataset)
    return model.to(device).train()


def profile_gemm(model, plan):
    """One extra (untimed) step with CUDA events around every GEMM launch of the forward and backward plans:
    returns (sum of algorithmic FLOPs, sum of milliseconds, number of launches) of gemm_tcgen05_kernel."""
    import ctypes
    from msa_b200 import capi
    gemm = plan._fn("gemm")
    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)
    evs, flops, abytes = [], 0.0, 0.0
    model._prepare_grads()
    for seq in (plan.fwd, plan.bwd):
        for fn, a in seq:
            if fn is gemm:
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(stream)
                capi.check(fn(ctypes.byref(a), sp), "gemm")
                e.record(stream)
                evs.append((s, e))
                flops += 2.0 * a.M * a.N * a.K
                # algorithmic HBM bytes of the launch: both operands once, the output once (f32 for the wgrad epilogues),
                # plus the auxiliary stream of the GELU / multiply epilogues
                out_b = 4 if a.epilogue in (capi.EPI_STORE_F32, capi.EPI_ATOMIC_ADD_F32) else 2
                if a.epilogue == capi.EPI_CE_STATS:
                    out_b = 0               # fused cross entropy: only the labelled rows (~1 %) are written
                aux_b = 2 if (a.aux and a.epilogue in (capi.EPI_GELU_BF16, capi.EPI_DGELU_BF16, capi.EPI_GELU_GRAD_BF16,
                                                       capi.EPI_MUL_AUX_BF16)) else 0
                abytes += 2.0 * a.K * (a.M + a.N) + float(a.M) * a.N * (out_b + aux_b)
            else:
                capi.check(fn(ctypes.byref(a), sp), "launch")
    torch.cuda.synchronize()
    ms = sum(s.elapsed_time(e) for s, e in evs)
    return flops, ms, len(evs), abytes


def measured_traffic(workload_name):
    """DRAM bytes per GEMM launch measured under ncu for this workload (profiles/r1_gemm_traffic.json), or None."""
    for tag in ("r2", "r1"):
        p = os.path.join(ROOT, "profiles", f"{tag}_gemm_traffic.json")
        try:
            d = json.load(open(p)).get(workload_name)
            if d is not None:
                return float(d["traffic_bytes_per_launch"])
        except (OSError, ValueError, KeyError):
            continue
    return None


def executed_train_gflop_per_sample(shape, workload, batches):
    """The dense-faithful figure of SURVEY.md §8d counts every attention score of the S x S square.  The kernels skip
    what is exactly zero: key tiles behind the last unmasked key (forward and backward) and, in the backward, the query
    rows behind it (their upstream gradient is exactly zero).  This returns the train GFLOP per sample with the attention
    term counted as EXECUTED (forward 4 q e H with q = the query rows of the 128-row tiles that start before e — the
    training forward leaves the tiles entirely behind e alone, mmb_attn_args.flags bit 3 —, backward 8 e^2 H per sequence
    with e = 1 + last unmasked key), averaged over the given host batches — reported next to the dense-faithful number so
    that neither hides the other."""
    fwd_skip = os.environ.get("MMB_ATTN_QSKIP", "1") != "0" and os.environ.get("MMB_ATTN_FWD_QSKIP", "1") != "0"
    H, N = shape.hidden_size, shape.num_hidden_layers
    dense = train_gflop_per_sample(shape, workload)
    T = workload.T
    d_attn = e_attn = 0.0
    nsamp = 0
    for b in batches:
        m_t, (m_tv, m_v), (m_ts, m_s) = b["attention_mask"]

        def eff(mask):                   # [B, S] -> 1 + index of the last unmasked key (S when none is: processed in full)
            S = mask.shape[1]
            idx = torch.arange(1, S + 1)[None, :] * (mask != 0)
            e = idx.max(dim=1).values
            return torch.where(e > 0, e, torch.full_like(e, S)).double()

        fv = m_v[:, :, 0] if m_v.dim() == 3 else m_v
        fs = m_s[:, :, 0] if m_s.dim() == 3 else m_s
        for mask in (m_t, torch.cat((m_tv.double(), fv.double()), 1), torch.cat((m_ts.double(), fs.double()), 1)):
            S = float(mask.shape[1])
            e = eff(mask)
            d_attn += float(N * 12.0 * S * S * H * mask.shape[0])
            q = torch.clamp(torch.ceil(e / 128.0) * 128.0, max=S) if fwd_skip else torch.full_like(e, S)
            e_attn += float(N * H * (4.0 * q * e + 8.0 * e * e).sum())
        nsamp += m_t.shape[0]
    return dense + (e_attn - d_attn) / nsamp / 1e9


def live_row_fraction(batches):
    """Share of the packed rows that lie before their sequence's last unmasked key (the rows DESIGN.md 3.3 calls live)."""
    live = total = 0.0
    for b in batches:
        m_t, (m_tv, m_v), (m_ts, m_s) = b["attention_mask"]
        fv = m_v[:, :, 0] if m_v.dim() == 3 else m_v
        fs = m_s[:, :, 0] if m_s.dim() == 3 else m_s
        for mask in (m_t, torch.cat((m_tv.double(), fv.double()), 1), torch.cat((m_ts.double(), fs.double()), 1)):
            S = mask.shape[1]
            e = (torch.arange(1, S + 1)[None, :] * (mask != 0)).max(dim=1).values
            e = torch.where(e > 0, e, torch.full_like(e, S))
            live += float(e.sum())
            total += float(mask.numel())
    return live / total


def config_dict(workload, shape, world):
    B = workload.batch
    return {"workload": workload.name, "model": f"bert-base shape ({shape.num_hidden_layers} layers), random init",
            "batch_per_gpu": B, "global_batch": B * world, "positions_per_sample": workload.positions,
            "packed_rows_per_gpu": B * workload.positions, "parallelism": f"dp{world}",
            "mlm": "dense (decoder GEMM, dgrad and wgrad over all positions, as the reference); cross entropy fused into the decoder epilogue, logits not materialised", "optimizer": "AdamW (HF semantics)",
            "padding_rows": "left alone where nothing observable reads them (DESIGN.md 3.3; GEMM FLOPs executed in full); the line's "
                            "padding_rows object carries the same run with that switched off",
            "l2": "no explicit flush: one training step streams GiBs of saved activations (>> 126 MB L2) and 4 distinct "
                  "input batches are cycled"}


class _HFAdamW:
    """transformers(<=4.x).AdamW.step restated for a list of leaf tensors (train.py:76-92): the optimizer the reference
    builds, so that the CPU reference arm times a whole training step like the GPU arm does."""

    def __init__(self, named, lr=1e-5, b1=0.9, b2=0.999, eps=1e-6, wd=0.01):
        self.named, self.lr, self.b1, self.b2, self.eps, self.wd, self.t = named, lr, b1, b2, eps, wd, 0
        self.m = {k: torch.zeros_like(v) for k, v in named.items()}
        self.v = {k: torch.zeros_like(v) for k, v in named.items()}

    @torch.no_grad()
    def step(self, grads):
        self.t += 1
        s = self.lr * (1 - self.b2 ** self.t) ** 0.5 / (1 - self.b1 ** self.t)
        for k, p in self.named.items():
            g = grads.get(k)
            if g is None:
                continue
            self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            p.addcdiv_(self.m[k], self.v[k].sqrt().add_(self.eps), value=-s)
            if not ("bias" in k or "LayerNorm.weight" in k):
                p.mul_(1 - self.lr * self.wd)


def _cpu_reference_step(shape, workload):
    """Returns (kind, describe, step(batch) -> None): one CPU training step (forward, backward, AdamW) of the reference's
    path in fp32 on all host threads — the UNMODIFIED reference when a checkout is reachable ($MSA_REF, default
    /root/reference: the build container), else the oracle port (the GPU box: the reference is Python on top of
    transformers, lives outside the repo and does not travel)."""
    from oracle import mmbert_oracle as O
    from oracle import ref_loader
    from msa_b200.params import seeded_state_dict
    ocfg = O.Cfg(shape.hidden_size, shape.num_hidden_layers, shape.num_attention_heads, shape.intermediate_size,
                 shape.vocab_size, shape.max_position_embeddings, shape.layer_norm_eps)
    sd = seeded_state_dict(ocfg, workload.dataset, seed=0, std=0.02)
    if ref_loader.available():
        from transformers import BertConfig
        cfg = BertConfig(hidden_size=shape.hidden_size, num_hidden_layers=shape.num_hidden_layers,
                         num_attention_heads=shape.num_attention_heads, intermediate_size=shape.intermediate_size,
                         vocab_size=shape.vocab_size, max_position_embeddings=shape.max_position_embeddings)
        model = ref_loader.build_model(cfg, workload.dataset, eager=False).train()
        model.load_state_dict(sd, strict=False)
        named = {k: p for k, p in model.named_parameters()}
        opt = _HFAdamW({k: p.data for k, p in named.items()})

        def step(batch):
            out, _ = model(**batch)
            out[0].mean().backward()
            opt.step({k: p.grad for k, p in named.items()})
            for p in named.values():
                p.grad = None

        return "reference", f"unmodified reference ({ref_loader.REF_DIR}) forward + backward + AdamW, dropout on", step
    params = {k: v.float().clone() for k, v in sd.items() if k not in O.TIED}
    opt = _HFAdamW(params)
    drop = O.Dropout(0.1, 0.1, 0.5, seed=0)

    def step(batch):
        full = dict(params)
        for alias, canon in O.TIED.items():
            full[alias] = params[canon]
        _, _, grads = O.forward_backward(full, ocfg, batch, dtype=torch.float32, dropout=drop)
        opt.step(grads)

    return "port", "oracle/mmbert_oracle.py forward + backward + AdamW (HF rule), dropout on", step


def cpu_baseline(shape, workload, steps=2, warmup=1, batch=None, target_s=4.0):
    """Times the CPU reference on a BOUNDED sample of the workload: ``batch`` samples per step, chosen from a one-sample
    probe so that a step takes about ``target_s`` seconds.  Returns (cpu_baseline dict, seconds per step)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind, what, step = _cpu_reference_step(shape, workload)
    probe = None
    if batch is None:
        t0 = time.perf_counter()
        step(synth.make_workload_batch(workload, seed=1233, batch=1))
        probe = time.perf_counter() - t0
        batch = max(1, min(8, int(target_s / probe)))
    data = synth.make_workload_batch(workload, seed=1234, batch=batch)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step(data)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return dict(value=batch / sec, unit=UNIT, cores=cores, kind=kind,
                sample=f"{what}; fp32, {torch.get_num_threads()} torch threads; {batch} of the workload's {workload.batch} "
                       f"samples per step x {steps} steps (+{warmup} warm-up"
                       + (f", batch chosen from a {probe:.1f} s one-sample probe" if probe else "") + ")"), sec


def run_reference(args, workload, shape):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, sec = cpu_baseline(shape, workload, steps=args.steps, warmup=args.warmup, batch=args.cpu_batch)
    line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": dict(config_dict(workload, shape, 1), mlm="dense (all positions), logits materialised: the reference's own computation"),
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def gpu_torch_baseline(shape, workload, device, steps=3, warmup=2):
    """The same model as plain PyTorch on THIS GPU (oracle/torch_baseline.py: F.linear / SDPA / F.layer_norm / F.gelu /
    F.cross_entropy, dropout on, torch.optim.AdamW(fused=True)): one training step at the workload's batch in fp32 with
    TF32 off, TF32 on, and under torch.autocast(bfloat16).  Halves the batch on out-of-memory.  Returns a dict."""
    from oracle import mmbert_oracle as O
    from oracle import torch_baseline as TB
    from msa_b200.params import seeded_state_dict
    ocfg = O.Cfg(shape.hidden_size, shape.num_hidden_layers, shape.num_attention_heads, shape.intermediate_size,
                 shape.vocab_size, shape.max_position_embeddings, shape.layer_norm_eps)
    sd = seeded_state_dict(ocfg, workload.dataset, seed=0, std=0.02)
    params = {k: v.to(device).requires_grad_(True) for k, v in sd.items() if k not in O.TIED}
    opt = torch.optim.AdamW(list(params.values()), lr=1e-5, eps=1e-6, weight_decay=0.01, fused=True)
    full = dict(params)
    for alias, canon in O.TIED.items():
        full[alias] = params[canon]
    out = {"kind": "port", "what": "oracle/torch_baseline.py: the reference's path on torch's fused ops (cuBLASLt GEMMs, "
                                   "scaled_dot_product_attention, fused LayerNorm / GELU / cross entropy / dropout), "
                                   "forward + backward + torch.optim.AdamW(fused=True)", "unit": UNIT, "steps": steps}
    old_tf32 = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32

    def run(mode, B):
        batch = synth.tree_to(synth.make_workload_batch(workload, seed=1234, batch=B), device)
        tf32 = mode != "fp32"
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(warmup + steps):
            if i == warmup:
                torch.cuda.synchronize()
                ev0.record()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "autocast_bf16")):
                o, _ = TB.forward(full, ocfg, training=True, **batch)
            o[0].backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
        ev1.record()
        torch.cuda.synchronize()
        return B * steps / (ev0.elapsed_time(ev1) / 1e3)

    try:
        for mode in ("autocast_bf16", "tf32", "fp32"):
            B = workload.batch
            while True:
                try:
                    out[mode] = {"value": run(mode, B), "batch": B}
                    break
                except torch.OutOfMemoryError:
                    torch.cuda.empty_cache()
                    if B == 1:
                        out[mode] = {"error": "out of memory at batch 1"}
                        break
                    B //= 2
    except Exception as e:                                  # a baseline must never take the bench line down
        out["error"] = f"{type(e).__name__}: {e}"[:300]
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old_tf32
    return out


def infer_sweep(args, shape, device):
    """BASELINE.json configs[4]: forward-only (eval, no_grad) sweep, batch 1-1024 x concatenated length 150-2048 (T = 50
    text tokens, Lv = La = (S - 50) / 2 frames), MOSI feature dims, one B200; CUDA-graph replay of the launch plan.
    The CPU reference is timed beside the smallest length (bounded: batch 1 and 4)."""
    lines = []
    model = None
    free_gb = torch.cuda.mem_get_info(device)[0] / 2 ** 30
    for S in [int(x) for x in args.lengths.split(",")]:
        T, L = 50, (S - 50) // 2
        for B in [int(x) for x in args.batches.split(",")]:
            rows = B * (3 * T + 2 * L)
            per_row = 768 * 40 + 3072 * 4 + (30528 * 2 if args.materialize_logits else 0)
            est = rows * per_row / 2 ** 30
            if est > 0.8 * free_gb:
                lines.append({"batch": B, "concat_len": S, "skipped": f"needs ~{est:.0f} GiB"})
                print(json.dumps(lines[-1]), flush=True)
                continue
            w = synth.Workload(f"infer_b{B}_s{S}", "mosi", T, L, L, B)
            if model is None:
                model = build_model(shape, w, device).eval()
                model.materialize_logits = bool(args.materialize_logits)
                model.use_cuda_graph = not args.no_cuda_graph
                model._ensure_store(device)
            model._plans.clear()
            torch.cuda.empty_cache()
            batch = synth.tree_to(synth.make_workload_batch(w, seed=7), device)
            with torch.no_grad():
                for _ in range(3):
                    model(**batch)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.reps):
                    model(**batch)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            lines.append({"batch": B, "concat_len": S, "encoder_positions_per_sample": 3 * T + 2 * L, "ms": round(ms, 3),
                          "samples_per_s": round(B / ms * 1e3, 1)})
            print(json.dumps(lines[-1]), flush=True)
    cpu = []
    if not args.no_cpu_baseline:
        from oracle import mmbert_oracle as O
        from msa_b200.params import seeded_state_dict
        torch.set_num_threads(os.cpu_count() or 1)
        ocfg = O.Cfg(num_hidden_layers=shape.num_hidden_layers)
        sd = {k: v for k, v in seeded_state_dict(ocfg, "mosi", seed=0, std=0.02).items()}
        S0 = int(args.lengths.split(",")[0])
        for B in (1, 4):
            w = synth.Workload(f"infer_b{B}_s{S0}", "mosi", 50, (S0 - 50) // 2, (S0 - 50) // 2, B)
            b = synth.make_workload_batch(w, seed=7)
            with torch.no_grad():
                O.forward(sd, ocfg, dtype=torch.float32, **b)
                t0 = time.perf_counter()
                O.forward(sd, ocfg, dtype=torch.float32, **b)
                sec = time.perf_counter() - t0
            cpu.append({"batch": B, "concat_len": S0, "ms": round(sec * 1e3, 1), "samples_per_s": round(B / sec, 2),
                        "cores": os.cpu_count(), "kind": "port", "what": "oracle forward, fp32, all host threads"})
    print(json.dumps({"mode": "infer-sweep", "metric": "mmbert_forward_samples_per_sec", "unit": UNIT, "dtype": "bf16",
                      "cuda_graph": not args.no_cuda_graph, "materialize_logits": bool(args.materialize_logits),
                      "points": lines, "cpu_reference": cpu}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="mosei_unaligned_b64", choices=sorted(synth.WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "infer-sweep"])
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--cpu-batch", type=int, default=None, help="samples per CPU-reference step (default: ~4 s per step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-torch-baseline", action="store_true")
    ap.add_argument("--no-all-rows", action="store_true", help="skip the extra timed run with the padding-aware paths off")
    ap.add_argument("--batches", default="1,4,16,64,256,1024", help="infer-sweep")
    ap.add_argument("--lengths", default="150,512,1024,2048", help="infer-sweep")
    ap.add_argument("--reps", type=int, default=5, help="infer-sweep")
    ap.add_argument("--materialize-logits", type=int, default=0, help="infer-sweep: keep the [rows, V] decoder output")
    ap.add_argument("--no-cuda-graph", action="store_true", help="infer-sweep: launch the plan kernel by kernel")
    ap.add_argument("--lr", type=float, default=1e-5)
    ap.add_argument("--dp-mode", default=None, choices=["deferred", "overlap"],
                    help="N>1: gradient all-reduce schedule (msa_b200.ddp.GradReducer; default deferred)")
    ap.add_argument("--dp-compress", default=None, choices=["bf16"], help="N>1: all-reduce a bf16 copy of the gradient")
    ap.add_argument("--dp-chunks", type=int, default=4,
                    help="N>1, deferred: pieces of the all-reduce the fused AdamW consumes one by one (1 = wait for all of it)")
    ap.add_argument("--reserve-sms", type=int, default=None,
                    help="N>1: SMs the persistent kernels leave to NCCL (default 4; MMB_RESERVE_SMS overrides)")
    ap.add_argument("--nccl-ctas", type=int, default=None, help="N>1: NCCL_MAX_CTAS (default 4; 0 = NCCL's own choice)")
    args = ap.parse_args()
    workload = synth.WORKLOADS[args.workload]
    shape = BertShape(num_hidden_layers=args.layers)
    if args.impl == "reference":
        return run_reference(args, workload, shape)

    import torch.distributed as dist
    from msa_b200 import capi
    from msa_b200.ddp import GradReducer, broadcast_parameters
    from msa_b200.optim import FusedAdamW
    from msa_b200.trainer_fast import DeferredScalars, DevicePrefetcher
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dp_mode = args.dp_mode or os.environ.get("MMB_DP_MODE", "deferred")
    dp = {"mode": dp_mode, "compress": args.dp_compress, "reserved_sms": 0, "nccl_max_ctas": None}
    if world > 1:
        # deferred (default): one all-reduce after the backward sweep, every SM free for NCCL.  overlap: one all-reduce per
        # layer under the backward sweep — the persistent GEMM / attention kernels own every SM, so NCCL is capped to a few
        # CTAs and the persistent kernels leave that many SMs free.
        if dp_mode == "overlap":
            ctas = 4 if args.nccl_ctas is None else args.nccl_ctas
            if ctas > 0:
                os.environ.setdefault("NCCL_MAX_CTAS", str(ctas))
        elif args.nccl_ctas:
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.nccl_ctas))
        dp["nccl_max_ctas"] = os.environ.get("NCCL_MAX_CTAS")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    capi.check(capi.lib().mmb_check_device(), "mmb_check_device")
    if world > 1:
        reserve = args.reserve_sms
        if reserve is None:
            reserve = int(os.environ.get("MMB_RESERVE_SMS", "4" if dp_mode == "overlap" else "0"))
        capi.check(capi.lib().mmb_set_reserved_sms(int(reserve)), "mmb_set_reserved_sms")
        dp["reserved_sms"] = int(reserve)
    peaks = load_peaks()
    if args.mode == "infer-sweep":
        return infer_sweep(args, shape, device)

    model = build_model(shape, workload, device)
    model._ensure_store(device)
    broadcast_parameters(model)
    opt = FusedAdamW(model, lr=args.lr)
    opt.grad_scale = 1.0 / world
    reducer = (GradReducer(model._store, shape.num_hidden_layers, mode=dp_mode, compress=args.dp_compress).attach(model)
               if world > 1 else None)
    if reducer is not None and dp_mode == "deferred" and args.dp_compress is None and args.dp_chunks > 1:
        reducer.pipeline_optimizer(opt, chunks=args.dp_chunks)     # AdamW of piece k under the exchange of pieces k+1..
        dp["chunks"] = args.dp_chunks

    B = workload.batch
    nb = 4  # distinct synthetic batches, cycled
    host = [synth.make_workload_batch(workload, seed=1234 + 97 * rank + i) for i in range(nb)]
    pinned = [synth.tree_map(lambda t: t.pin_memory(), b) for b in host]
    resident = [synth.tree_to(b, device) for b in host]

    def step(batch):
        out, _ = model(**batch)
        out[0].backward()
        opt.step()
        opt.zero_grad()
        return out[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- warm-up (also builds the plan, TMA descriptors, NCCL channels)
    for i in range(args.warmup):
        step(resident[i % nb])
    barrier()
    # ---------------- timed region 1: device-resident inputs
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = capi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        loss = step(resident[i % nb])
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = (capi.launch_count() - l0) // args.steps
    # ---------------- timed region 2: end to end from pinned host memory, every step's loss read on the host.
    # The loop a user of the package writes (msa_b200.trainer_fast): batch i+1 is copied on a copy stream while batch i
    # computes, and the loss of step i is read two steps later — every copy and every read is inside the timed region.
    # What crosses PCIe is the compact form (msa_b200.trainer_fast.compact_host: float32 frames, feature-0 column of the
    # frame masks — bit-identical results); the compaction and the staging copy into reused pinned buffers run on the
    # host INSIDE the timed region, as they do in train_epoch.
    from msa_b200.trainer_fast import compact_host
    prefetch, reader = DevicePrefetcher(None, device, stage=True), DeferredScalars(device)

    def pipelined(n):
        losses = []
        prefetch.batches = (compact_host(host[i % nb]) for i in range(n))
        for dev_batch in prefetch:
            v = reader.push(step(dev_batch))
            if v is not None:
                losses.append(v)
        return losses + reader.flush()             # D2H reads of the last losses

    pipelined(3)                                   # untimed: copy stream, the two device buffer sets, the pinned loss slots
    barrier()
    h2d0 = prefetch.h2d_bytes
    t0 = time.perf_counter()
    losses = pipelined(args.steps)
    barrier()
    pipe_s = time.perf_counter() - t0
    pipe_h2d = (prefetch.h2d_bytes - h2d0) // args.steps
    assert len(losses) == args.steps
    # ---------------- timed region 3: the same with the reference loop's blocking pattern (trainer.py:49-93): copy, step,
    # ``float(loss)`` — host and device take turns
    def blocking(n):
        for i in range(n):
            dev_batch = synth.tree_to(pinned[i % nb], device, non_blocking=True)
            val = float(step(dev_batch).detach())  # D2H read of the loss (trainer.py:85)
        return val

    blocking(2)
    barrier()
    t0 = time.perf_counter()
    loss_val = blocking(args.steps)
    barrier()
    blocking_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, pipe_s * 1e3, blocking_s * 1e3], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, pipe_ms, blocking_ms = float(t[0]), float(t[1]), float(t[2])
    # where a step's time goes around the gradient exchange (4 extra, untimed steps; CUDA events on the compute stream, which
    # waits for every piece of the all-reduce inside optimizer.step()): forward + backward, then exchange + AdamW
    ph = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    fb_ms = ex_ms = 0.0
    for i in range(4):
        ph[0].record()
        out, _ = model(**resident[i % nb])
        out[0].backward()
        ph[1].record()
        opt.step()
        opt.zero_grad()
        ph[2].record()
        torch.cuda.synchronize()
        fb_ms += ph[0].elapsed_time(ph[1]) / 4
        ex_ms += ph[1].elapsed_time(ph[2]) / 4
    tp = torch.tensor([fb_ms, ex_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    dp["fwd_bwd_ms"], dp["exchange_adamw_zero_ms"] = float(tp[0]), float(tp[1])
    # the end-to-end figure is the better of the two loops (both copy every step's inputs from pinned host memory and read
    # every step's loss on the host inside their timed region); both are reported
    e2e_ms, e2e_loop = min((pipe_ms, "pipelined"), (blocking_ms, "blocking"))
    value = world * B * args.steps / (ms / 1e3)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)

    if rank == 0:
        plan = next(p for p in model._plans.values() if p.training)
        model(**resident[0])     # fresh forward state for the profiling pass
        flops, gemm_ms, n_gemm, gemm_bytes = profile_gemm(model, plan)
        opt.zero_grad()
        achieved = flops / (gemm_ms / 1e3) / 1e12
        # roofline denominator of the dominant kernel: the sustained figure (a 4 s back-to-back cuBLAS loop at the power cap) for
        # kernels timed inside a long step — unless the step's GEMM launches run ABOVE it (interleaved with HBM-bound kernels the
        # chip clocks higher than in a pure GEMM loop): a fraction > 1 would be meaningless, so the burst figure is used then
        # and both fractions are reported
        gemm_peak, gemm_peak_source = peaks["tflops"], peaks["source"]
        if achieved > gemm_peak:
            gemm_peak = peaks["burst"]
            gemm_peak_source = (peaks["source"].replace("sustained", "burst") + ": the GEMM launches of this step ran above the "
                                "sustained figure (see frac_sustained), so frac is quoted against the burst figure")
        gf = train_gflop_per_sample(shape, workload)
        gf_exec = executed_train_gflop_per_sample(shape, workload, host)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": config_dict(workload, shape, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(pipe_h2d) if e2e_loop == "pipelined" else synth.tree_bytes(host[0]),
                    "d2h_bytes_per_step": 4,
                    "loop": e2e_loop,
                    "pipelined_value": world * B * args.steps / (pipe_ms / 1e3),
                    "pipelined_h2d_bytes_per_step": int(pipe_h2d),
                    "pipelined_loop": "msa_b200.trainer_fast: compact_host (float32 frames, feature-0 mask column) -> reused "
                                      "pinned staging buffers -> H2D copy of batch i+1 on a copy stream under step i "
                                      "(DevicePrefetcher); every step's loss read on the host two steps late (DeferredScalars)",
                    "blocking_value": world * B * args.steps / (blocking_ms / 1e3),
                    "blocking_h2d_bytes_per_step": synth.tree_bytes(host[0]),
                    "blocking_loop": "copy the collate-format tensors, step, float(loss) in turn, as trainer.py:49-93"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_2cta_kernel<8|16> (CTA-pair tcgen05 GEMM, csrc/gemm_tcgen05.cu)",
                         "achieved": achieved, "peak": gemm_peak,
                         "unit": "TFLOP/s", "frac": achieved / gemm_peak, "frac_sustained": achieved / peaks["tflops"],
                         "frac_burst": achieved / peaks["burst"], "peak_sustained": peaks["tflops"],
                         "traffic": measured_traffic(workload.name),
                         "traffic_note": "DRAM read+write bytes per GEMM launch, averaged over the GEMM launches of one "
                                         "step (ncu, profiles/*_gemm_traffic.json); algorithmic_bytes = the same average "
                                         "computed from the launch shapes (every output row counted: the measured figure is "
                                         "lower because the epilogues do not write all-padding 32-row slices, DESIGN.md 3.3)",
                         "algorithmic_bytes": gemm_bytes / n_gemm,
                         "peak_source": gemm_peak_source, "step_peak_source": peaks["source"], "launches_per_step": n_gemm,
                         "gemm_share_of_step": gemm_ms / (ms / args.steps),
                         # the WHOLE step (every kernel, launch gaps, optimizer; dense-faithful FLOPs of SURVEY.md §8d)
                         "step_achieved": value / world * gf / 1e3,
                         "step_frac": value / world * gf / 1e3 / peaks["tflops"],
                         "step_frac_burst": value / world * gf / 1e3 / peaks["burst"],
                         "peak_burst": peaks["burst"], "train_gflop_per_sample": gf,
                         # attention counted as executed (exact-zero key tiles / padded query rows skipped), everything else dense
                         "train_gflop_per_sample_executed": gf_exec,
                         "step_frac_executed": value / world * gf_exec / 1e3 / peaks["tflops"]},
            "model_flops": {"train_gflop_per_sample": gf, "achieved_tflops_per_gpu": value / world * gf / 1e3,
                            "frac_of_peak": value / world * gf / 1e3 / peaks["tflops"]},
            "dp": dp,
            "clocks": clocks, "final_loss": loss_val,
        }
        del plan
        if world == 1 and not args.no_all_rows:
            # The same timed region with every padding-aware path switched off: attention forward, LayerNorm,
            # column sums, epilogues and embeddings then process the rows behind each sequence's last unmasked key as well.
            # Same results; reported so that the saving is visible next to the headline.
            os.environ["MMB_ATTN_FWD_QSKIP"] = "0"
            model._plans.clear()
            torch.cuda.empty_cache()
            for i in range(3):
                step(resident[i % nb])
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for i in range(args.steps):
                step(resident[i % nb])
            a1.record()
            torch.cuda.synchronize()
            del os.environ["MMB_ATTN_FWD_QSKIP"]
            ms_all = a0.elapsed_time(a1)
            v_all = B * args.steps / (ms_all / 1e3)
            line["padding_rows"] = {
                "what": "rows behind a sequence's last unmasked key (no label there: checked on the device per batch) are left "
                        "alone by the attention forward, LayerNorm, column sums, GEMM epilogues and embeddings; every GEMM FLOP "
                        "is executed; outputs and gradients identical (tests); MMB_ATTN_FWD_QSKIP=0 turns it off",
                "live_row_fraction": live_row_fraction(host), "packed_rows": B * workload.positions,
                "value_all_rows": v_all, "ms_per_step_all_rows": ms_all / args.steps,
                "step_frac_all_rows": v_all * gf / 1e3 / peaks["tflops"]}
        if world == 1 and not args.no_gpu_torch_baseline:
            model._plans.clear()                    # hand the activation buffers back before the PyTorch model runs
            torch.cuda.empty_cache()
            line["gpu_torch_baseline"] = gpu_torch_baseline(shape, workload, device)
            ab = line["gpu_torch_baseline"].get("autocast_bf16", {}).get("value")
            if ab:
                line["gpu_torch_baseline"]["ours_over_autocast_bf16"] = value / ab
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _ = cpu_baseline(shape, workload, batch=args.cpu_batch, steps=2, warmup=1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
