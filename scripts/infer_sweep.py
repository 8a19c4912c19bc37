"""BASELINE.json configs[4]: forward-only (eval) sweep over batch 1-1024 and concatenated length 150-2048
(T = 50 text tokens, Lv = La = (S_concat - 50) / 2 frames), MOSI feature dims, bert-base, one B200.
Prints one JSON line per (batch, S_concat): samples/s and latency; combinations whose activation + logits buffers would
not fit in HBM are skipped (and listed).  The reference on the host CPU is timed by bench.py --impl reference.

usage: python scripts/infer_sweep.py [--batches 1,4,16,64,256,1024] [--lengths 150,512,1024,2048] [--reps 5]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from msa_b200 import synth
from msa_b200.params import BertShape

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="1,4,16,64,256,1024")
ap.add_argument("--lengths", default="150,512,1024,2048")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--mem-gb", type=float, default=150.0)
args = ap.parse_args()
dev = torch.device("cuda", 0)
shape = BertShape(num_hidden_layers=12)
model = None
for S in [int(x) for x in args.lengths.split(",")]:
    T, L = 50, (S - 50) // 2
    for B in [int(x) for x in args.batches.split(",")]:
        rows = B * (3 * T + 2 * L)
        # eval plan: logits bf16 [rows, 30528] + ~40 bytes/element of layer scratch at H = 768, I = 3072
        est = rows * (30528 * 2 + 768 * 40 + 3072 * 4) / 2 ** 30
        if est > args.mem_gb:
            print(json.dumps({"batch": B, "concat_len": S, "skipped": f"needs ~{est:.0f} GiB"}), flush=True)
            continue
        w = synth.Workload(f"infer_b{B}_s{S}", "mosi", T, L, L, B)
        if model is None:
            model = bench.build_model(shape, w, dev).eval()
            model._ensure_store(dev)
        model._plans.clear()
        torch.cuda.empty_cache()
        batch = synth.tree_to(synth.make_workload_batch(w, seed=7), dev)
        with torch.no_grad():
            for _ in range(2):
                model(**batch)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                model(**batch)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        print(json.dumps({"batch": B, "concat_len": S, "encoder_positions_per_sample": 3 * T + 2 * L, "ms": round(ms, 3),
                          "samples_per_s": round(B / ms * 1e3, 1)}), flush=True)
