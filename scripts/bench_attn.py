"""Times the attention kernels alone (CUDA events) on a packed batch shaped like one MMBert layer.
usage: python scripts/bench_attn.py [mosei|mosi] [p_drop]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msa_b200 import capi

shape = sys.argv[1] if len(sys.argv) > 1 else "mosei"
p_drop = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
B, nh, H = 64, 12, 768
T, L = (50, 500) if shape == "mosei" else (50, 50)
g = torch.Generator().manual_seed(0)
lens, valid = [], []
for _ in range(B):
    lens.append(T); valid.append(int(torch.randint(10, T + 1, (1,), generator=g)))
for _ in range(2 * B):
    lens.append(T + L); valid.append(T + int(torch.randint(50, L + 1, (1,), generator=g)) if shape == "mosei" else T + int(torch.randint(10, L + 1, (1,), generator=g)))
cu = [0]
for n in lens:
    cu.append(cu[-1] + n)
rows = cu[-1]
dev = "cuda"
qkv = (torch.randn(rows, 3 * H, device=dev) * 0.5).to(torch.bfloat16)
keybias = torch.zeros(rows, device=dev)
for i, (n, v) in enumerate(zip(lens, valid)):
    keybias[cu[i] + v:cu[i] + n] = -10000.0
cu_t = torch.tensor(cu, device=dev, dtype=torch.int32)
kv_end = torch.tensor(valid, device=dev, dtype=torch.int32)
ctx = torch.empty(rows, H, device=dev, dtype=torch.bfloat16)
lse = torch.empty(nh, rows, device=dev)
dctx = torch.randn(rows, H, device=dev).to(torch.bfloat16)
dqkv = torch.empty(rows, 3 * H, device=dev, dtype=torch.bfloat16)
bwd_ws = capi.attn_bwd_workspace(rows, nh, dev)
flops_fwd = sum(4.0 * n * n * H for n in lens)           # dense S x S, as the reference computes
for flags in (0,):
    a = capi.attn_args(qkv, ctx, lse, keybias, cu_t, H, nh, max(lens), dctx=dctx, dqkv=dqkv, bwd_ws=bwd_ws, kv_end=kv_end,
                       p_drop=p_drop, seed=1, rng_stream=1, flags=flags)
    for name in ("attn_fwd", "attn_bwd"):
        for _ in range(3):
            capi.call(name, a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for _ in range(n):
            capi.call(name, a)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = flops_fwd * (1.0 if name == "attn_fwd" else 2.5)
        print(f"{shape} p={p_drop} {name}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s (dense-equivalent)")
