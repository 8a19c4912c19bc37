"""Stand-alone timing of the HBM-bound row kernels at one workload's packed row count: achieved GB/s against the
algorithmic bytes of DESIGN.md §3 (drln_fwd 12 B/element, drln_bwd 18 B/element, colsum 2 B/element).
usage: python scripts/bench_rowops.py [rows] [H]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msa_b200 import capi

M = int(sys.argv[1]) if len(sys.argv) > 1 else 73600
H = int(sys.argv[2]) if len(sys.argv) > 2 else 768
dev = "cuda"
torch.manual_seed(0)
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
f32 = lambda *s: torch.randn(*s, device=dev)
y, g1, out, d_y = bf(M, H), bf(M, H), bf(M, H), bf(M, H)
res, g2, out32, d_res = f32(M, H), f32(M, H), f32(M, H), f32(M, H)
gamma, beta = f32(H), f32(H)
mean, rstd = f32(M), f32(M).abs() + 0.5
dgamma, dbeta, dbias = torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev)
big = bf(M, 4 * H)
cs = torch.zeros(4 * H, device=dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


cases = [
    ("drln_fwd p=0.1", 12, lambda: capi.drln_fwd(y, res, gamma, beta, out, mean, rstd, 1e-12, p_drop=0.1, seed=5, rng_stream=2, out_f32=out32)),
    ("drln_bwd p=0.1", 18, lambda: capi.drln_bwd(g1, g2, y, res, mean, rstd, gamma, d_y, d_res, dgamma, dbeta, dbias, p_drop=0.1, seed=5, rng_stream=2)),
    ("drln_bwd p=0", 18, lambda: capi.drln_bwd(g1, g2, y, res, mean, rstd, gamma, d_y, d_res, dgamma, dbeta, dbias)),
    ("drln_bwd no g2 (last layer)", 14, lambda: capi.drln_bwd(g1, None, y, res, mean, rstd, gamma, d_y, d_res, dgamma, dbeta, dbias, p_drop=0.1, seed=5, rng_stream=2)),
]
for name, bpe, fn in cases:
    us = timeit(fn)
    print(f"{name:32s} {us:8.1f} us  {M * H * bpe / us / 1e3:8.1f} GB/s")
us = timeit(lambda: capi.colsum(big, cs))
print(f"{'colsum [M, 4H]':32s} {us:8.1f} us  {M * 4 * H * 2 / us / 1e3:8.1f} GB/s")
us = timeit(lambda: capi.colsum(big[:, :3 * H], cs[:3 * H]))
print(f"{'colsum [M, 3H] (ld 4H)':32s} {us:8.1f} us  {M * 3 * H * 2 / us / 1e3:8.1f} GB/s")
