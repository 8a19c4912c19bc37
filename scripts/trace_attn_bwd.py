"""Bring-up aid: per-role event timeline of CTA 0 of the attention backward kernel (library built with
MMB_NVCC_EXTRA=-DMMB_ATTN_TRACE).  Prints, per step, cycle offsets of: producer (wait-empty done, full arrived),
score warp (step-full seen, slot-empty seen, issued), acc warp (staged seen, issued), compute warp 0 (scores seen,
loaded, staged arrived)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from msa_b200 import capi
import scripts.bench_attn  # noqa: F401  (runs the benchmark once, leaving the last launch's trace)
buf = np.zeros((4, 1024, 4), dtype=np.uint64)
L = capi.lib()
L.mmb_debug_attn_trace.argtypes = [ctypes.c_void_p]
torch.cuda.synchronize()
assert L.mmb_debug_attn_trace(buf.ctypes.data) == 0
t0 = int(buf[buf > 0].min())
rel = lambda x: int(x) - t0 if x else -1
n = int(sys.argv[3]) if len(sys.argv) > 3 else 48
last = max(g for g in range(1024) if buf[3, g, 3] > 0)
print("steps traced:", last + 1, "total cycles:", int(buf[3, last, 3]) - t0, "avg per step:", (int(buf[3, last, 3]) - t0) / (last + 1))
print("items (compute warp 0): item | top  rowfullOK  recordsRead")
for i in range(12):
    I = buf[0, i]
    print(f"  item {i:3d} | {rel(I[0]):7d} {rel(I[1]):7d} {rel(I[3]):7d}")
print("step | prod: top emptyOK fullArr | score: top stepOK slotOK issued | acc: top stagedOK issued | comp: top scoresOK loaded stagedArr")
for g in range(n):
    P, S, A, C = buf[0, g], buf[1, g], buf[2, g], buf[3, g]
    print(f"{g:4d} | {rel(P[0]):7d} {rel(P[1]):7d} {rel(P[2]):7d} | {rel(S[0]):7d} {rel(S[1]):7d} {rel(S[2]):7d} {rel(S[3]):7d} | "
          f"{rel(A[0]):7d} {rel(A[1]):7d} {rel(A[2]):7d} | {rel(C[0]):7d} {rel(C[1]):7d} {rel(C[2]):7d} {rel(C[3]):7d}")
