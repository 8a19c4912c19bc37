"""Per-launch CUDA-event timing of one real training step (forward + backward plans), aggregated per kernel entry point
and, for the GEMM, per shape.  usage: python scripts/step_table.py [workload] [reps]"""
import collections, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from msa_b200 import capi, synth
from msa_b200.params import BertShape

wname = sys.argv[1] if len(sys.argv) > 1 else "mosi_aligned_b64"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
workload = synth.WORKLOADS[wname]
shape = BertShape(num_hidden_layers=12)
dev = torch.device("cuda", 0)
model = bench.build_model(shape, workload, dev)
model._ensure_store(dev)
batch = synth.tree_to(synth.make_workload_batch(workload, seed=1234), dev)
for _ in range(2):
    out, _ = model(**batch); out[0].backward()
plan = next(p for p in model._plans.values() if p.training)
names = {getattr(capi.lib(), n)._name if hasattr(getattr(capi.lib(), n), "_name") else n: n for n in capi.DECLARED_FUNCTIONS}
fn_name = {}
for n in capi.DECLARED_FUNCTIONS:
    try:
        fn_name[ctypes.addressof(getattr(capi.lib(), n))] = n
    except Exception:
        pass
stream = torch.cuda.current_stream(); sp = ctypes.c_void_p(stream.cuda_stream)
agg = collections.OrderedDict()
for rep in range(reps):
    model(**batch)
    model._prepare_grads()
    evs = []
    for seq in (plan.fwd, plan.bwd):
        for fn, a in seq:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream); capi.check(fn(ctypes.byref(a), sp), "launch"); e.record(stream)
            key = fn.__name__ if hasattr(fn, "__name__") else str(fn)
            if "gemm" in key:
                epi = {0: "bf16", 1: "gelu", 2: "relu", 3: "f32", 4: "atomic", 5: "dgelu", 6: "gelu+grad", 7: "mul_aux", 8: "ce_stats"}[a.epilogue]
                key = f"gemm M={a.M} N={a.N} K={a.K} {'MN' if a.a_major else 'K'}{'MN' if a.b_major else 'K'} {epi} sk={a.split_k}"
                fl = 2.0 * a.M * a.N * a.K
            else:
                fl = 0.0
            evs.append((key, s, e, fl))
    torch.cuda.synchronize()
    if rep == 0:
        continue
    for key, s, e, fl in evs:
        d = agg.setdefault(key, [0, 0.0, 0.0])
        d[0] += 1; d[1] += s.elapsed_time(e); d[2] += fl
n = reps - 1
tot = sum(d[1] for d in agg.values()) / n
print(f"# {wname}: sum of per-launch event times = {tot:.3f} ms per step (launch gaps excluded)")
for key, (c, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tf = f"{fl / ms / 1e9:7.1f} TF/s" if fl else ""
    print(f"{ms / n:8.3f} ms {100 * ms / n / tot:5.1f}%  x{c // n:3d}  {1e3 * ms / c:8.1f} us  {tf}  {key}")
