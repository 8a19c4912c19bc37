"""ONE training step (forward, backward, fused AdamW, gradient zeroing) between cudaProfilerStart / Stop, after warm-up
steps — the unit every ncu capture of profiles/ is taken on:
    ncu --profile-from-start off ... python scripts/one_step.py [workload] [warmup]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from msa_b200 import synth
from msa_b200.optim import FusedAdamW
from msa_b200.params import BertShape

wname = sys.argv[1] if len(sys.argv) > 1 else "mosei_unaligned_b64"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 3
workload = synth.WORKLOADS[wname]
dev = torch.device("cuda", 0)
model = bench.build_model(BertShape(num_hidden_layers=12), workload, dev)
model._ensure_store(dev)
opt = FusedAdamW(model, lr=1e-5)
batches = [synth.tree_to(synth.make_workload_batch(workload, seed=1234 + i), dev) for i in range(2)]


def step(b):
    out, _ = model(**b)
    out[0].backward()
    opt.step()
    opt.zero_grad()


for i in range(warm):
    step(batches[i % 2])
torch.cuda.synchronize()
torch.cuda.profiler.start()
step(batches[warm % 2])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("one step of", wname, "done")
