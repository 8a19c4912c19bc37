"""torchrun --nproc-per-node N scripts/ddp_check.py : data-parallel parity on N GPUs.
Every rank runs the packed step on its own micro-batch with the bucketed, overlapped all-reduce attached; the
averaged gradient must equal the mean of the per-rank gradients computed WITHOUT communication (SURVEY.md §8e:
DDP semantics, per-rank CE means and per-rank CPC negatives)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from msa_b200 import synth
from msa_b200.api import MMBertForPretraining
from msa_b200.ddp import GradReducer, broadcast_parameters
from msa_b200.params import BertShape, seeded_state_dict

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shape = BertShape(num_hidden_layers=3, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
sd = seeded_state_dict(shape, "mosi", seed=4, std=0.02)


def build():
    m = MMBertForPretraining(shape)
    m.bert.set_joint_embeddings("mosi")
    m.bert.jointEmbeddings.dropout.p = 0.0
    m.load_state_dict(sd)
    return m.to(dev).train()


batches = [synth.tree_to(synth.make_batch(8, 24, 24, 24, 47, 74, seed=100 + r, min_len=6), dev) for r in range(world)]
m = build()
m._ensure_store(dev)
broadcast_parameters(m)
GradReducer(m._store, shape.num_hidden_layers).attach(m)
out, _ = m(**batches[rank])
out[0].backward()
torch.cuda.synchronize()
got = m._store.grad[:m._store.trainable_end].clone() / world

ref_m = build()
ref = None
for r in range(world):
    for p in ref_m.parameters():
        p.grad = None
    o, _ = ref_m(**batches[r])
    o[0].backward()
    g = ref_m._store.grad[:ref_m._store.trainable_end].clone()
    ref = g if ref is None else ref + g
ref /= world
torch.cuda.synchronize()
err = float((got - ref).abs().max() / ref.abs().max())
none_ok = all(p.grad is None for n, p in m.named_parameters() if "W_cv" in n or "W_cs" in n or "seq_relationship" in n)
print(f"rank {rank}: max rel diff DP vs mean-of-local grads = {err:.3e}; untouched params stay None: {none_ok}", flush=True)
# fp32 atomics make each local gradient run-to-run nondeterministic at the 1e-6 level
assert err < 1e-3 and none_ok
# pipelined optimizer: the all-reduce in 4 pieces, FusedAdamW waiting for one piece at a time — same parameters afterwards
from msa_b200.optim import FusedAdamW
after = []
for chunks in (1, 1, 4):
    mm = build()
    mm._ensure_store(dev)
    broadcast_parameters(mm)
    opt = FusedAdamW(mm, lr=1e-3)
    opt.grad_scale = 1.0 / world
    red = GradReducer(mm._store, shape.num_hidden_layers, mode="deferred").attach(mm)
    if chunks > 1:
        red.pipeline_optimizer(opt, chunks=chunks)
    for _ in range(2):
        o, _ = mm(**batches[rank])
        o[0].backward()
        opt.step()
        opt.zero_grad()
    torch.cuda.synchronize()
    after.append(mm._store.flat[:mm._store.trainable_end].clone())
# (Adam's first steps are sign-like: a gradient element within the 1e-6 run-to-run noise of the split-K atomics around zero
# moves its parameter by +-lr either way, so two PLAIN runs already differ by ~lr; the pipelined run must not differ more)
base = float((after[0] - after[1]).abs().max() / after[0].abs().max())
perr = float((after[0] - after[2]).abs().max() / after[0].abs().max())
nbad0 = int(((after[0] - after[1]).abs() > 1e-5).sum())
nbad = int(((after[0] - after[2]).abs() > 1e-5).sum())
print(f"rank {rank}: max rel parameter diff after 2 steps: plain vs plain {base:.3e} ({nbad0} elements > 1e-5), "
      f"plain vs pipelined {perr:.3e} ({nbad} elements > 1e-5)", flush=True)
assert perr < max(3 * base, 1e-5) and nbad < max(3 * nbad0, 100)
dist.barrier()
dist.destroy_process_group()
