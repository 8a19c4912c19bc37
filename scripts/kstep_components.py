"""Which dropout site drives the seed-to-seed spread of the K-step MAE comparison (tests/kstep.py)?  Small model, several
seeds, one site at a time:  python scripts/kstep_components.py [seeds]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import mmbert_oracle as O
from tests import kstep

seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
ocfg = O.Cfg(num_hidden_layers=2)        # bert-base width
for name, p in (("hidden 0.1", (0.1, 0.0, 0.0)), ("attention 0.1", (0.0, 0.1, 0.0)), ("joint 0.5", (0.0, 0.0, 0.5)),
                ("all", (0.1, 0.1, 0.5))):
    r = kstep.run(ocfg, "mosi", K=6, B=6, T=16, L=16, seeds=seeds, lr=5e-5, p=p, dtype=torch.float32)
    print(f"{name:14s} mae0 {r['mae_initial']:.6f} cuda {r['mean_cuda']:.6f} +- {r['std_cuda']:.6f}   oracle {r['mean_oracle']:.6f} +- {r['std_oracle']:.6f}   gap {r['gap']:.6f}", flush=True)
