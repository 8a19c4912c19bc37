"""Records the K-step sentiment-MAE comparison (tests/kstep.py) at full bert-base (12 layers) with the reference's dropout
rates and stepping rule:   python scripts/kstep_mae.py [--layers 12] [--K 8] [--seeds 3] > gpurun_out/r2_kstep_mae.json"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mmbert_oracle as O  # noqa: E402
from tests import kstep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=12)
ap.add_argument("--K", type=int, default=8)
ap.add_argument("--seeds", type=int, default=3)
ap.add_argument("--B", type=int, default=8)
ap.add_argument("--T", type=int, default=20)
ap.add_argument("--lr", type=float, default=5e-4)
ap.add_argument("--dataset", default="mosi")
ap.add_argument("--dropout", type=int, default=1, help="0: dropout off on both sides (deterministic comparison, one seed)")
a = ap.parse_args()
t0 = time.time()
if a.dropout:
    r = kstep.run(O.Cfg(num_hidden_layers=a.layers), a.dataset, K=a.K, B=a.B, T=a.T, L=a.T, seeds=a.seeds, lr=a.lr)
else:
    r = kstep.run(O.Cfg(num_hidden_layers=a.layers), a.dataset, K=a.K, B=a.B, T=a.T, L=a.T, seeds=1, lr=a.lr, p=(0.0, 0.0, 0.0))
r["seconds"] = time.time() - t0
print(json.dumps(r))
