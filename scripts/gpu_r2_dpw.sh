#!/bin/bash
# round 2: data-parallel runs of the other two training shapes (BASELINE.json configs[1], [3]) at N GPUs of one box
tag=${1:-r2dpw}
N=${2:-2}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for w in ur_funny_b64 mosi_aligned_b64; do
  timeout 200 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-torch-baseline 2>/dev/null | tail -1 > gpurun_out/${tag}_${w}_n1.json
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --workload $w --steps 20 --warmup 5 2>gpurun_out/${tag}_${w}_n$N.err | tail -1 > gpurun_out/${tag}_${w}_n$N.json
  python - <<PY
import json
a = json.load(open("gpurun_out/${tag}_${w}_n1.json")); b = json.load(open("gpurun_out/${tag}_${w}_n$N.json"))
print("$w", "n1", round(a["value"], 1), round(a["ms_per_step"], 2), "| n$N", round(b["value"], 1), round(b["ms_per_step"], 2), "e2e", round(b["e2e"]["value"], 1),
      "efficiency", round(b["value"] / ($N * a["value"]), 3), b["dp"])
PY
done
