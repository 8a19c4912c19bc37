#!/bin/bash
# 2-GPU sanity of the final code: tests, 1-GPU and 2-GPU bench lines (deferred all-reduce), DP gradient = mean of per-rank gradients
#   gpurun --gpus 2 --timeout 1200 -- 'bash scripts/gpu_r2_dpcheck.sh r2v'
tag=${1:-r2v}
N=${2:-2}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.txt 2>&1
tail -3 gpurun_out/${tag}_pytest.txt
timeout 300 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-gpu-torch-baseline 2>/dev/null | tail -1 > gpurun_out/${tag}_n1.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 12 --warmup 4 2>gpurun_out/${tag}_n$N.err | tail -1 > gpurun_out/${tag}_n$N.json
python - <<PY
import json
for n in (1, $N):
    try:
        d = json.load(open(f"gpurun_out/${tag}_n{n}.json"))
        print("n", d["n_gpus"], round(d["value"], 1), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), d.get("dp"), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(n, "failed:", e)
PY
MMB_DP_MODE=deferred timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/ddp_check.py > gpurun_out/${tag}_ddp_check.txt 2>&1
grep -E "max rel|Error|error|ok|OK" gpurun_out/${tag}_ddp_check.txt | cut -c1-200 | tail -8
