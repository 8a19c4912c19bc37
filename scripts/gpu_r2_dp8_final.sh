#!/bin/bash
# final code at 8 GPUs of one box (weak scaling, B = 64 per GPU, MOSEI shape): one short 1-GPU line, then 8 ranks
#   gpurun --gpus 8 --timeout 600 -- 'bash scripts/gpu_r2_dp8_final.sh r2f8'
tag=${1:-r2f8}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-torch-baseline --no-all-rows 2>/dev/null | tail -1 > gpurun_out/${tag}_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 10 --warmup 3 2>gpurun_out/${tag}_n8.err | tail -1 > gpurun_out/${tag}_n8.json
python - <<PY
import json
for n in (1, 8):
    try:
        d = json.load(open(f"gpurun_out/${tag}_n{n}.json"))
        print("n", d["n_gpus"], round(d["value"], 1), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), d.get("dp"), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(n, "failed:", e)
PY
tail -2 gpurun_out/${tag}_n8.err
