#!/bin/bash
# round 2: data-parallel schedule A/B at N GPUs: deferred (fp32 / bf16) against overlapped buckets
tag=${1:-r2dp2}
N=${2:-2}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() {  # name, extra args
  name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 12 --warmup 4 "$@" 2>gpurun_out/${tag}_${name}.err | tail -1 > gpurun_out/${tag}_${name}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${name}.json"))
    print("${name}", "n", d["n_gpus"], round(d["value"], 1), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), d["dp"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("${name} failed:", e)
PY
}
timeout 300 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-gpu-torch-baseline 2>/dev/null | tail -1 > gpurun_out/${tag}_n1.json
python -c "
import json; d=json.load(open('gpurun_out/${tag}_n1.json')); print('n1', round(d['value'],1), round(d['ms_per_step'],2))"
run deferred
run deferred_nopipe --dp-chunks 1
run deferred_bf16 --dp-compress bf16
run overlap --dp-mode overlap
run deferred_again
MMB_DP_MODE=deferred timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/ddp_check.py > gpurun_out/${tag}_ddp_check.txt 2>&1
grep -E "max rel|Error|error" gpurun_out/${tag}_ddp_check.txt | cut -c1-200
tail -3 gpurun_out/${tag}_deferred.err
