#!/bin/bash
# round 2: data-parallel runs (N GPUs of one box): bench at the MOSEI shape, reserved-SM A/B, DDP parity check
tag=${1:-r2dp}
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
run c3_default
run c3_reserve0 --reserve-sms 0 --nccl-ctas 0
run c3_reserve8 --reserve-sms 8 --nccl-ctas 8
run c3_reserve2 --reserve-sms 2 --nccl-ctas 2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/ddp_check.py > gpurun_out/${tag}_ddp_check.txt 2>&1
tail -5 gpurun_out/${tag}_ddp_check.txt
tail -3 gpurun_out/${tag}_c3_default.err
