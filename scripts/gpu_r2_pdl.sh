#!/bin/bash
# Same-box A/B of programmatic dependent launch (MMB_PDL) and of the forward attention's padded-tile skip
# (MMB_ATTN_FWD_QSKIP):  gpurun --timeout 1200 -- 'bash scripts/gpu_r2_pdl.sh r2q'
tag=${1:-r2q}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.txt 2>&1
tail -5 gpurun_out/${tag}_pytest.txt
run() {   # name, env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 300 python bench.py --no-cpu-baseline --no-gpu-torch-baseline "$@" 2>gpurun_out/${tag}_$name.err | tail -1 > gpurun_out/${tag}_$name.json
  python - "$name" gpurun_out/${tag}_$name.json <<PY
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(sys.argv[1], round(d["value"], 1), "samples/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 1),
          "gemm", round(d["roofline"]["achieved"], 1), "step_frac", round(d["roofline"]["step_frac"], 4),
          "exec", round(d["roofline"]["step_frac_executed"], 4), "loss", d.get("final_loss"), d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "failed:", e)
PY
}
S="--steps 12 --warmup 4"
run c3_base   MMB_PDL=0 MMB_ATTN_FWD_QSKIP=0 -- $S
run c3_pdl    MMB_PDL=1 MMB_ATTN_FWD_QSKIP=0 -- $S
run c3_skip   MMB_PDL=0 MMB_ATTN_FWD_QSKIP=1 -- $S
run c3_both   MMB_PDL=1 MMB_ATTN_FWD_QSKIP=1 -- $S
run c3_base2  MMB_PDL=0 MMB_ATTN_FWD_QSKIP=0 -- $S
run c3_both2  MMB_PDL=1 MMB_ATTN_FWD_QSKIP=1 -- $S
run c2_base   MMB_PDL=0 -- $S --workload mosi_aligned_b64
run c2_pdl    MMB_PDL=1 -- $S --workload mosi_aligned_b64
run c2_base2  MMB_PDL=0 -- $S --workload mosi_aligned_b64
run c2_pdl2   MMB_PDL=1 -- $S --workload mosi_aligned_b64
for m in 1 2; do
  MMB_PDL=$m timeout 200 python bench.py --mode infer-sweep --batches 1,4,16 --lengths 150,512 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/${tag}_infer_pdl$m.json
  python - $m gpurun_out/${tag}_infer_pdl$m.json <<PY
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print("infer MMB_PDL=" + sys.argv[1], [(p["batch"], p["concat_len"], p["ms"]) for p in d["points"]])
except Exception as e:
    print("infer failed:", e)
PY
done
