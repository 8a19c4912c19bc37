#!/bin/bash
# Padding-aware row kernels (MMB_ROW_SKIP) and forward attention tile skip (MMB_ATTN_FWD_QSKIP): tests, same-box A/B,
# per-launch tables.   gpurun --timeout 1200 -- 'bash scripts/gpu_r2_rowskip.sh r2r'
tag=${1:-r2r}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.txt 2>&1
tail -4 gpurun_out/${tag}_pytest.txt
run() {
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
          "gemm", round(d["roofline"]["achieved"], 1), "step_frac", round(d["roofline"]["step_frac"], 4), "loss", d.get("final_loss"), d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "failed:", e)
PY
}
S="--steps 12 --warmup 4"
run c3_off   MMB_ATTN_FWD_QSKIP=0 -- $S
run c3_all   -- $S
run c3_off2  MMB_ATTN_FWD_QSKIP=0 -- $S
run c3_all2  -- $S
run c2_off   MMB_ATTN_FWD_QSKIP=0 -- $S --workload mosi_aligned_b64
run c2_all   -- $S --workload mosi_aligned_b64
run c4_off   MMB_ATTN_FWD_QSKIP=0 -- $S --workload ur_funny_b64
run c4_all   -- $S --workload ur_funny_b64
timeout 150 python scripts/step_table.py mosei_unaligned_b64 > gpurun_out/${tag}_step_table_mosei_unaligned_b64.txt 2>&1
head -32 gpurun_out/${tag}_step_table_mosei_unaligned_b64.txt
