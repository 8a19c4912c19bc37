#!/bin/bash
# Same-box A/B of two staging tiles per epilogue warp (MMB_GEMM_STAGE2) + per-launch tables for the forward tile skip:
#   gpurun --timeout 1200 -- 'bash scripts/gpu_r2_stage2.sh r2s'
tag=${1:-r2s}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
MMB_GEMM_STAGE2=1 timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x -k "gemm or tail_skip or golden" > gpurun_out/${tag}_pytest_gemm.txt 2>&1
tail -2 gpurun_out/${tag}_pytest_gemm.txt
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
run c3_s0  MMB_GEMM_STAGE2=0 -- $S
run c3_s1  MMB_GEMM_STAGE2=1 -- $S
run c3_s2  MMB_GEMM_STAGE2=2 -- $S
run c3_s0b MMB_GEMM_STAGE2=0 -- $S
run c3_s2b MMB_GEMM_STAGE2=2 -- $S
run c2_s0  MMB_GEMM_STAGE2=0 -- $S --workload mosi_aligned_b64
run c2_s2  MMB_GEMM_STAGE2=2 -- $S --workload mosi_aligned_b64
MMB_ATTN_FWD_QSKIP=0 timeout 150 python scripts/step_table.py mosei_unaligned_b64 > gpurun_out/${tag}_step_table_noskip.txt 2>&1
MMB_ATTN_FWD_QSKIP=1 timeout 150 python scripts/step_table.py mosei_unaligned_b64 > gpurun_out/${tag}_step_table_skip.txt 2>&1
grep -h "attn_fwd\|sum of" gpurun_out/${tag}_step_table_noskip.txt gpurun_out/${tag}_step_table_skip.txt
MMB_ATTN_FWD_QSKIP=1 MMB_GEMM_STAGE2=2 timeout 150 python scripts/step_table.py mosei_unaligned_b64 > gpurun_out/${tag}_step_table_stage2.txt 2>&1
head -24 gpurun_out/${tag}_step_table_stage2.txt
