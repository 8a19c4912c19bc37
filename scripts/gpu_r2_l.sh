#!/bin/bash
tag=${1:-r2l}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm" 2>&1 | tail -3
echo "== wide boxes, 3 stages"; timeout 300 python scripts/bringup_gemm.py z_ffn1_gg z_c2_ffn1_gg y_ffn1_gelu_1stream z_ragged_gg 2>&1 | cut -c1-230
echo "== wide boxes, 2 stages"; MMB_GEMM_STAGES=2 timeout 300 python scripts/bringup_gemm.py z_ffn1_gg 2>&1 | cut -c1-230
echo "== 32-col boxes, 5 stages"; MMB_GEMM_WIDEBOX=0 timeout 300 python scripts/bringup_gemm.py z_ffn1_gg z_c2_ffn1_gg y_ffn1_gelu_1stream 2>&1 | cut -c1-230
echo "== 32-col boxes, 3 stages"; MMB_GEMM_WIDEBOX=0 MMB_GEMM_STAGES=3 timeout 300 python scripts/bringup_gemm.py z_ffn1_gg 2>&1 | cut -c1-230
b() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-gpu-torch-baseline 2>gpurun_out/${tag}_$name.err | tail -1 > gpurun_out/${tag}_$name.json; }
b c3_wide MMB_X=1
b c3_narrow MMB_GEMM_WIDEBOX=0
b c3_wide2 MMB_X=1
python - <<PY
import json
for w in ("c3_wide", "c3_narrow", "c3_wide2"):
    try:
        d = json.load(open(f"gpurun_out/${tag}_{w}.json"))
        print(w, round(d["value"], 1), round(d["ms_per_step"], 2), "gemm", round(d["roofline"]["achieved"], 1), "step_frac", round(d["roofline"]["step_frac"], 3), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(w, "failed:", e)
PY
