#!/bin/bash
tag=${1:-r2i}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm" > gpurun_out/${tag}_pytest_gemm.txt 2>&1
grep -E "^E  |^FAILED|passed|failed" gpurun_out/${tag}_pytest_gemm.txt | cut -c1-400 | head -20
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.txt 2>&1
grep -E "^E  |^FAILED|passed|failed" gpurun_out/${tag}_pytest.txt | cut -c1-400 | head -30
b() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-gpu-torch-baseline 2>gpurun_out/${tag}_$name.err | tail -1 > gpurun_out/${tag}_$name.json; }
b c3_fused MMB_X=1
b c3_sep MMB_GEMM_COLSUM=0
b c3_sep_st5 MMB_GEMM_COLSUM=0 MMB_GEMM_STAGES=5
b c3_fused2 MMB_X=1
timeout 120 python scripts/step_table.py mosei_unaligned_b64 > gpurun_out/${tag}_step_table_mosei_unaligned_b64.txt 2>&1
python - <<PY
import json
for w in ("c3_fused", "c3_sep", "c3_sep_st5", "c3_fused2"):
    try:
        d = json.load(open(f"gpurun_out/${tag}_{w}.json"))
        print(w, round(d["value"], 1), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1),
              "gemm", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 3), "step_frac", round(d["roofline"]["step_frac"], 3), d["clocks"])
    except Exception as e:
        print(w, "failed:", e)
PY
tail -n 5 gpurun_out/${tag}_c3_fused.err
head -12 gpurun_out/${tag}_step_table_mosei_unaligned_b64.txt
