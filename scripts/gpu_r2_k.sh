#!/bin/bash
tag=${1:-r2k}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python scripts/bringup_gemm.py z_ffn1_gg y_ffn1_gg_nostore y_ffn1_gg_mathonly y_ffn1_gelu_1stream y_ffn1_gg_8warps z_ffn1_plain z_qkv 2>&1 | cut -c1-260
timeout 200 python scripts/bench_rowops.py 2>&1 | tail -8
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.txt 2>&1
grep -E "^E  |^FAILED|passed|failed" gpurun_out/${tag}_pytest.txt | cut -c1-300 | head -30
timeout 300 python bench.py --no-cpu-baseline --no-gpu-torch-baseline 2>gpurun_out/${tag}_c3.err | tail -1 > gpurun_out/${tag}_c3.json
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_c3.json"))
print("c3", round(d["value"], 1), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "gemm", round(d["roofline"]["achieved"], 1), "step_frac", round(d["roofline"]["step_frac"], 3), d["clocks"])
PY
