#!/bin/bash
# round 2, GPU call C: whole GPU suite with full failure output, bench + step table of the MOSEI shape.
tag=${1:-r2c}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 -x --deselect tests/test_model_gpu.py::test_bf16_forward_backward_12_layer_bert_base > gpurun_out/${tag}_pytest.txt 2>&1
tail -5 gpurun_out/${tag}_pytest.txt
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -k "12_layer or k_steps_bert_base or tracks_changing" > gpurun_out/${tag}_pytest2.txt 2>&1
grep -E "^E  |^FAILED|passed|failed" gpurun_out/${tag}_pytest2.txt | head -60
timeout 300 python bench.py --no-cpu-baseline --no-gpu-torch-baseline 2>gpurun_out/${tag}_bench_c3.err | tail -1 > gpurun_out/${tag}_bench_c3.json
MMB_ATTN_QSKIP=0 timeout 300 python bench.py --no-cpu-baseline --no-gpu-torch-baseline 2>gpurun_out/${tag}_bench_c3_noqskip.err | tail -1 > gpurun_out/${tag}_bench_c3_noqskip.json
timeout 120 python scripts/step_table.py mosei_unaligned_b64 > gpurun_out/${tag}_step_table_mosei_unaligned_b64.txt 2>&1
python - <<PY
import json
for w in ("c3", "c3_noqskip"):
    try:
        d = json.load(open(f"gpurun_out/${tag}_bench_{w}.json"))
        print(w, round(d["value"], 1), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1),
              "gemm", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 3), "step_frac", round(d["roofline"]["step_frac"], 3), d["clocks"])
    except Exception as e:
        print(w, "failed:", e)
PY
tail -n 5 gpurun_out/${tag}_bench_c3.err
head -12 gpurun_out/${tag}_step_table_mosei_unaligned_b64.txt
