#!/bin/bash
tag=${1:-r2h}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python scripts/bringup_gemm.py z_ragged_mul z_ffn2d_mul s_ffn2d_mul z_ffn1_gg 2>&1 | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.txt 2>&1
grep -E "^E  |^FAILED|passed|failed" gpurun_out/${tag}_pytest.txt | cut -c1-400 | head -30
timeout 300 python bench.py --no-cpu-baseline --no-gpu-torch-baseline 2>gpurun_out/${tag}_bench_c3.err | tail -1 > gpurun_out/${tag}_bench_c3.json
timeout 300 python bench.py --workload mosi_aligned_b64 --no-cpu-baseline --no-gpu-torch-baseline 2>gpurun_out/${tag}_bench_c2.err | tail -1 > gpurun_out/${tag}_bench_c2.json
for w in mosi_aligned_b64 mosei_unaligned_b64; do
timeout 120 python scripts/step_table.py $w > gpurun_out/${tag}_step_table_$w.txt 2>&1
done
python - <<PY
import json
for w in ("c3", "c2"):
    try:
        d = json.load(open(f"gpurun_out/${tag}_bench_{w}.json"))
        print(w, round(d["value"], 1), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1),
              "gemm", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 3), "step_frac", round(d["roofline"]["step_frac"], 3), d["clocks"])
    except Exception as e:
        print(w, "failed:", e)
PY
tail -n 5 gpurun_out/${tag}_bench_c3.err
head -32 gpurun_out/${tag}_step_table_mosei_unaligned_b64.txt
head -32 gpurun_out/${tag}_step_table_mosi_aligned_b64.txt
