#!/bin/bash
tag=${1:-r2m}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.txt 2>&1
grep -E "^E  |^FAILED|passed|failed" gpurun_out/${tag}_pytest.txt | cut -c1-300 | head -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-400
timeout 200 python scripts/bench_rowops.py 2>&1 | tail -8
timeout 300 python bench.py --no-cpu-baseline --no-gpu-torch-baseline 2>gpurun_out/${tag}_c3.err | tail -1 > gpurun_out/${tag}_c3.json
timeout 300 python bench.py --mode infer-sweep > gpurun_out/${tag}_infer_sweep.jsonl 2>gpurun_out/${tag}_infer_sweep.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_c3.json"))
print("c3", round(d["value"], 1), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "gemm", round(d["roofline"]["achieved"], 1), "step_frac", round(d["roofline"]["step_frac"], 3), d["clocks"])
PY
tail -30 gpurun_out/${tag}_infer_sweep.jsonl | cut -c1-300
tail -3 gpurun_out/${tag}_infer_sweep.err
