#!/bin/bash
# round 2, GPU call A: full GPU test-suite + first bench lines of the new default (MOSEI-unaligned) with both baselines.
tag=${1:-r2a}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
nproc >> gpurun_out/${tag}_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -40 > gpurun_out/${tag}_pytest.txt
cat gpurun_out/${tag}_pytest.txt | tail -25
timeout 600 python bench.py 2>gpurun_out/${tag}_bench_c3.err | tail -1 > gpurun_out/${tag}_bench_c3.json
timeout 300 python bench.py --workload mosi_aligned_b64 --no-cpu-baseline 2>gpurun_out/${tag}_bench_c2.err | tail -1 > gpurun_out/${tag}_bench_c2.json
for w in mosi_aligned_b64 mosei_unaligned_b64; do
  timeout 120 python scripts/step_table.py $w > gpurun_out/${tag}_step_table_$w.txt 2>&1
done
timeout 400 python bench.py --mode infer-sweep > gpurun_out/${tag}_infer_sweep.jsonl 2>gpurun_out/${tag}_infer_sweep.err
timeout 300 python bench.py --mode infer-sweep --no-cuda-graph --no-cpu-baseline --batches 1,4,16 --lengths 150,2048 > gpurun_out/${tag}_infer_sweep_nograph.jsonl 2>&1
python - <<PY
import json
for w in ("c3", "c2"):
    try:
        d = json.load(open(f"gpurun_out/${tag}_bench_{w}.json"))
        print(w, round(d["value"], 1), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), d["e2e"]["loop"],
              "gemm", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 3), "step_frac", round(d["roofline"]["step_frac"], 3))
        print("   torch:", {k: v for k, v in d.get("gpu_torch_baseline", {}).items() if k in ("autocast_bf16", "tf32", "fp32", "error", "ours_over_autocast_bf16")})
        print("   cpu:", d.get("cpu_baseline"))
    except Exception as e:
        print(w, "failed:", e)
PY
tail -3 gpurun_out/${tag}_bench_c3.err gpurun_out/${tag}_infer_sweep.err
tail -30 gpurun_out/${tag}_infer_sweep.jsonl | cut -c1-400
