#!/bin/bash
# One gpurun call that refreshes everything profiles/ is built from (≈ 6 GPU-minutes):
#   gpurun --timeout 560 -- 'bash scripts/gpu_refresh.sh r2'
# writes gpurun_out/<tag>_*: GPU test summary, bench lines for both headline shapes, ncu launch lists, per-launch step
# tables, and one `ncu --set full` capture of the GEMM launches of a step (source-level: -lineinfo is always on).
# Afterwards, here:  python scripts/summarize_launches.py gpurun_out/<tag>_launches_<workload>.csv
#                    python scripts/ncu_table.py gpurun_out/<tag>_full_gemm.ncu-rep
tag=${1:-r2}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${tag}_pytest.txt
cat gpurun_out/${tag}_pytest.txt
timeout 200 python bench.py --workload mosi_aligned_b64 2>gpurun_out/${tag}_bench_c2.err | tail -1 > gpurun_out/${tag}_bench_c2.json
timeout 200 python bench.py --workload mosei_unaligned_b64 --no-cpu-baseline 2>gpurun_out/${tag}_bench_c3.err | tail -1 > gpurun_out/${tag}_bench_c3.json
for w in mosi_aligned_b64 mosei_unaligned_b64; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv \
      --log-file gpurun_out/${tag}_launches_$w.csv python bench.py --steps 2 --warmup 3 --workload $w --no-cpu-baseline > /dev/null 2>&1
  timeout 120 python scripts/step_table.py $w > gpurun_out/${tag}_step_table_$w.txt 2>&1
done
# the 14 GEMM launches of one encoder layer's forward + backward at the MOSEI shape (skip the warm-up steps' launches)
timeout 240 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_2cta -s 400 -c 14 \
    -o gpurun_out/${tag}_full_gemm -f python bench.py --steps 1 --warmup 3 --workload mosei_unaligned_b64 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import json
for w in ("c2", "c3"):
    try:
        d = json.load(open(f"gpurun_out/${tag}_bench_{w}.json"))
        print(w, round(d["value"], 1), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), d["e2e"]["loop"],
              "gemm", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 3), "mfu", round(d["model_flops"]["frac_of_peak"], 3))
    except Exception as e:
        print(w, "failed:", e)
PY
ls -la gpurun_out | grep ${tag}_
