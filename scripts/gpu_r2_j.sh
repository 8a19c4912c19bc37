#!/bin/bash
tag=${1:-r2j}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm" 2>&1 | tail -3
for lib in msa_b200/lib/libmmbert_gelu_as.so msa_b200/lib/libmmbert_sm100.so; do
  echo "== $lib"
  MMB_LIB=$PWD/$lib timeout 300 python scripts/bringup_gemm.py z_ffn1_gg z_c2_ffn1_gg 2>&1 | cut -c1-260
done
b() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-gpu-torch-baseline 2>gpurun_out/${tag}_$name.err | tail -1 > gpurun_out/${tag}_$name.json; }
b c3_new MMB_X=1
b c3_old MMB_LIB=$PWD/msa_b200/lib/libmmbert_gelu_as.so
b c3_new2 MMB_X=1
b c3_old2 MMB_LIB=$PWD/msa_b200/lib/libmmbert_gelu_as.so
python - <<PY
import json
for w in ("c3_new", "c3_old", "c3_new2", "c3_old2"):
    try:
        d = json.load(open(f"gpurun_out/${tag}_{w}.json"))
        print(w, round(d["value"], 1), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1),
              "gemm", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 3), "step_frac", round(d["roofline"]["step_frac"], 3), d["clocks"])
    except Exception as e:
        print(w, "failed:", e)
PY
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.txt 2>&1
grep -E "^E  |^FAILED|passed|failed" gpurun_out/${tag}_pytest.txt | cut -c1-400 | head -30
