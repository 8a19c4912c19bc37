#!/bin/bash
# round 2, GPU call E: TMA-store epilogue: kernel tests, A/B per shape, whole suite, bench.
tag=${1:-r2e}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/${tag}_pytest_kernels.txt 2>&1
grep -E "^E  |^FAILED|passed|failed" gpurun_out/${tag}_pytest_kernels.txt | cut -c1-400 | head -20
timeout 900 python scripts/bringup_gemm.py z_ragged_gg z_ragged_bias z_ragged_mul z_vocab z_ffn1_gg s_ffn1_gg z_ffn2d_mul s_ffn2d_mul z_qkv s_qkv z_wo s_wo z_ffn2 s_ffn2 z_c2_qkv s_c2_qkv z_c2_ffn1_gg s_c2_ffn1_gg 2>&1 | cut -c1-300
cp gpurun_out/bringup_gemm.json gpurun_out/${tag}_gemm_tma_ab.json
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.txt 2>&1
grep -E "^E  |^FAILED|passed|failed" gpurun_out/${tag}_pytest.txt | cut -c1-400 | head -30
timeout 300 python bench.py --no-cpu-baseline --no-gpu-torch-baseline 2>gpurun_out/${tag}_bench_c3.err | tail -1 > gpurun_out/${tag}_bench_c3.json
MMB_GEMM_STORE=stg timeout 300 python bench.py --no-cpu-baseline --no-gpu-torch-baseline 2>gpurun_out/${tag}_bench_c3_stg.err | tail -1 > gpurun_out/${tag}_bench_c3_stg.json
timeout 300 python bench.py --workload mosi_aligned_b64 --no-cpu-baseline --no-gpu-torch-baseline 2>gpurun_out/${tag}_bench_c2.err | tail -1 > gpurun_out/${tag}_bench_c2.json
timeout 120 python scripts/step_table.py mosei_unaligned_b64 > gpurun_out/${tag}_step_table_mosei_unaligned_b64.txt 2>&1
python - <<PY
import json
for w in ("c3", "c3_stg", "c2"):
    try:
        d = json.load(open(f"gpurun_out/${tag}_bench_{w}.json"))
        print(w, round(d["value"], 1), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1),
              "gemm", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 3), "step_frac", round(d["roofline"]["step_frac"], 3), d["clocks"])
    except Exception as e:
        print(w, "failed:", e)
PY
tail -n 5 gpurun_out/${tag}_bench_c3.err
head -32 gpurun_out/${tag}_step_table_mosei_unaligned_b64.txt
