#!/bin/bash
# round 2, GPU call D: GPU suite, GEMM epilogue dissection at the MOSEI shapes.
tag=${1:-r2d}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/${tag}_pytest.txt 2>&1
grep -E "^E  |^FAILED|passed|failed" gpurun_out/${tag}_pytest.txt | cut -c1-600 | head -40
timeout 900 python scripts/bringup_gemm.py z_ffn1_gg z_ffn1_gg_nostore z_ffn1_gg_nold z_ffn2d_mul z_ffn2d_mul_nostore z_ffn2d_mul_nold z_qkv z_qkv_nostore z_qkv_nold z_ffn1_plain z_ffn2 z_ffn2_nold z_wo z_wo_nold 2>&1 | cut -c1-300
cp gpurun_out/bringup_gemm.json gpurun_out/${tag}_gemm_dissect.json
