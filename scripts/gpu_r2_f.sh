#!/bin/bash
# round 2, GPU call F: ncu --set full of the epilogue-bound GEMM shapes (stand-alone launches of scripts/bringup_gemm.py)
tag=${1:-r2f}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for c in z_ffn1_gg z_ffn2d_mul z_qkv z_ffn2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_2cta -s 3 -c 1 \
      -o gpurun_out/${tag}_$c -f python scripts/bringup_gemm.py --one $c > gpurun_out/${tag}_$c.log 2>&1
  tail -2 gpurun_out/${tag}_$c.log | cut -c1-200
done
ls -la gpurun_out | grep ${tag}_
