#!/bin/bash
# Regenerates the inputs of profiles/r2_* in ONE gpurun call (1 GPU, ~8 GPU-minutes):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_profile_r2.sh r2p'
# then here: python scripts/make_profiles_r2.py gpurun_out r2p
tag=${1:-r2p}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.txt 2>&1
tail -3 gpurun_out/${tag}_pytest.txt
# bench lines: the default run (both baselines), and the other two training shapes
timeout 600 python bench.py 2>gpurun_out/${tag}_bench_mosei_unaligned_b64.err | tail -1 > gpurun_out/${tag}_bench_mosei_unaligned_b64.json
for w in mosi_aligned_b64 ur_funny_b64; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline 2>gpurun_out/${tag}_bench_$w.err | tail -1 > gpurun_out/${tag}_bench_$w.json
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${tag}_bench_reference_arm.json
# per-launch CUDA-event tables and ncu launch lists (duration + DRAM bytes of every kernel of ONE step)
for w in mosei_unaligned_b64 mosi_aligned_b64 ur_funny_b64; do
  timeout 150 python scripts/step_table.py $w > gpurun_out/${tag}_step_table_$w.txt 2>&1
done
for w in mosei_unaligned_b64 mosi_aligned_b64; do
  timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      --profile-from-start off --csv --log-file gpurun_out/${tag}_launches_$w.csv python scripts/one_step.py $w > /dev/null 2>&1
done
# ncu --set full: the first encoder layer's forward kernels and the last layer's backward kernels at the MOSEI shape
K='regex:gemm_tcgen05_2cta|attn_fwd_ws|attn_bwd_ws|attn_bwd_prep|drln_fwd|drln_bwd|colsum_bf16'
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$K" -s 0 -c 8 \
    -o gpurun_out/${tag}_full_fwd -f python scripts/one_step.py mosei_unaligned_b64 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$K" -s 92 -c 15 \
    -o gpurun_out/${tag}_full_bwd -f python scripts/one_step.py mosei_unaligned_b64 > /dev/null 2>&1
timeout 200 python scripts/bench_rowops.py > gpurun_out/${tag}_rowops.txt 2>&1
python - <<PY
import json
for w in ("mosei_unaligned_b64", "mosi_aligned_b64", "ur_funny_b64"):
    try:
        d = json.load(open(f"gpurun_out/${tag}_bench_{w}.json"))
        print(w, round(d["value"], 1), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1),
              "gemm", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 3), "step_frac", round(d["roofline"]["step_frac"], 3),
              round(d["roofline"]["step_frac_executed"], 3), d["clocks"])
        print("   torch:", {k: v for k, v in d.get("gpu_torch_baseline", {}).items() if k in ("autocast_bf16", "tf32", "fp32", "error", "ours_over_autocast_bf16")})
        print("   cpu:", d.get("cpu_baseline"))
    except Exception as e:
        print(w, "failed:", e)
PY
ls -la gpurun_out | grep ${tag}_
