cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 70 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r1d_pytest.txt
cat gpurun_out/r1d_pytest.txt
for w in mosei_unaligned_b64 mosi_aligned_b64; do
  for m in warp lane; do
    MMB_GEMM_ISSUE=$m timeout 40 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r1d_${w}_${m}.json
    python -c "import json;d=json.load(open('gpurun_out/r1d_${w}_${m}.json'));print('$w $m',round(d['value'],1),round(d['roofline']['achieved'],1),d['final_loss'])"
  done
done
