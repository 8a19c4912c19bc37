set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 330 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_sched_pytest.txt
cat gpurun_out/r1_sched_pytest.txt
for w in mosei_unaligned_b64 mosi_aligned_b64; do
  for s in 0 1 0 1; do
    MMB_ATTN_SCHED=$s timeout 150 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r1_sched_${w}_${s}_$RANDOM.json
  done
done
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 > gpurun_out/r1_sched_smoke.txt
grep -h -o '"value": [0-9.]*\|"workload": "[a-z_0-9]*"' gpurun_out/r1_sched_mo*.json | paste - - 
ls gpurun_out | grep sched
