cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 45 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/r1e_pytest.txt
cat gpurun_out/r1e_pytest.txt
for w in mosei_unaligned_b64 mosi_aligned_b64; do
    timeout 25 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r1e_${w}.json
    python -c "import json;d=json.load(open('gpurun_out/r1e_${w}.json'));print('$w',round(d['value'],1),round(d['roofline']['achieved'],1),d['final_loss'])"
done
