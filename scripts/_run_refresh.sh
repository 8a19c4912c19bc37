set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r1b_pytest.txt
cat gpurun_out/r1b_pytest.txt
timeout 200 python bench.py --workload mosi_aligned_b64 2>gpurun_out/r1b_bench_c2.err | tail -1 > gpurun_out/r1b_bench_c2.json
timeout 200 python bench.py --workload mosei_unaligned_b64 --no-cpu-baseline 2>gpurun_out/r1b_bench_c3.err | tail -1 > gpurun_out/r1b_bench_c3.json
for w in mosi_aligned_b64 mosei_unaligned_b64; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/r1b_launches_$w.csv python bench.py --steps 2 --warmup 3 --workload $w --no-cpu-baseline > /dev/null 2>&1
  timeout 120 python scripts/step_table.py $w > gpurun_out/r1b_step_table_$w.txt 2>&1
done
head -c 400 gpurun_out/r1b_bench_c2.json; echo; head -c 400 gpurun_out/r1b_bench_c3.json; echo
tail -3 gpurun_out/r1b_bench_c2.err
ls -la gpurun_out | grep r1b
