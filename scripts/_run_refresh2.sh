cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "schedule or work_lists" 2>&1 | tail -3 > gpurun_out/r1c_pytest_sched.txt
cat gpurun_out/r1c_pytest_sched.txt
timeout 100 python bench.py --workload mosi_aligned_b64 2>gpurun_out/r1c_bench_c2.err | tail -1 > gpurun_out/r1c_bench_c2.json
timeout 80 python bench.py --workload mosei_unaligned_b64 --no-cpu-baseline 2>gpurun_out/r1c_bench_c3.err | tail -1 > gpurun_out/r1c_bench_c3.json
head -c 300 gpurun_out/r1c_bench_c2.json; echo; head -c 300 gpurun_out/r1c_bench_c3.json; echo
tail -3 gpurun_out/r1c_bench_c2.err gpurun_out/r1c_bench_c3.err
