#!/bin/bash
# end-of-round refresh: sanitizer over the extended smoke script, default bench line, inference sweep (BASELINE config 5)
tag=${1:-r2z}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 120 python scripts/sanitize_smoke.py 2>&1 | tail -3
for tool in memcheck synccheck; do
  echo "=== compute-sanitizer --tool $tool python scripts/sanitize_smoke.py" > gpurun_out/${tag}_$tool.txt
  timeout 500 compute-sanitizer --tool $tool --print-limit 30 python scripts/sanitize_smoke.py >> gpurun_out/${tag}_$tool.txt 2>&1
  echo "exit code $?" >> gpurun_out/${tag}_$tool.txt
  tail -6 gpurun_out/${tag}_$tool.txt | cut -c1-200
done
timeout 500 python bench.py 2>gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench_mosei_unaligned_b64.json
python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_mosei_unaligned_b64.json')); r=d['roofline']
print(round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'gemm', round(r['achieved'],1), r['frac'], r.get('frac_sustained'), r['peak_source'][:60], 'step', round(r['step_frac'],3), d['clocks'])"
timeout 600 python bench.py --mode infer-sweep 2>gpurun_out/${tag}_infer.err | tail -1 > gpurun_out/${tag}_infer_sweep.json
python -c "
import json; d=json.load(open('gpurun_out/${tag}_infer_sweep.json'))
print([(p['batch'], p['concat_len'], p['ms']) for p in d['points']])
print(d.get('cpu_reference'))"
