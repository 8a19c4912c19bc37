#!/bin/bash
tag=${1:-r2san}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 120 python scripts/sanitize_smoke.py 2>&1 | tail -3
for tool in memcheck synccheck racecheck; do
  echo "=== compute-sanitizer --tool $tool python scripts/sanitize_smoke.py" > gpurun_out/${tag}_$tool.txt
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python scripts/sanitize_smoke.py >> gpurun_out/${tag}_$tool.txt 2>&1
  echo "exit code $?" >> gpurun_out/${tag}_$tool.txt
  tail -12 gpurun_out/${tag}_$tool.txt | cut -c1-220
done
