#!/bin/bash
# GPU call S: grid softmin with the packed cost table — parity, 256^3 timing, sanitizer
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "grid or image or barycenter or img" 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_grid.json 2> gpurun_out/bench_grid.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_grid.json')); s=d['config']['secondary']; print(d['value']); print(s['cfg5_grid_256'])"
for tool in racecheck memcheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 6 python tools/sanitize_smoke.py grid > gpurun_out/sanitize_${tool}_grid.log 2>&1
  echo "$tool grid: $(grep 'RACECHECK SUMMARY\|ERROR SUMMARY' gpurun_out/sanitize_${tool}_grid.log | tail -1)"
done
