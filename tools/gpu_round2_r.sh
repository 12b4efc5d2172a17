#!/bin/bash
# GPU call R: ranges mode — a warp owns R*32 consecutive rows, warps without rows idle (cluster slivers, small clusters)
set -u
mkdir -p gpurun_out
python tools/run_multiscale_once.py 1000000
timeout 1200 python -m pytest tests/test_gpu_reference_goldens.py tests/test_gpu_parity.py -q -m gpu -x -k "multiscale or ranges or batched or keops or kernel or instantiations or online" 2>&1 | tail -4
for tool in racecheck memcheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 6 python tools/sanitize_smoke.py ranges loss > gpurun_out/sanitize_${tool}_ranges_loss.log 2>&1
  echo "$tool ranges+loss: $(grep 'RACECHECK SUMMARY\|ERROR SUMMARY' gpurun_out/sanitize_${tool}_ranges_loss.log | tail -1)"
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_ranges.json 2> gpurun_out/bench_ranges.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_ranges.json')); s=d['config']['secondary']; print(d['value']); print(s['cfg4_multiscale']); print(s['small_and_batched'])"
